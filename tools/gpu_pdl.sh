python bench.py --no-cpu-baseline --no-train-record 2>/dev/null | cut -c1-190
SGF_NO_PDL=1 python bench.py --no-cpu-baseline --no-train-record 2>/dev/null | cut -c1-190
python bench.py --no-cpu-baseline --no-train-record --no-graph 2>/dev/null | cut -c1-190
