for i in 1 2; do
python bench.py --no-cpu-baseline --no-train-record 2>/dev/null | cut -c1-190
SGF_LN_GRID=1 python bench.py --no-cpu-baseline --no-train-record 2>/dev/null | cut -c1-190
done
