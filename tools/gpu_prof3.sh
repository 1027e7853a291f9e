set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r02_s3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train-record --profile-step > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tcgen05 -c 2 -f -o gpurun_out/prof_attn_r02_s3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train-record --profile-step > gpurun_out/ncu_at.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train_r02_s3.csv python tools/prof_train.py > gpurun_out/ncu_tl.log 2>&1
ls -la gpurun_out/*r02_s3*
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
timeout 600 python bench.py --config 4 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
timeout 900 python bench.py --config 5 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
python - <<'PY'
import json
for n in ("cfg3", "cfg4", "cfg5"):
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, d["ms_per_step"], d["value"], d["roofline"]["step"])
    except Exception as e:
        print(n, "failed", e)
PY
