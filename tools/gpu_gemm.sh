set -x
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm --csv --log-file gpurun_out/ab_gemm.csv python tools/ab_gemm_once.py tile persist > gpurun_out/ab_gemm.log 2>&1
tail -2 gpurun_out/ab_gemm.log | cut -c1-2000
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/ab_gemm.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[4][:60], r[-1])
PY
