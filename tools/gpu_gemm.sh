set -x
ncu --set full --clock-control none --import-source on -k regex:gemm_ts -s 7 -c 7 -f -o gpurun_out/prof_gemm_ts_r02 python tools/prof_gemm_ts.py > gpurun_out/ncu_gts.log 2>&1
tail -3 gpurun_out/ncu_gts.log
