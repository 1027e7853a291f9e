set -x
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:gelu_ln_bwd_wide2" -s 2 -c 1 -f -o gpurun_out/prof_wide2_r01 python tools/prof_train.py > gpurun_out/ncu_rows.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:row_layernorm_bwd_reg" -s 6 -c 2 -f -o gpurun_out/prof_rowbwd_r01 python tools/prof_train.py > gpurun_out/ncu_rows2.log 2>&1
ls -la gpurun_out/prof_wide2_r01* gpurun_out/prof_rowbwd_r01*
