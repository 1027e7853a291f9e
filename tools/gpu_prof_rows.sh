set -x
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train_r01_rows.csv python tools/prof_train.py > gpurun_out/ncu_tl.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:row_layernorm_bwd_reg|gelu_ln_bwd_wide|row_layernorm_reg|gelu_ln_fwd_wide|transpose_cast" -s 60 -c 8 -f -o gpurun_out/prof_rows_r01 python tools/prof_train.py > gpurun_out/ncu_rows.log 2>&1
ls -la gpurun_out/prof_rows_r01*
