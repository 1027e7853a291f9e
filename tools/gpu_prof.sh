set -x
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train_r01.csv python tools/prof_train.py > gpurun_out/ncu_tl.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:row_layernorm_bwd -s 20 -c 4 -f -o gpurun_out/prof_rowbwd_r01 python tools/prof_train.py > gpurun_out/ncu_rb.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_bwd -s 24 -c 4 -f -o gpurun_out/prof_attnbwd_r01 python tools/prof_train.py > gpurun_out/ncu_ab.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:upsample_ce_bwd -c 1 -f -o gpurun_out/prof_cebwd_r01 python tools/prof_train.py > gpurun_out/ncu_ce.log 2>&1
tail -3 gpurun_out/ncu_*.log
