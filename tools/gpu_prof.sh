set -x
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train-record --profile-step > gpurun_out/ncu_l.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train_r02_final.csv python tools/prof_train.py > gpurun_out/ncu_tl.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_ts_kernel -s 100 -c 8 -f -o gpurun_out/prof_gemm_ts_r02_final python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train-record --profile-step > gpurun_out/ncu_g2.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tcgen05 -c 2 -f -o gpurun_out/prof_attn_r02_final python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train-record --profile-step > gpurun_out/ncu_at.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:row_layernorm_reg -s 4 -c 2 -f -o gpurun_out/prof_ln_r02_final python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train-record --profile-step > gpurun_out/ncu_ln.log 2>&1
ls -la gpurun_out/*r02_final*
