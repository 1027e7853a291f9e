set -x
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile-step > gpurun_out/ncu_l.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train_r01_final.csv python tools/prof_train.py > gpurun_out/ncu_tl.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm2p -c 2 -f -o gpurun_out/prof_gemm2p_r01_final python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile-step > gpurun_out/ncu_g2.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tcgen05 -c 2 -f -o gpurun_out/prof_attn_r01_final python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile-step > gpurun_out/ncu_at.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:gemm_tcgen05_kernel<128" -s 40 -c 3 -f -o gpurun_out/prof_gemm1_r01_final python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile-step > gpurun_out/ncu_g1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_bwd -s 24 -c 2 -f -o gpurun_out/prof_attnbwd_r01_final python tools/prof_train.py > gpurun_out/ncu_ab.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:gemm_mm|gelu_ln_bwd_wide2|row_layernorm_bwd_reg|row_layernorm_reg|upsample_ce_bwd" -s 40 -c 6 -f -o gpurun_out/prof_trainmisc_r01_final python tools/prof_train.py > gpurun_out/ncu_tm.log 2>&1
ls -la gpurun_out/*final*
