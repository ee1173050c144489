# attention: integer TF32 rounding, forward with two CTAs per SM (default) vs one (CGAN_ATTN_CTAS=1); ncu capture of the result
timeout 120 python profiles/attn_check.py bwd > gpurun_out/r16_attn_bwd.txt 2>&1; echo "bwd rc=$?" >> gpurun_out/r16_attn_bwd.txt
timeout 200 python profiles/attn_check.py time > gpurun_out/r16_attn_time.txt 2>&1; echo "time rc=$?" >> gpurun_out/r16_attn_time.txt
CGAN_ATTN_CTAS=1 timeout 200 python profiles/attn_check.py time > gpurun_out/r16_attn_time_1cta.txt 2>&1
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -8 > gpurun_out/r16_t_attn.log
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_" -o gpurun_out/r2_prof_attn2 python profiles/prof_attention.py > gpurun_out/r16_ncu.log 2>&1
grep -c "err" gpurun_out/r16_attn_bwd.txt; grep "err" gpurun_out/r16_attn_bwd.txt | sort -k3 | tail -3; cat gpurun_out/r16_attn_time.txt gpurun_out/r16_attn_time_1cta.txt | cut -c1-200; tail -3 gpurun_out/r16_t_attn.log
