# ring-pipelined attention kernels: parity, timings, tests
timeout 120 python profiles/attn_check.py bwd > gpurun_out/r15_attn_bwd.txt 2>&1; echo "bwd rc=$?" >> gpurun_out/r15_attn_bwd.txt
timeout 200 python profiles/attn_check.py time > gpurun_out/r15_attn_time.txt 2>&1; echo "time rc=$?" >> gpurun_out/r15_attn_time.txt
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -8 > gpurun_out/r15_t_attn.log
timeout 300 python -m pytest tests/test_tf32_parity_gpu.py -m gpu -q -s -k "biggan" 2>&1 | grep -v "^$" | cut -c1-300 | tail -8 > gpurun_out/r15_t_biggan.log
cat gpurun_out/r15_attn_bwd.txt gpurun_out/r15_attn_time.txt | cut -c1-200; tail -4 gpurun_out/r15_t_attn.log; tail -6 gpurun_out/r15_t_biggan.log
