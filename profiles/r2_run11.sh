# fused attention kernels: per-kernel errors (separate processes so that a trap in one stage does not hide the others), timings, tests
timeout 120 python profiles/attn_check.py fwd > gpurun_out/r11_attn_fwd.txt 2>&1; echo "fwd rc=$?" >> gpurun_out/r11_attn_fwd.txt
timeout 120 python profiles/attn_check.py bwd > gpurun_out/r11_attn_bwd.txt 2>&1; echo "bwd rc=$?" >> gpurun_out/r11_attn_bwd.txt
timeout 200 python profiles/attn_check.py time > gpurun_out/r11_attn_time.txt 2>&1; echo "time rc=$?" >> gpurun_out/r11_attn_time.txt
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -15 > gpurun_out/r11_t_attn.log
cat gpurun_out/r11_attn_fwd.txt gpurun_out/r11_attn_bwd.txt gpurun_out/r11_attn_time.txt | cut -c1-300; tail -5 gpurun_out/r11_t_attn.log
