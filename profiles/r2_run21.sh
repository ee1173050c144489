# attention with P / dS kept in TMEM (tcgen05.st + TMEM-A tcgen05.mma); S3GAN on the device
timeout 120 python profiles/attn_check.py bwd > gpurun_out/r21_attn_bwd.txt 2>&1; echo "bwd rc=$?" >> gpurun_out/r21_attn_bwd.txt
timeout 200 python profiles/attn_check.py time > gpurun_out/r21_attn_time.txt 2>&1; echo "time rc=$?" >> gpurun_out/r21_attn_time.txt
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention or random_uniform" 2>&1 | tail -8 > gpurun_out/r21_t_attn.log
timeout 300 python -m pytest tests/test_s3gan.py tests/test_ssgan.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r21_t_s3gan.log
timeout 300 python -m pytest tests/test_tf32_parity_gpu.py -m gpu -q -k "biggan" 2>&1 | tail -6 > gpurun_out/r21_t_biggan.log
grep -c "err" gpurun_out/r21_attn_bwd.txt; grep "err" gpurun_out/r21_attn_bwd.txt | sort -k3 -g | tail -4; tail -2 gpurun_out/r21_attn_bwd.txt; cat gpurun_out/r21_attn_time.txt | cut -c1-200; tail -3 gpurun_out/r21_t_attn.log; tail -3 gpurun_out/r21_t_s3gan.log; tail -3 gpurun_out/r21_t_biggan.log
