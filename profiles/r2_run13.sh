# 8-softmax-warp attention kernels: parity, timings, BigGAN network parity (fused path inside a network), BigGAN cycle
timeout 120 python profiles/attn_check.py bwd > gpurun_out/r13_attn_bwd.txt 2>&1; echo "bwd rc=$?" >> gpurun_out/r13_attn_bwd.txt
timeout 200 python profiles/attn_check.py time > gpurun_out/r13_attn_time.txt 2>&1; echo "time rc=$?" >> gpurun_out/r13_attn_time.txt
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -8 > gpurun_out/r13_t_attn.log
timeout 400 python -m pytest tests/test_tf32_parity_gpu.py tests/test_gan_step_gpu.py -m gpu -q -s -k "biggan" 2>&1 | grep -v "^$" | cut -c1-300 | tail -14 > gpurun_out/r13_t_biggan.log
timeout 500 python bench.py --workload biggan_imagenet128 --steps 3 --warmup 3 --headline-only --no-cpu-baseline --no-eval > gpurun_out/r13_bench_biggan.json 2> gpurun_out/r13_bench_biggan.err
cat gpurun_out/r13_attn_bwd.txt gpurun_out/r13_attn_time.txt | cut -c1-200; tail -4 gpurun_out/r13_t_attn.log; tail -8 gpurun_out/r13_t_biggan.log; cut -c1-400 gpurun_out/r13_bench_biggan.json; tail -3 gpurun_out/r13_bench_biggan.err
