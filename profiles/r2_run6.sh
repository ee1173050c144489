timeout 300 python -m pytest tests/test_tf32_parity_gpu.py -m gpu -q -s -k "baseline" --maxfail=4 2>&1 | grep -v "^$" | cut -c1-300 | tail -40 > gpurun_out/r6_t_pair.log
echo "pair test rc=$?" >> gpurun_out/r6_t_pair.log
nvidia-smi --query-gpu=name,memory.used --format=csv >> gpurun_out/r6_t_pair.log 2>&1
timeout 300 python -m pytest tests/test_tf32_parity_gpu.py -m gpu -q -s -k network_parity 2>&1 | grep -v "^$" | cut -c1-400 | tail -30 > gpurun_out/r6_t_net.log
CGAN_TC_PAIR=1 timeout 200 python profiles/microbench.py > gpurun_out/r6_micro_pair.txt 2>&1
timeout 200 python profiles/microbench.py > gpurun_out/r6_micro.txt 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --headline-only --no-cpu-baseline > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.err
CGAN_TC_PAIR=1 timeout 400 python bench.py --steps 10 --warmup 3 --headline-only --no-cpu-baseline --no-eval > gpurun_out/r6_bench_pair.json 2> gpurun_out/r6_bench_pair.err
tail -n 6 gpurun_out/r6_t_pair.log gpurun_out/r6_t_net.log
