CGAN_TC_PAIR_MT=1 timeout 300 python -m pytest tests/test_tf32_parity_gpu.py -m gpu -q -k "baseline" 2>&1 | tail -5 > gpurun_out/r8_t_pair_mt1.log
CGAN_TC_PAIR=1 CGAN_TC_PAIR_MT=1 timeout 200 python profiles/microbench.py 2>&1 | grep -v "^\[{" | head -15 | cut -c1-250 > gpurun_out/r8_micro_pair_mt1.txt
CGAN_TC_PAIR=1 CGAN_TC_PAIR_MT=2 timeout 200 python profiles/microbench.py 2>&1 | grep -v "^\[{" | head -15 | cut -c1-250 > gpurun_out/r8_micro_pair_mt2.txt
cat gpurun_out/r8_t_pair_mt1.log
