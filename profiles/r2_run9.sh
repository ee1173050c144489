# coalescing epilogue: parity (variants incl. the per-thread-row epilogue, network in-situ checks), microbench A/B, headline
timeout 400 python -m pytest tests/test_tf32_parity_gpu.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r9_t_tf32.log
timeout 200 python profiles/microbench.py 2>&1 | grep -v "^\[{" | head -22 | cut -c1-250 > gpurun_out/r9_micro_epi1.txt
CGAN_TC_EPI=0 timeout 200 python profiles/microbench.py 2>&1 | grep -v "^\[{" | head -15 | cut -c1-250 > gpurun_out/r9_micro_epi0.txt
timeout 600 python bench.py --headline-only > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err
cat gpurun_out/r9_t_tf32.log; cat gpurun_out/r9_bench.json | cut -c1-600
