# image-side convolutions through patch tensors on tcgen05 (thin_tc.cu): parity, A/B microbench, network parity, headline
timeout 400 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "thin or conv or tcgen05 or tc_" 2>&1 | tail -12 > gpurun_out/r17_t_kernels.log
timeout 200 python profiles/microbench.py 2>&1 | grep -v "^\[{" | grep -E "256->3 |3->128|3x3 256->256 @32|bn_train" | cut -c1-250 > gpurun_out/r17_micro_thin1.txt
CGAN_TC_THIN=0 timeout 200 python profiles/microbench.py 2>&1 | grep -v "^\[{" | grep -E "256->3 |3->128" | cut -c1-250 > gpurun_out/r17_micro_thin0.txt
timeout 500 python -m pytest tests/test_tf32_parity_gpu.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r17_t_tf32.log
timeout 400 python bench.py --headline-only --no-cpu-baseline --no-eval > gpurun_out/r17_bench.json 2> gpurun_out/r17_bench.err
tail -6 gpurun_out/r17_t_kernels.log; cat gpurun_out/r17_micro_thin1.txt gpurun_out/r17_micro_thin0.txt; tail -5 gpurun_out/r17_t_tf32.log; cut -c1-330 gpurun_out/r17_bench.json; tail -2 gpurun_out/r17_bench.err
