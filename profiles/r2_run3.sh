python -m pytest tests/test_tf32_parity_gpu.py -m gpu --maxfail=6 -q -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r3_t_tf32.log
python -m pytest tests/test_kernels_gpu.py tests/test_gan_step_gpu.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r3_t_rest.log
python profiles/microbench.py > gpurun_out/r3_micro.txt 2>&1
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
CGAN_PROFILE_RANGE=1 timeout 600 $NCU --log-file gpurun_out/r3_launches_cifar.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eval --eager --headline-only > gpurun_out/r3_ncu_cifar.log 2>&1
python bench.py --steps 10 --warmup 3 --headline-only > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err
tail -n 3 gpurun_out/r3_t_tf32.log gpurun_out/r3_t_rest.log
