python -m pytest tests/ -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t_all.log
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
CGAN_PROFILE_RANGE=1 timeout 600 $NCU --log-file gpurun_out/r2_launches_cifar.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eval --eager --headline-only > gpurun_out/r2_ncu_cifar.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
tail -n 4 gpurun_out/r2_t_all.log
