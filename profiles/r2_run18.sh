# whole GPU suite after the thin-TC dispatch + attention rings; launch list of the cifar cycle
timeout 1200 python -m pytest tests/ -m gpu -q 2>&1 | tail -15 > gpurun_out/r18_t_all.log
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
CGAN_PROFILE_RANGE=1 timeout 400 $NCU --log-file gpurun_out/r18_launches_cifar.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eval --eager --headline-only > gpurun_out/r18_ncu_cifar.log 2>&1
tail -n 8 gpurun_out/r18_t_all.log
