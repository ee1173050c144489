# re-entry baseline: whole GPU suite, default bench line, per-kernel launch list of the cifar cycle
timeout 900 python -m pytest tests/ -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r10_t_all.log
timeout 600 python bench.py > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
CGAN_PROFILE_RANGE=1 timeout 400 $NCU --log-file gpurun_out/r10_launches_cifar.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eval --eager --headline-only > gpurun_out/r10_ncu_cifar.log 2>&1
tail -n 4 gpurun_out/r10_t_all.log; cut -c1-1500 gpurun_out/r10_bench.json
