NCU="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
CGAN_PROFILE_RANGE=1 timeout 800 $NCU --log-file gpurun_out/r14_launches_biggan.csv python bench.py --workload biggan_imagenet128 --steps 1 --warmup 3 --no-cpu-baseline --no-eval --eager --headline-only > gpurun_out/r14_ncu_biggan.log 2>&1
tail -3 gpurun_out/r14_ncu_biggan.log | cut -c1-300
