#!/bin/bash
# ncu launch lists (gpu__time_duration per launch, cold-cache and serialised: compare SHARES) of one timed cycle of the
# two headline workloads and of one FID evaluation batch.  Output: gpurun_out/launches_{cifar,biggan,eval}.csv
mkdir -p gpurun_out
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
CGAN_PROFILE_RANGE=1 timeout 600 $NCU --log-file gpurun_out/launches_cifar.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eval --eager > gpurun_out/ncu_cifar.log 2>&1
timeout 300 $NCU --log-file gpurun_out/launches_eval.csv python profiles/prof_eval.py > gpurun_out/ncu_eval.log 2>&1
CGAN_PROFILE_RANGE=1 timeout 900 $NCU --log-file gpurun_out/launches_biggan.csv python bench.py --workload biggan_imagenet128 --steps 1 --warmup 3 --no-cpu-baseline --no-eval --eager > gpurun_out/ncu_biggan.log 2>&1
wc -l gpurun_out/launches_*.csv; tail -n 3 gpurun_out/ncu_eval.log
