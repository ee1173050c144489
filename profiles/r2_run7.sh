timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_tc|wgrad_tc" -o gpurun_out/r2_prof_tc python profiles/prof_kernels.py > gpurun_out/r7_ncu.log 2>&1
ls -la gpurun_out/r2_prof_tc.ncu-rep >> gpurun_out/r7_ncu.log
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
timeout 300 $NCU --log-file gpurun_out/r7_launches_eval.csv python profiles/prof_eval.py 256 > gpurun_out/r7_ncu_eval.log 2>&1
CGAN_PROFILE_RANGE=1 timeout 600 $NCU --log-file gpurun_out/r7_launches_biggan.csv python bench.py --workload biggan_imagenet128 --steps 1 --warmup 3 --no-cpu-baseline --no-eval --eager --headline-only > gpurun_out/r7_ncu_biggan.log 2>&1
du -sh gpurun_out; tail -n 3 gpurun_out/r7_ncu.log gpurun_out/r7_ncu_eval.log
