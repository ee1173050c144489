#!/bin/bash
# End-of-round refresh on one B200: full GPU tests, the default bench line, ncu evidence for the dominant kernels,
# the secondary workloads, and the launch list of one resnet_cifar10 cycle.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu --timeout 150 --timeout-method=thread > gpurun_out/t_final.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/t_final.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad_tc_kernel" -c 18 -f -o gpurun_out/prof_tc_final python profiles/prof_kernels.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
bash profiles/run_workloads.sh
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
CGAN_PROFILE_RANGE=1 timeout 420 $NCU --log-file gpurun_out/launches_cifar_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eval --eager > gpurun_out/ncu_cifar_final.log 2>&1; echo "launch list rc=$?"
wc -l gpurun_out/launches_cifar_final.csv
python -c "
import json; d = json.load(open('gpurun_out/bench_final.json')); print({k: d[k] for k in ('value','ms_per_step','e2e','gpu_launches','roofline','cpu_baseline','eval','clocks')})"
