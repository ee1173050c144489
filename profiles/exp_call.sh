#!/bin/bash
# one-off experiment driver (A/B of the multi-tile tcgen05 kernels); results in gpurun_out/
mkdir -p gpurun_out
timeout 420 python -m pytest tests -q -m gpu --timeout 150 --timeout-method=thread > gpurun_out/t_exp.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/t_exp.log
if grep -q "Timeout" gpurun_out/t_exp.log; then echo "TIMEOUT in tests: skipping benches"; exit 1; fi
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/exp_cifar_mt2.json 2> gpurun_out/exp_cifar_mt2.err; echo "cifar mt2 rc=$?"
CGAN_TC_MT=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eval > gpurun_out/exp_cifar_mt1.json 2> gpurun_out/exp_cifar_mt1.err; echo "cifar mt1 rc=$?"
timeout 300 python bench.py --workload biggan_imagenet128 --steps 3 --warmup 3 --no-cpu-baseline --no-eval > gpurun_out/exp_biggan_mt2.json 2> gpurun_out/exp_biggan_mt2.err; echo "biggan mt2 rc=$?"
python - <<'PY'
import json
for f in ["exp_cifar_mt2", "exp_cifar_mt1", "exp_biggan_mt2"]:
  try:
    d = json.load(open("gpurun_out/%s.json" % f))
    print(f, "img/s %.0f  ms %.2f  e2e %.0f  kernel_ms %.4f frac %.3f  eval %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["eval"] and d["eval"]["fid_samples_per_sec"]))
  except Exception as e:
    print(f, "FAILED", e)
PY
