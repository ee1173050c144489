#!/bin/bash
# Secondary BASELINE workloads on one GPU at their per-GPU batch (configs[2..4]); results -> gpurun_out/wl_*.json
mkdir -p gpurun_out
for wl in sndcgan_celebahq128 resnet_lsun-bedroom128 biggan_imagenet128; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/wl_$wl.json 2> gpurun_out/wl_$wl.err
  tail -c 400 gpurun_out/wl_$wl.err
  python - <<PY
import json
try:
  d = json.load(open("gpurun_out/wl_$wl.json"))
  print("$wl", "images/s %.1f" % d["value"], "ms/step %.1f" % d["ms_per_step"], "e2e %.1f" % d["e2e"]["value"], "eval", d["eval"] and d["eval"]["fid_samples_per_sec"], "graph", d["config"]["cuda_graph"], d["losses"])
except Exception as e:
  print("$wl FAILED", e)
PY
done
