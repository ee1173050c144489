# final state: smoke, the driver's default bench line, the two other BASELINE workloads
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r20_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r20_smoke.log
timeout 900 python bench.py > gpurun_out/r20_bench.json 2> gpurun_out/r20_bench.err
timeout 400 python bench.py --workload sndcgan_celebahq128 --steps 5 --warmup 3 --headline-only --no-cpu-baseline > gpurun_out/r20_bench_sndcgan.json 2> gpurun_out/r20_bench_sndcgan.err
timeout 400 python bench.py --workload resnet_lsun-bedroom128 --steps 5 --warmup 3 --headline-only --no-cpu-baseline > gpurun_out/r20_bench_lsun.json 2> gpurun_out/r20_bench_lsun.err
tail -2 gpurun_out/r20_smoke.log; for f in bench bench_sndcgan bench_lsun; do python - <<PY
import json
d=json.loads(open('gpurun_out/r20_$f.json').read().strip().splitlines()[-1])
print('$f', d['metric'], round(d['ms_per_step'],2), round(d['value']), (d.get('eval') or {}).get('fid_samples_per_sec'), d['roofline']['frac'] if d.get('roofline') else None, {k:(round(v.get('ms_per_step',0),1), round(v.get('value',0))) for k,v in (d.get('workloads') or {}).items()})
PY
done
