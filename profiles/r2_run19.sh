# two GPUs: NCCL gradient all-reduce + NVLink peer-memory BN exchange inside the captured cycle, in-run dp_equivalence
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --headline-only --dp-check --no-eval > gpurun_out/r19_bench2.json 2> gpurun_out/r19_bench2.err
echo "rc=$?" >> gpurun_out/r19_bench2.err
timeout 200 python profiles/microbench.py 2>&1 | grep -v "^\[{" | grep -E "256->3 |3->128" | cut -c1-250 > gpurun_out/r19_micro_thin.txt
cut -c1-400 gpurun_out/r19_bench2.json; python -c "
import json; d=json.loads(open('gpurun_out/r19_bench2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['dp_equivalence'])"; tail -4 gpurun_out/r19_bench2.err; cat gpurun_out/r19_micro_thin.txt
