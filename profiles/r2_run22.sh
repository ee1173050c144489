timeout 120 python profiles/attn_check.py bwd > gpurun_out/r22_attn_bwd.txt 2>&1; echo "bwd rc=$?" >> gpurun_out/r22_attn_bwd.txt
timeout 200 python profiles/attn_check.py time > gpurun_out/r22_attn_time.txt 2>&1
grep "err" gpurun_out/r22_attn_bwd.txt | sort -k3 -g | tail -2; tail -1 gpurun_out/r22_attn_bwd.txt; cat gpurun_out/r22_attn_time.txt | cut -c1-200
