timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_" -o gpurun_out/r2_prof_attn python profiles/prof_attention.py > gpurun_out/r12_ncu.log 2>&1
ls -la gpurun_out/r2_prof_attn.ncu-rep >> gpurun_out/r12_ncu.log; tail -3 gpurun_out/r12_ncu.log
