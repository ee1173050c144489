timeout 1000 python -m pytest tests/ -m gpu -q 2>&1 | tail -12 > gpurun_out/r23_t_all.log
tail -n 6 gpurun_out/r23_t_all.log
