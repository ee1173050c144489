timeout 170 python -m pytest tests/test_gan_step_gpu.py -m gpu -q -s 2>&1 | grep -E "flips over|passed|failed|Error|error" | cut -c1-200 > gpurun_out/r24_t_step.log
cat gpurun_out/r24_t_step.log
