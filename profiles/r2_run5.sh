CGAN_TEST_VERBOSE=1 python -m pytest tests/test_tf32_parity_gpu.py -m gpu -q -s -k network_parity 2>&1 | grep -v "^$" | grep -v "^   generator\|^   discriminator" | cut -c1-400 > gpurun_out/r5_t_tf32.log
python -m pytest tests/test_kernels_gpu.py tests/test_gan_step_gpu.py tests/test_ssgan.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r5_t_rest.log
tail -n 5 gpurun_out/r5_t_tf32.log gpurun_out/r5_t_rest.log
