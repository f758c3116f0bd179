# Round-2 1-GPU call 14: multi-frame 3x3x3 conv
mkdir -p gpurun_out
ICB_CONV_MF=1 timeout 300 python -m pytest tests/test_gpu_vae.py tests/test_gpu_pipeline.py -q -m gpu > gpurun_out/c14_tests_mf1.log 2>&1; echo "exit $?" >> gpurun_out/c14_tests_mf1.log
ICB_CONV_MF=0 timeout 100 python tools/gpu_check_kernels.py perf_conv96_fullres > gpurun_out/c14_conv_mf0.log 2>&1
ICB_CONV_MF=1 timeout 100 python tools/gpu_check_kernels.py perf_conv96_fullres > gpurun_out/c14_conv_mf1.log 2>&1
ICB_CONV_MF=0 timeout 200 python tools/vae_bench.py > gpurun_out/c14_vae_mf0.log 2>&1
ICB_CONV_MF=1 timeout 200 python tools/vae_bench.py > gpurun_out/c14_vae_mf1.log 2>&1
grep -h "passed\|failed\|^exit\|Error\|assert" gpurun_out/c14_tests_mf1.log | tail -6
tail -qn1 gpurun_out/c14_conv_mf0.log gpurun_out/c14_conv_mf1.log gpurun_out/c14_vae_mf0.log gpurun_out/c14_vae_mf1.log | cut -c1-260
