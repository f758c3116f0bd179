mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_vae.py tests/test_gpu_pipeline.py -q -m gpu > gpurun_out/c19_tests.log 2>&1; echo "exit $?" >> gpurun_out/c19_tests.log
timeout 100 python tools/gpu_check_kernels.py perf_conv96_fullres > gpurun_out/c19_conv.log 2>&1
timeout 200 python tools/vae_bench.py > gpurun_out/c19_vae.log 2>&1
grep -h "passed\|failed\|^exit" gpurun_out/c19_tests.log | tail -3
tail -qn1 gpurun_out/c19_conv.log gpurun_out/c19_vae.log | cut -c1-260
