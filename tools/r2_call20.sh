mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_vae.py tests/test_gpu_pipeline.py -q -m gpu > gpurun_out/c20_tests.log 2>&1; echo "exit $?" >> gpurun_out/c20_tests.log
timeout 200 python tools/gpu_check_kernels.py perf_conv > gpurun_out/c20_conv.log 2>&1
timeout 200 python tools/vae_bench.py > gpurun_out/c20_vae.log 2>&1
grep -h "passed\|failed\|^exit" gpurun_out/c20_tests.log | tail -3
grep -h "perf_conv" gpurun_out/c20_conv.log | cut -c1-260; tail -qn1 gpurun_out/c20_vae.log | cut -c1-260
