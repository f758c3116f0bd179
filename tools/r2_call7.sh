# Round-2 1-GPU call 7: padded 64-channel conv A/B, full suite on the new defaults, the default bench line
mkdir -p gpurun_out
ICB_CONV_PAD64=1 timeout 300 python -m pytest tests/test_gpu_vae.py tests/test_gpu_pipeline.py -q -m gpu -s > gpurun_out/c7_tests_pad64.log 2>&1; echo "exit $?" >> gpurun_out/c7_tests_pad64.log
ICB_CONV_PAD64=0 timeout 200 python tools/vae_bench.py > gpurun_out/c7_vae_pad0.log 2>&1
ICB_CONV_PAD64=1 timeout 200 python tools/vae_bench.py > gpurun_out/c7_vae_pad1.log 2>&1
timeout 500 python -m pytest tests -q -m gpu > gpurun_out/c7_tests.log 2>&1; echo "exit $?" >> gpurun_out/c7_tests.log
timeout 600 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/c7_bench.err
grep -h "passed\|failed\|^exit" gpurun_out/c7_tests_pad64.log gpurun_out/c7_tests.log | tail -4
tail -qn1 gpurun_out/c7_vae_pad0.log gpurun_out/c7_vae_pad1.log | cut -c1-200
grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"call_ms": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2_bench_1gpu.json | tr '\n' ' '
