mkdir -p gpurun_out
ICB_FMHA_PERSIST=1 timeout 300 python -m pytest tests/test_gpu_dit.py tests/test_gpu_pipeline.py -q -m gpu > gpurun_out/c21_tests.log 2>&1; echo "exit $?" >> gpurun_out/c21_tests.log
ICB_FMHA_PERSIST=0 timeout 200 python bench.py --skip-e2e --skip-parity --skip-raster --steps 8 > gpurun_out/c21_bench_p0.json 2> gpurun_out/c21_bench_p0.err
ICB_FMHA_PERSIST=1 timeout 200 python bench.py --skip-e2e --skip-parity --skip-raster --steps 8 > gpurun_out/c21_bench_p1.json 2> gpurun_out/c21_bench_p1.err
grep -h "passed\|failed\|^exit\|Error" gpurun_out/c21_tests.log | tail -4
for f in gpurun_out/c21_bench_p0.json gpurun_out/c21_bench_p1.json; do grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"fmha_cross": [0-9.]*' $f | head -3 | tr '\n' ' '; echo; done
