# Round-2 1-GPU call 12: validation of the committed state (suite, smoke, both bench arms, library bars incl. conv)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/c12_tests.log 2>&1; echo "exit $?" >> gpurun_out/c12_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c12_smoke.log 2>&1; echo "exit $?" >> gpurun_out/c12_smoke.log
timeout 600 python tools/gpu_check_kernels.py perf_ > gpurun_out/c12_kernel_perf.log 2>&1; cp gpurun_out/kernel_check.json gpurun_out/c12_library_bars.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/c12_bench_ref.err
timeout 600 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/c12_bench.err
tail -3 gpurun_out/c12_tests.log; tail -2 gpurun_out/c12_smoke.log
grep -h "perf_conv\|perf_fmha" gpurun_out/c12_kernel_perf.log | cut -c1-300
cut -c1-300 gpurun_out/r2_bench_reference.json
grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"call_ms": [0-9.]*\|"frac": [0-9.]*\|"render_ms_93cams": [0-9.]*' gpurun_out/r2_bench_1gpu.json | tr '\n' ' '
