# 1-GPU validation of the committed state: GPU suite, smoke, both bench arms
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/v_tests.log 2>&1; echo "exit $?" >> gpurun_out/v_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.log 2>&1; echo "exit $?" >> gpurun_out/v_smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/v_bench_ref.err
timeout 600 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/v_bench.err
tail -3 gpurun_out/v_tests.log; tail -2 gpurun_out/v_smoke.log
grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"call_ms": [0-9.]*\|"frac": [0-9.]*\|"render_ms_93cams": [0-9.]*' gpurun_out/r2_bench_1gpu.json | tr '\n' ' '
