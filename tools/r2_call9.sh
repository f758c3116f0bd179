# Round-2 1-GPU call 9: max-free (lazy) softmax: correctness under both settings, isolated and in-step timing
mkdir -p gpurun_out
ICB_FMHA_LAZY=0 timeout 300 python -m pytest tests/test_gpu_dit.py -q -m gpu > gpurun_out/c9_tests_lazy0.log 2>&1; echo "exit $?" >> gpurun_out/c9_tests_lazy0.log
ICB_FMHA_LAZY=1 timeout 300 python -m pytest tests/test_gpu_dit.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -q -m gpu > gpurun_out/c9_tests_lazy1.log 2>&1; echo "exit $?" >> gpurun_out/c9_tests_lazy1.log
for V in "0 2" "1 0" "1 1" "1 2"; do set -- $V
  ICB_FMHA_LAZY=$1 ICB_FMHA_EMU=$2 timeout 200 python tools/gpu_check_kernels.py perf_fmha_full fmha_2048 fmha_tails fmha_seg2 > gpurun_out/c9_fmha_l$1_m$2.log 2>&1
done
ICB_FMHA_LAZY=1 timeout 200 python bench.py --skip-e2e --skip-parity --skip-raster --steps 8 > gpurun_out/c9_bench_lazy1.json 2> gpurun_out/c9_bench_lazy1.err
ICB_FMHA_LAZY=0 timeout 200 python bench.py --skip-e2e --skip-parity --skip-raster --steps 8 > gpurun_out/c9_bench_lazy0.json 2> gpurun_out/c9_bench_lazy0.err
grep -h "passed\|failed\|^exit\|Error" gpurun_out/c9_tests_lazy0.log gpurun_out/c9_tests_lazy1.log | tail -8
for f in gpurun_out/c9_fmha_l*.log; do echo $f; grep -h -o '"rel_l2": [0-9.e-]*\|"tflops": [0-9.]*' $f | tr '\n' ' '; echo; done
for f in gpurun_out/c9_bench_lazy0.json gpurun_out/c9_bench_lazy1.json; do grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"achieved": [0-9.]*' $f | head -3 | tr '\n' ' '; echo; done
