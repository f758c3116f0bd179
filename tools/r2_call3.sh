# Round-2 1-GPU call 3: FMHA variants after the max-tree fix, what-if bounds, new tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dit.py tests/test_gpu_raster.py -q -m gpu -s > gpurun_out/c3_tests.log 2>&1; echo "exit $?" >> gpurun_out/c3_tests.log
for V in "0 0" "0 1" "0 2" "1 0" "1 1" "1 2"; do set -- $V
  ICB_FMHA_EARLY=$1 ICB_FMHA_EMU=$2 timeout 200 python tools/gpu_check_kernels.py perf_fmha_full fmha_2048 fmha_tails fmha_seg2 > gpurun_out/c3_fmha_e$1_m$2.log 2>&1
done
for W in 1 2 3 4 5; do
  ICB_FMHA_WHATIF=$W timeout 100 python tools/gpu_check_kernels.py perf_fmha_full > gpurun_out/c3_whatif_$W.log 2>&1
done
grep -h "passed\|failed\|^exit\|coordinate buffer" gpurun_out/c3_tests.log | tail
for f in gpurun_out/c3_fmha_e*.log; do echo $f; grep -h -o '"rel_l2": [0-9.e-]*\|"tflops": [0-9.]*' $f | tr '\n' ' '; echo; done
for f in gpurun_out/c3_whatif_*.log; do echo $f; grep -h -o '"tflops": [0-9.]*' $f | head -1; done
