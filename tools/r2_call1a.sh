# Round-2 1-GPU call:  gpurun --timeout 1200 -- 'bash tools/r2_call1a.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/c1_smi.log 2>&1
timeout 400 python -m pytest tests -q -m gpu -x -s > gpurun_out/c1_tests.log 2>&1; echo "exit $?" >> gpurun_out/c1_tests.log
ICB_RMSROPE_V2=1 timeout 300 python -m pytest tests/test_gpu_dit.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -q -m gpu > gpurun_out/c1_tests_v2.log 2>&1; echo "exit $?" >> gpurun_out/c1_tests_v2.log
ICB_RMSROPE_V2=1 timeout 200 python bench.py --skip-e2e --skip-parity > gpurun_out/c1_bench_v2.json 2> gpurun_out/c1_bench_v2.err
timeout 600 python tools/gpu_check_kernels.py perf_ > gpurun_out/c1_kernel_perf.log 2>&1
cp gpurun_out/kernel_check.json gpurun_out/c1_library_bars.json 2>/dev/null
for V in "1 0" "1 1" "1 2"; do set -- $V
  ICB_FMHA_EARLY=$1 ICB_FMHA_EMU=$2 timeout 200 python tools/gpu_check_kernels.py perf_fmha_full fmha_2048 fmha_tails fmha_seg2 > gpurun_out/c1_fmha_e$1_m$2.log 2>&1
  cp gpurun_out/kernel_check.json gpurun_out/c1_fmha_e$1_m$2.json 2>/dev/null
done
tail -n 5 gpurun_out/c1_tests.log gpurun_out/c1_tests_v2.log
grep -h "perf_fmha_full" gpurun_out/c1_kernel_perf.log gpurun_out/c1_fmha_e*.log
grep -h -o '"value": [0-9.]*' gpurun_out/c1_bench_v2.json
