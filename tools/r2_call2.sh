# Round-2 1-GPU call 2:  gpurun --timeout 1200 -- 'bash tools/r2_call2.sh'
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu -s > gpurun_out/c2_tests.log 2>&1; echo "exit $?" >> gpurun_out/c2_tests.log
ICB_RMSROPE_V2=1 timeout 300 python -m pytest tests/test_gpu_dit.py tests/test_gpu_pipeline.py -q -m gpu -s > gpurun_out/c2_tests_v2.log 2>&1; echo "exit $?" >> gpurun_out/c2_tests_v2.log
nvcc -gencode arch=compute_100a,code=sm_100a -I infinicube_b200/csrc -o /tmp/emu_test tools/emu_test.cu > gpurun_out/c2_emu.log 2>&1 && /tmp/emu_test >> gpurun_out/c2_emu.log 2>&1
for W in 0 1 2 3 4 5; do
  ICB_FMHA_WHATIF=$W timeout 100 python tools/gpu_check_kernels.py perf_fmha_full > gpurun_out/c2_whatif_$W.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fmha_fwd -s 1 -c 1 -o gpurun_out/r2_fmha_base python tools/gpu_check_kernels.py perf_fmha_full > gpurun_out/c2_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
grep -h "passed\|failed\|^exit\|rel-L2\|teacher" gpurun_out/c2_tests.log gpurun_out/c2_tests_v2.log | tail -20
cat gpurun_out/c2_emu.log
grep -h "RESULT" gpurun_out/c2_whatif_*.log
