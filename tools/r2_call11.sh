# Round-2 1-GPU call 11: FMHA emulation share 2/3/4 eighths, P hand-over after 64 vs 96 keys
mkdir -p gpurun_out
for M in 2 3 4; do
  ICB_FMHA_EMU=$M timeout 200 python tools/gpu_check_kernels.py perf_fmha_full fmha_2048 fmha_tails fmha_seg2 > gpurun_out/c11_fmha_h3_m$M.log 2>&1
done
for M in 2 3; do
  ICB_LIB_PATH=$PWD/infinicube_b200/build/libinfinicube_b200_head2.so ICB_FMHA_EMU=$M timeout 200 python tools/gpu_check_kernels.py perf_fmha_full fmha_2048 fmha_tails fmha_seg2 > gpurun_out/c11_fmha_h2_m$M.log 2>&1
done
timeout 300 python -m pytest tests/test_gpu_dit.py -q -m gpu > gpurun_out/c11_tests.log 2>&1; echo "exit $?" >> gpurun_out/c11_tests.log
for f in gpurun_out/c11_fmha_h*.log; do echo $f; grep -h -o '"rel_l2": [0-9.e-]*\|"tflops": [0-9.]*' $f | tr '\n' ' '; echo; done
tail -2 gpurun_out/c11_tests.log
