mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3x3_mf -s 1 -c 1 -o gpurun_out/r2_conv96_mf python tools/gpu_check_kernels.py perf_conv96_fullres > gpurun_out/c16_ncu.log 2>&1
ls -la gpurun_out/r2_conv96_mf.ncu-rep; tail -2 gpurun_out/c16_ncu.log
