mkdir -p gpurun_out
timeout 200 python tools/gpu_check_kernels.py perf_gemm_shard > gpurun_out/c1g_gemm.log 2>&1
cp gpurun_out/kernel_check.json gpurun_out/r2_gemm_shard_perf.json
cat gpurun_out/c1g_gemm.log | cut -c1-3000
