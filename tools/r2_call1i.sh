mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dit.py -x -q -m gpu -k "gemm or token_shard or block or forward" > gpurun_out/c1i_tests.log 2>&1; tail -3 gpurun_out/c1i_tests.log
timeout 300 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_active.max -k regex:gemm --csv --log-file gpurun_out/r2_gemm_shard_launches_after.csv python tools/gemm_shard_probe.py 4 > gpurun_out/c1i_ncu.log 2>&1
tail -2 gpurun_out/c1i_ncu.log
