mkdir -p gpurun_out
timeout 120 python tools/gemm_shard_probe.py 12 > gpurun_out/c1h_plain.log 2>&1
timeout 300 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_active.min,sm__cycles_elapsed.max -k regex:gemm --csv --log-file gpurun_out/r2_gemm_shard_launches.csv python tools/gemm_shard_probe.py 4 > gpurun_out/c1h_ncu.log 2>&1
grep SHAPE gpurun_out/c1h_plain.log
wc -l gpurun_out/r2_gemm_shard_launches.csv
