# Round-2 final 8-GPU call: parity, the default bench line with a warm e2e call, stage 2 from one process
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 250 $TR --master-port 29512 tools/check_cfg_parallel.py --full --out gpurun_out/r2_shard_parity_8gpu.json > gpurun_out/c8b_parity_full.log 2>&1; echo "exit $?" >> gpurun_out/c8b_parity_full.log
timeout 500 $TR --master-port 29513 bench.py --gpus 8 --steps 10 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/c8b_bench8.err
timeout 500 python tools/stage2_run.py --world 8 --size 512 --out gpurun_out/r2_stage2_8gpu.json > gpurun_out/c8b_stage2.log 2>&1; echo "exit $?" >> gpurun_out/c8b_stage2.log
grep -h "SHARD_PARITY\|^exit" gpurun_out/c8b_parity_full.log | cut -c1-700
grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"call_ms": [0-9.]*\|"rel_l2_velocity": [0-9.e-]*\|"frac": [0-9.]*' gpurun_out/r2_bench_8gpu.json | tr '\n' ' '; echo
grep -h "STAGE2\|^exit" gpurun_out/c8b_stage2.log | cut -c1-900
tail -2 gpurun_out/c8b_bench8.err
