# Round-2 8-GPU call: parity vs single GPU, the headline scaling point in both layouts, configs[3] (14B) and configs[2] (stage 2)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/c8_smi.log 2>&1
timeout 250 $TR --master-port 29512 tools/check_cfg_parallel.py --full --out gpurun_out/r2_shard_parity_8gpu.json > gpurun_out/c8_parity_full.log 2>&1; echo "exit $?" >> gpurun_out/c8_parity_full.log
timeout 400 $TR --master-port 29513 bench.py --gpus 8 --steps 10 --skip-e2e-warmup > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/c8_bench8.err
timeout 250 $TR --master-port 29514 bench.py --gpus 8 --steps 10 --cfg-parallel 0 --skip-e2e --skip-parity > gpurun_out/r2_bench_8gpu_temporal_x8.json 2> gpurun_out/c8_bench8_t8.err
ICB_KV_P2P=0 timeout 250 $TR --master-port 29515 bench.py --gpus 8 --steps 10 --skip-e2e --skip-parity > gpurun_out/r2_bench_8gpu_nccl.json 2> gpurun_out/c8_bench8_nccl.err
timeout 500 $TR --master-port 29516 bench.py --gpus 8 --model 14b --steps 5 --skip-e2e-warmup > gpurun_out/r2_bench_14b_8gpu.json 2> gpurun_out/c8_bench14b.err
timeout 500 python tools/stage2_run.py --world 8 --size 512 --out gpurun_out/r2_stage2_8gpu.json > gpurun_out/c8_stage2.log 2>&1; echo "exit $?" >> gpurun_out/c8_stage2.log
grep -h "SHARD_PARITY\|^exit" gpurun_out/c8_parity_full.log | cut -c1-900
for f in gpurun_out/r2_bench_8gpu.json gpurun_out/r2_bench_8gpu_temporal_x8.json gpurun_out/r2_bench_8gpu_nccl.json gpurun_out/r2_bench_14b_8gpu.json; do echo $f; grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"call_ms": [0-9.]*\|"rel_l2_velocity": [0-9.e-]*' $f | tr '\n' ' '; echo; done
grep -h "STAGE2\|^exit" gpurun_out/c8_stage2.log | cut -c1-900
tail -2 gpurun_out/c8_bench14b.err
