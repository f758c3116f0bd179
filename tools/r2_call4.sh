# Round-2 4-GPU call: parity of both layouts vs the single GPU (default exchange = peer-memory push), one bench line,
# and the single-process stage-2 run (RankPool on hardware)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29511 tools/check_cfg_parallel.py --out gpurun_out/r2_shard_parity_4gpu_small.json > gpurun_out/c4_parity_small.log 2>&1; echo "exit $?" >> gpurun_out/c4_parity_small.log
timeout 250 $TR --master-port 29512 tools/check_cfg_parallel.py --full --out gpurun_out/r2_shard_parity_4gpu.json > gpurun_out/c4_parity_full.log 2>&1; echo "exit $?" >> gpurun_out/c4_parity_full.log
timeout 250 $TR --master-port 29513 bench.py --gpus 4 --steps 6 --skip-e2e > gpurun_out/c4_bench4.json 2> gpurun_out/c4_bench4.err
ICB_KV_P2P=0 timeout 250 $TR --master-port 29514 bench.py --gpus 4 --steps 6 --skip-e2e --skip-parity > gpurun_out/c4_bench4_nccl.json 2> gpurun_out/c4_bench4_nccl.err
timeout 400 python tools/stage2_run.py --world 4 --size 256 --out gpurun_out/r2_stage2_4gpu.json > gpurun_out/c4_stage2.log 2>&1; echo "exit $?" >> gpurun_out/c4_stage2.log
grep -h "SHARD_PARITY\|^exit" gpurun_out/c4_parity_*.log
grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"kv_exchange": "[a-z -]*"\|"parity": {[^}]*}[^}]*}' gpurun_out/c4_bench4.json gpurun_out/c4_bench4_nccl.json
grep -h "STAGE2\|^exit\|Error\|error" gpurun_out/c4_stage2.log | tail -5
tail -3 gpurun_out/c4_bench4.err
