# Round-2 2-GPU call:  gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_call1b.sh'
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29511 tools/check_cfg_parallel.py > gpurun_out/c1_parity_nccl_small.log 2>&1; echo "exit $?" >> gpurun_out/c1_parity_nccl_small.log
timeout 200 $TR --master-port 29512 tools/check_cfg_parallel.py --full --out gpurun_out/r2_shard_parity_2gpu_nccl.json > gpurun_out/c1_parity_nccl_full.log 2>&1; echo "exit $?" >> gpurun_out/c1_parity_nccl_full.log
timeout 150 $TR --master-port 29513 tools/check_p2p.py > gpurun_out/c1_p2p_small.log 2>&1; echo "exit $?" >> gpurun_out/c1_p2p_small.log
timeout 200 $TR --master-port 29514 tools/check_p2p.py --full > gpurun_out/c1_p2p_full.log 2>&1; echo "exit $?" >> gpurun_out/c1_p2p_full.log
ICB_KV_P2P=1 timeout 200 $TR --master-port 29515 tools/check_cfg_parallel.py --full --out gpurun_out/r2_shard_parity_2gpu_p2p.json > gpurun_out/c1_parity_p2p_full.log 2>&1; echo "exit $?" >> gpurun_out/c1_parity_p2p_full.log
ICB_CFG_PARALLEL=0 timeout 200 $TR --master-port 29516 bench.py --gpus 2 --steps 6 --skip-e2e > gpurun_out/c1_bench2_nccl.json 2> gpurun_out/c1_bench2_nccl.err
ICB_CFG_PARALLEL=0 ICB_KV_P2P=1 timeout 200 $TR --master-port 29517 bench.py --gpus 2 --steps 6 --skip-e2e > gpurun_out/c1_bench2_p2p.json 2> gpurun_out/c1_bench2_p2p.err
grep -h "SHARD_PARITY\|P2P_CHECK\|^exit" gpurun_out/c1_parity_*.log gpurun_out/c1_p2p_*.log
grep -h -o '"value": [0-9.]*' gpurun_out/c1_bench2_nccl.json gpurun_out/c1_bench2_p2p.json
tail -3 gpurun_out/c1_bench2_p2p.err
