# Round-2 first GPU call (2 GPUs):  gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_call1.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/c1_smi.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
# 1. GPU suite (new shard / depth / R4 tests included)
timeout 400 python -m pytest tests -q -m gpu -x -s > gpurun_out/c1_tests.log 2>&1; echo "exit $?" >> gpurun_out/c1_tests.log
# 2. RMSNorm+RoPE v2 under the DiT suites
ICB_RMSROPE_V2=1 timeout 300 python -m pytest tests/test_gpu_dit.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -q -m gpu > gpurun_out/c1_tests_v2.log 2>&1; echo "exit $?" >> gpurun_out/c1_tests_v2.log
# 3. sharded == single GPU through NCCL, both layouts, small (bit-identical) and bench-sized
timeout 200 $TR --master-port 29511 tools/check_cfg_parallel.py > gpurun_out/c1_parity_nccl_small.log 2>&1; echo "exit $?" >> gpurun_out/c1_parity_nccl_small.log
timeout 200 $TR --master-port 29512 tools/check_cfg_parallel.py --full --out gpurun_out/r2_shard_parity_2gpu_nccl.json > gpurun_out/c1_parity_nccl_full.log 2>&1; echo "exit $?" >> gpurun_out/c1_parity_nccl_full.log
# 4. peer-memory push: vs NCCL (timed), then vs single GPU
timeout 150 $TR --master-port 29513 tools/check_p2p.py > gpurun_out/c1_p2p_small.log 2>&1; echo "exit $?" >> gpurun_out/c1_p2p_small.log
timeout 200 $TR --master-port 29514 tools/check_p2p.py --full > gpurun_out/c1_p2p_full.log 2>&1; echo "exit $?" >> gpurun_out/c1_p2p_full.log
ICB_KV_P2P=1 timeout 200 $TR --master-port 29515 tools/check_cfg_parallel.py --full --out gpurun_out/r2_shard_parity_2gpu_p2p.json > gpurun_out/c1_parity_p2p_full.log 2>&1; echo "exit $?" >> gpurun_out/c1_parity_p2p_full.log
# 5. 2-GPU bench, plain temporal shard (the layout with an exchange at N = 2), both exchange paths
ICB_CFG_PARALLEL=0 timeout 200 $TR --master-port 29516 bench.py --gpus 2 --steps 6 --skip-e2e > gpurun_out/c1_bench2_nccl.json 2> gpurun_out/c1_bench2_nccl.err
ICB_CFG_PARALLEL=0 ICB_KV_P2P=1 timeout 200 $TR --master-port 29517 bench.py --gpus 2 --steps 6 --skip-e2e > gpurun_out/c1_bench2_p2p.json 2> gpurun_out/c1_bench2_p2p.err
# 6. 1-GPU: v2 bench, library bars
ICB_RMSROPE_V2=1 timeout 200 python bench.py --skip-e2e --skip-parity > gpurun_out/c1_bench_v2.json 2> gpurun_out/c1_bench_v2.err
timeout 600 python tools/gpu_check_kernels.py perf_ > gpurun_out/c1_kernel_perf.log 2>&1
cp gpurun_out/kernel_check.json gpurun_out/c1_library_bars.json 2>/dev/null
tail -n 5 gpurun_out/c1_tests.log gpurun_out/c1_tests_v2.log
grep -h "SHARD_PARITY\|P2P_CHECK\|^exit" gpurun_out/c1_parity_*.log gpurun_out/c1_p2p_*.log
grep -h -o '"value": [0-9.]*' gpurun_out/c1_bench2_nccl.json gpurun_out/c1_bench2_p2p.json gpurun_out/c1_bench_v2.json
