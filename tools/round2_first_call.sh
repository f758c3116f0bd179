# First GPU call of round 2 (2 GPUs, about 3 minutes of box time = 6 GPU-minutes):
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/round2_first_call.sh'
# 1. opt-in RMSNorm+RoPE v2 under the full GPU suite and the 1-GPU bench  -> make it the default if green
# 2. peer-memory K / V^T exchange: bit-identity vs the NCCL path (small, then bench-sized), then the 2-GPU bench with
#    the plain temporal shard (ICB_CFG_PARALLEL=0, otherwise 2 GPUs have no exchange at all) on both paths
mkdir -p gpurun_out
ICB_RMSROPE_V2=1 timeout 300 python -m pytest tests -q -m gpu > gpurun_out/r2_v2_tests.log 2>&1; echo "exit $?" >> gpurun_out/r2_v2_tests.log
ICB_RMSROPE_V2=1 timeout 200 python bench.py --skip-e2e > gpurun_out/r2_bench_v2.json 2> gpurun_out/r2_bench_v2.err
timeout 200 python bench.py --skip-e2e > gpurun_out/r2_bench_v1.json 2> gpurun_out/r2_bench_v1.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29521 tools/check_p2p.py > gpurun_out/r2_p2p_small.log 2>&1; echo "exit $?" >> gpurun_out/r2_p2p_small.log
timeout 200 $TR --master-port 29522 tools/check_p2p.py --full > gpurun_out/r2_p2p_full.log 2>&1; echo "exit $?" >> gpurun_out/r2_p2p_full.log
ICB_CFG_PARALLEL=0 timeout 200 $TR --master-port 29523 bench.py --gpus 2 --steps 6 --skip-e2e > gpurun_out/r2_bench2_nccl.json 2> gpurun_out/r2_bench2_nccl.err
ICB_CFG_PARALLEL=0 ICB_KV_P2P=1 timeout 200 $TR --master-port 29524 bench.py --gpus 2 --steps 6 --skip-e2e > gpurun_out/r2_bench2_p2p.json 2> gpurun_out/r2_bench2_p2p.err
tail -n 4 gpurun_out/r2_v2_tests.log gpurun_out/r2_p2p_small.log gpurun_out/r2_p2p_full.log
grep -h -o '"value": [0-9.]*' gpurun_out/r2_bench_v1.json gpurun_out/r2_bench_v2.json gpurun_out/r2_bench2_nccl.json gpurun_out/r2_bench2_p2p.json
