# A/B of the peer-push order on a x4 temporal shard (3 peers per rank, the same exchange as the 8-GPU default layout)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
ICB_P2P_SERIAL=1 timeout 200 $TR --master-port 29531 tools/check_cfg_parallel.py --full --out gpurun_out/r2_shard_parity_4gpu_serial.json > gpurun_out/c4s_parity.log 2>&1; echo "exit $?" >> gpurun_out/c4s_parity.log
for rep in 1 2; do for s in 1 0; do
ICB_P2P_SERIAL=$s timeout 200 $TR --master-port 2954$s bench.py --gpus 4 --cfg-parallel 0 --steps 8 --skip-e2e --skip-parity --skip-raster > gpurun_out/c4s_serial${s}_$rep.json 2> gpurun_out/c4s_serial${s}_$rep.err
echo "serial=$s rep=$rep: $(grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/c4s_serial${s}_$rep.json | head -2 | tr '\n' ' ')"
done; done
grep -h "SHARD_PARITY\|^exit" gpurun_out/c4s_parity.log | cut -c1-600
