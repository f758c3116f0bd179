# 8-GPU default bench line after the single-push-stream change
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29513 bench.py --gpus 8 --steps 10 > gpurun_out/r2_bench_8gpu_serial.json 2> gpurun_out/c8c_bench8.err
grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"call_ms": [0-9.]*\|"rel_l2_velocity": [0-9.e-]*\|"frac": [0-9.]*' gpurun_out/r2_bench_8gpu_serial.json | tr '\n' ' '; echo
tail -2 gpurun_out/c8c_bench8.err
