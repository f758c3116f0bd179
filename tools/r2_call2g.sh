mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29513 bench.py --gpus 2 --steps 10 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/c2g_bench2.err
timeout 200 $TR --master-port 29514 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/c2g_ref.json 2> gpurun_out/c2g_ref.err
grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"call_ms": [0-9.]*\|"rel_l2_velocity": [0-9.e-]*\|"parallelism": "[^"]*"' gpurun_out/r2_bench_2gpu.json | tr '\n' ' '; echo
cut -c1-200 gpurun_out/c2g_ref.json; tail -2 gpurun_out/c2g_bench2.err
