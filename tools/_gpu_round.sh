mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_text_encoder.py -q -m gpu > gpurun_out/t5_tests.log 2>&1; echo "t5 tests exit $?" >> gpurun_out/t5_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_r1_final_2gpu.json 2> gpurun_out/bench2.err; echo "bench2 exit $?" >> gpurun_out/bench2.err
tail -5 gpurun_out/t5_tests.log; tail -5 gpurun_out/bench2.err; head -c 400 gpurun_out/bench_r1_final_2gpu.json
