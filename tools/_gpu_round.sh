mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_text_encoder.py -q -m gpu > gpurun_out/t5_tests.log 2>&1; echo "t5 tests exit $?" >> gpurun_out/t5_tests.log
timeout 240 python tools/t5_bench.py > gpurun_out/t5_bench.log 2>&1; echo "t5 bench exit $?" >> gpurun_out/t5_bench.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:t5_attention -s 6 -c 1 -f -o gpurun_out/r1_t5_attention_v2 python tools/t5_bench.py 4 > gpurun_out/ncu_t5.log 2>&1
tail -n 6 gpurun_out/t5_tests.log; tail -n 3 gpurun_out/t5_bench.log
