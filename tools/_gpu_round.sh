mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_knn.py -q -m gpu > gpurun_out/knn_tests.log 2>&1; echo "knn tests exit $?" >> gpurun_out/knn_tests.log
timeout 120 python tools/knn_bench.py > gpurun_out/knn_bench.log 2>&1; echo "knn bench exit $?" >> gpurun_out/knn_bench.log
tail -25 gpurun_out/knn_tests.log; tail -3 gpurun_out/knn_bench.log
