# Round-2 1-GPU call 15: merged-slice multi-frame conv; encode launch list
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_vae.py tests/test_gpu_pipeline.py -q -m gpu > gpurun_out/c15_tests.log 2>&1; echo "exit $?" >> gpurun_out/c15_tests.log
timeout 100 python tools/gpu_check_kernels.py perf_conv96_fullres > gpurun_out/c15_conv.log 2>&1
timeout 200 python tools/vae_bench.py > gpurun_out/c15_vae.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c15_vae_encode_launches.csv python tools/vae_bench.py --encode-only --once > gpurun_out/c15_vae_enc_ncu.log 2>&1
grep -h "passed\|failed\|^exit" gpurun_out/c15_tests.log | tail -3
tail -qn1 gpurun_out/c15_conv.log gpurun_out/c15_vae.log | cut -c1-260
