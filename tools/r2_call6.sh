# Round-2 1-GPU call 6: VAE investigations
mkdir -p gpurun_out
ICB_RMSNORM_CL_REG=1 timeout 300 python -m pytest tests/test_gpu_vae.py tests/test_gpu_pipeline.py -q -m gpu -s > gpurun_out/c6_tests_reg.log 2>&1; echo "exit $?" >> gpurun_out/c6_tests_reg.log
ICB_RMSNORM_CL_REG=0 timeout 200 python tools/vae_bench.py > gpurun_out/c6_vae_tiled_reg0.log 2>&1
ICB_RMSNORM_CL_REG=1 timeout 200 python tools/vae_bench.py > gpurun_out/c6_vae_tiled_reg1.log 2>&1
timeout 200 python tools/vae_bench.py --untiled --decode-only > gpurun_out/c6_vae_untiled_warm.log 2>&1
PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True timeout 200 python tools/vae_bench.py --untiled --decode-only > gpurun_out/c6_vae_untiled_warm_expseg.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c6_vae_tiled_launches.csv python tools/vae_bench.py --decode-only --once > gpurun_out/c6_vae_tiled_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm --launch-skip 34 -c 1 -o gpurun_out/r2_conv96 python tools/vae_bench.py --untiled --decode-only --once > gpurun_out/c6_ncu_conv.log 2>&1
grep -h "passed\|failed\|^exit" gpurun_out/c6_tests_reg.log | tail -3
tail -qn1 gpurun_out/c6_vae_tiled_reg0.log gpurun_out/c6_vae_tiled_reg1.log gpurun_out/c6_vae_untiled_warm.log gpurun_out/c6_vae_untiled_warm_expseg.log | cut -c1-200
ls -la gpurun_out/r2_conv96.ncu-rep; tail -2 gpurun_out/c6_ncu_conv.log
