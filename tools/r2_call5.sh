# Round-2 1-GPU call 5
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu -s > gpurun_out/c5_tests.log 2>&1; echo "exit $?" >> gpurun_out/c5_tests.log
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/mufu_bench tools/mufu_bench.cu > gpurun_out/c5_mufu.log 2>&1 && /tmp/mufu_bench >> gpurun_out/c5_mufu.log 2>&1
timeout 300 python bench.py --model 14b --steps 3 --skip-e2e > gpurun_out/c5_bench_14b_1gpu.json 2> gpurun_out/c5_bench_14b_1gpu.err
ICB_FMHA_EMU=0 timeout 200 python bench.py --skip-e2e --skip-parity --skip-raster --steps 8 > gpurun_out/c5_bench_emu0.json 2> gpurun_out/c5_bench_emu0.err
ICB_FMHA_EMU=2 timeout 200 python bench.py --skip-e2e --skip-parity --skip-raster --steps 8 > gpurun_out/c5_bench_emu2.json 2> gpurun_out/c5_bench_emu2.err
ICB_FMHA_EMU=1 timeout 200 python bench.py --skip-e2e --skip-parity --skip-raster --steps 8 > gpurun_out/c5_bench_emu1.json 2> gpurun_out/c5_bench_emu1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c5_vae_untiled_launches.csv python tools/vae_bench.py --untiled --decode-only --once > gpurun_out/c5_vae_untiled.log 2>&1
timeout 200 python tools/vae_bench.py > gpurun_out/c5_vae_tiled.log 2>&1
timeout 600 python tools/raster_sweep.py > gpurun_out/c5_raster_sweep.log 2>&1
grep -h "passed\|failed\|^exit" gpurun_out/c5_tests.log | tail -4
cat gpurun_out/c5_mufu.log
for f in gpurun_out/c5_bench_14b_1gpu.json gpurun_out/c5_bench_emu0.json gpurun_out/c5_bench_emu1.json gpurun_out/c5_bench_emu2.json; do echo $f; grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"achieved": [0-9.]*' $f | head -3 | tr '\n' ' '; echo; done
tail -2 gpurun_out/c5_bench_14b_1gpu.err
tail -1 gpurun_out/c5_vae_untiled.log; tail -1 gpurun_out/c5_vae_tiled.log; tail -4 gpurun_out/c5_raster_sweep.log | cut -c1-400
