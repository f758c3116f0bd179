mkdir -p gpurun_out
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:fmha|gemm_bf16|ln_modulate|rmsnorm|add_bcast|gemv|patchify|sinusoid|cast_f32|copy_convert' -c 1900 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --skip-e2e --skip-parity --skip-raster > gpurun_out/c1j_bench_under_ncu.log 2>&1
echo "exit $?"; wc -l gpurun_out/r2_launches.csv
