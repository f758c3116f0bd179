# One gpurun call that re-verifies a round on a fresh B200 box (about 2.5 GPU-minutes):
#   gpurun --timeout 900 -- 'bash tools/_gpu_round.sh'
# smoke() -> full GPU test suite (as the driver runs it) -> default bench -> ncu launch list of the same command.
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests.log
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1300 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --skip-e2e > gpurun_out/ncu_list.log 2>&1; echo "ncu exit $?" >> gpurun_out/ncu_list.log
tail -n 3 gpurun_out/smoke.log gpurun_out/gpu_tests.log gpurun_out/bench.err gpurun_out/ncu_list.log
