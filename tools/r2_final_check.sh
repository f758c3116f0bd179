# last call of the round: GPU suite + smoke on the final library
mkdir -p gpurun_out
timeout 100 python -m pytest tests -q -m gpu -x > gpurun_out/f_tests.log 2>&1; echo "exit $?" >> gpurun_out/f_tests.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "exit $?" >> gpurun_out/f_smoke.log
tail -3 gpurun_out/f_tests.log; tail -2 gpurun_out/f_smoke.log
