# compute-sanitizer memcheck over the GPU suite (the full-size property tests excluded: 50-100x slowdown)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest -q -m gpu tests \
  -k "not full_size and not fullsize and not 30_layer and not device_sampling_full_size" \
  > gpurun_out/sanitize_memcheck_suite.log 2>&1; echo "exit $?" >> gpurun_out/sanitize_memcheck_suite.log
grep -h "ERROR SUMMARY\|passed\|failed\|^exit" gpurun_out/sanitize_memcheck_suite.log | tail -5
