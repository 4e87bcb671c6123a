# compute-sanitizer over the parity suite (small cases; the 10 M-atom tests are deselected)
set -x
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --target-processes all python -m pytest tests/test_parity_gpu.py tests/test_tohost_gpu.py tests/test_api_gpu.py -x -q -m gpu -k "not fuzz and not million and not 1m" > gpurun_out/san_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/san_memcheck.log
tail -5 gpurun_out/san_memcheck.log
grep -c "Invalid\|out of bounds" gpurun_out/san_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "fcc or high_density or edge" > gpurun_out/san_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/san_racecheck.log
tail -4 gpurun_out/san_racecheck.log
