# round-2 first GPU pass: new graphed path tests, reference suite, full gpu suite, bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graphed.py tests/test_reference_suite.py -x -q -m gpu > gpurun_out/r2a_new_tests.log 2>&1; echo "new tests rc=$?"
tail -30 gpurun_out/r2a_new_tests.log
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_graphed.py --deselect tests/test_reference_suite.py > gpurun_out/r2a_old_tests.log 2>&1; echo "old tests rc=$?"
tail -5 gpurun_out/r2a_old_tests.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2a_bench.err
cat gpurun_out/r2a_bench.json
