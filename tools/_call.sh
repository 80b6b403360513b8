mkdir -p gpurun_out
timeout 600 python tools/bench_kernels.py 8760 1038240 60 > gpurun_out/c22_kernels.log 2>&1
timeout 600 python tools/bench_kernels.py 8760 259200 128 >> gpurun_out/c22_kernels.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c22_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c22_pytest.log
timeout 900 python bench.py --no-cpu --no-e2e > gpurun_out/c22_bench.log 2>&1; echo "exit $?" >> gpurun_out/c22_bench.log
cat gpurun_out/c22_kernels.log; grep -E "^E  |passed|failed|^FAILED" gpurun_out/c22_pytest.log | head; tail -2 gpurun_out/c22_bench.log | cut -c1-1500
