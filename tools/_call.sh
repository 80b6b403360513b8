mkdir -p gpurun_out
timeout 600 python tools/bench_small.py > gpurun_out/c12_small.log 2>&1; grep sym_eig gpurun_out/c12_small.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c12_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c12_pytest.log
grep -E "^E  |passed|failed|^FAILED" gpurun_out/c12_pytest.log | head
