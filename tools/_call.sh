mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c25_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c25_pytest.log
grep -E "^E  |passed|failed|^FAILED" gpurun_out/c25_pytest.log | head -30
