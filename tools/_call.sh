mkdir -p gpurun_out
timeout 300 python tools/debug_tsc.py > gpurun_out/c15_tsc.log 2>&1; cat gpurun_out/c15_tsc.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c16_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c16_pytest.log
grep -E "^E  |passed|failed|^FAILED" gpurun_out/c16_pytest.log | head -30
