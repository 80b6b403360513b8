mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c7_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c7_pytest.log
timeout 600 python tools/bench_models.py rot > gpurun_out/c7_models.log 2>&1; echo "exit $?" >> gpurun_out/c7_models.log
grep -E "^E  |passed|failed|^FAILED" gpurun_out/c7_pytest.log | head -30; tail -5 gpurun_out/c7_models.log
