mkdir -p gpurun_out
timeout 120 ./build/mma_probe > gpurun_out/c24_probe.log 2>&1; cat gpurun_out/c24_probe.log
