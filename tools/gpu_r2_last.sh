mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_last.log 2>&1; tail -n 25 gpurun_out/pytest_gpu_last.log | cut -c1-220
