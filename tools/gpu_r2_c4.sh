mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_compress.py tests/test_gpu_reference_batch_scenarios.py -x -q 2>&1 | tail -n 4
(N=262144 timeout 600 python tools/gpu_c4_dict.py) 2>&1 | grep -v "^compress phases" | tail -n 8
