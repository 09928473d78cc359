mkdir -p gpurun_out
(MB=512 timeout 600 python tools/gpu_c5_frame.py) 2>&1 | tail -n 8
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
