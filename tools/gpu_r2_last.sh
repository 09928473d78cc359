mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
(MB=256 CK=1 timeout 600 python tools/gpu_c5_frame.py) 2>&1 | tail -n 7
