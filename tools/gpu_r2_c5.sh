mkdir -p gpurun_out
(MB=512 timeout 600 python tools/gpu_c5_frame.py) 2>&1 | tail -n 6
(N=2048 timeout 300 python tools/gpu_c3_decode.py) 2>&1 | grep -v phase | tail -n 3
(N=2048 ZB200_BLOCK_PATH=0 timeout 300 python tools/gpu_c3_decode.py) 2>&1 | grep -v phase | tail -n 1
(N=200 timeout 300 python tools/gpu_c3_decode.py) 2>&1 | grep -v phase | tail -n 2
(N=200 ZB200_BLOCK_PATH=0 timeout 300 python tools/gpu_c3_decode.py) 2>&1 | grep -v phase | tail -n 1
