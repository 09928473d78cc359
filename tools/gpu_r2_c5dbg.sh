mkdir -p gpurun_out
(MB=384 ZB200_LIB=python_zstandard_b200/libzb200_dbg.so timeout 600 python tools/gpu_c5_frame.py) > gpurun_out/c5dbg.log 2>&1
grep -c blk gpurun_out/c5dbg.log; tail -n 3 gpurun_out/c5dbg.log
