set -x
mkdir -p gpurun_out
(N=8192 SIZE=131072 MIX=1 timeout 300 python tools/gpu_compress_quick.py) > gpurun_out/q_mix.log 2>&1
(N=4096 SIZE=131072 MIX=0 timeout 300 python tools/gpu_compress_quick.py) > gpurun_out/q_text.log 2>&1
(ZB200_LIB=$PWD/python_zstandard_b200/libzb200_timers.so N=1184 SIZE=131072 MIX=1 timeout 300 python tools/gpu_phase_encode2.py) > gpurun_out/ph_mix.log 2>&1
(ZB200_LIB=$PWD/python_zstandard_b200/libzb200_timers.so N=1184 SIZE=131072 MIX=0 timeout 300 python tools/gpu_phase_encode2.py) > gpurun_out/ph_text.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.log 2>&1
tail -5 gpurun_out/q_mix.log gpurun_out/q_text.log gpurun_out/ph_mix.log gpurun_out/ph_text.log gpurun_out/pytest_gpu.log
tail -c 3000 gpurun_out/bench.log
