mkdir -p gpurun_out
(N=${NQ:-8192} SIZE=131072 MIX=1 timeout 300 python tools/gpu_compress_quick.py) > gpurun_out/q_mix.log 2>&1
(N=4096 SIZE=131072 MIX=0 timeout 300 python tools/gpu_compress_quick.py) > gpurun_out/q_text.log 2>&1
export ZB200_LIB=$PWD/python_zstandard_b200/libzb200_timers.so
(N=1184 SIZE=131072 MIX=1 timeout 300 python tools/gpu_phase_encode2.py) > gpurun_out/ph_mix.log 2>&1
(N=1184 SIZE=131072 MIX=0 timeout 300 python tools/gpu_phase_encode2.py) > gpurun_out/ph_text.log 2>&1
tail -n 3 gpurun_out/q_mix.log; tail -n 3 gpurun_out/q_text.log; cat gpurun_out/ph_mix.log gpurun_out/ph_text.log
