mkdir -p gpurun_out
export ZB200_LIB=$PWD/python_zstandard_b200/libzb200_timers.so
(N=1184 SIZE=131072 MIX=1 timeout 300 python tools/gpu_phase_encode2.py) > gpurun_out/ph_mix.log 2>&1
(N=1184 SIZE=131072 MIX=0 timeout 300 python tools/gpu_phase_encode2.py) > gpurun_out/ph_text.log 2>&1
cat gpurun_out/ph_mix.log gpurun_out/ph_text.log
