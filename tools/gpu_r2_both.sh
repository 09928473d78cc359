mkdir -p gpurun_out
N=296 MIX=0 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:zb_compress_smem --launch-skip 1 --launch-count 1 -o gpurun_out/r2_compress_smem -f python tools/gpu_prof_encode.py > gpurun_out/ncu.log 2>&1
tail -n 3 gpurun_out/ncu.log
bash tools/gpu_r2_phase_only.sh
