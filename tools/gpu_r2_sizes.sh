mkdir -p gpurun_out
for SZ in 16384 32768 65536; do
  N=$((1073741824 / SZ / 2))
  (N=$N SIZE=$SZ MIX=1 timeout 300 python tools/gpu_compress_quick.py) 2>&1 | tail -n 1 | sed "s/^/v2 $SZ: /"
  (ZB200_ENCODER_V1=1 N=$N SIZE=$SZ MIX=1 timeout 300 python tools/gpu_compress_quick.py) 2>&1 | tail -n 1 | sed "s/^/v1 $SZ: /"
done
(ZB200_ENCODER_V1=1 N=4096 SIZE=131072 MIX=1 timeout 300 python tools/gpu_compress_quick.py) 2>&1 | tail -n 1 | sed "s/^/v1 131072: /"
