mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_device_buffers.py -x -q 2>&1 | tail -n 3
ZB_BENCH_PROFILE_DEVICE_API=1 ZB_BENCH_COMPRESS_SEGMENTS=8192 ZB_BENCH_DICT_RECORDS=131072 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_dev.json 2> gpurun_out/r2_bench_dev.log; echo "bench rc $?"; grep -A 22 "function calls" gpurun_out/r2_bench_dev.log | cut -c1-150
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_dev.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "device_api", "large_frame"):
    print(k, json.dumps(d.get(k))[:700])
PY
