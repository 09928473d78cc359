mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_dev.json 2> gpurun_out/r2_bench_dev.log; echo "bench rc $?"; tail -n 5 gpurun_out/r2_bench_dev.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_dev.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "device_api", "large_frame"):
    print(k, json.dumps(d.get(k))[:600])
print("compress", json.dumps({k: d["compress"].get(k) for k in ("value", "e2e", "size_vs_reference_level3")} if "compress" in d else None)[:500])
PY
