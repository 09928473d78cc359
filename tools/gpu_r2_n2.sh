mkdir -p gpurun_out
ZB_BENCH_COMPRESS_SEGMENTS=8192 ZB_BENCH_DICT_RECORDS=131072 ZB_BENCH_FRAME_MB=64 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc $?"; tail -n 3 gpurun_out/bench_n2.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
for k in ("value", "n_gpus", "e2e", "device_api", "sharded", "large_frame"):
    print(k, json.dumps(d.get(k))[:400])
print("compress", d["compress"]["value"], d["compress"]["e2e"]["value"])
PY
ZB_BENCH_COMPRESS_SEGMENTS=8192 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 2>/dev/null | cut -c1-300
