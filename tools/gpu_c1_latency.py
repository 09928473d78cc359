"""BASELINE config 1 (the reference's own CPU-runnable case): ONE 64 KiB level-3 frame through decompress() / compress(),
call latency on the GPU path against the reference on one core.  A single small frame is latency, not throughput:
one lane's serial chain plus the launches and two PCIe hops.
   python tools/gpu_c1_latency.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, corpus
from oracle import RefZstd
import python_zstandard_b200 as zstd
ref = RefZstd()
data = corpus.text_corpus(1 << 20)[4096:4096 + 65536].tobytes()
frame = ref.compress(data, level=3)
d = zstd.ZstdDecompressor(); c = zstd.ZstdCompressor(level=3)
assert d.decompress(frame) == data and ref.decompress(c.compress(data), 65536) == data


def med(fn, n=200):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    ts.sort(); return ts[len(ts) // 2] * 1e6, ts[len(ts) // 10] * 1e6


for name, gpu, cpu in (("decompress 64 KiB", lambda: d.decompress(frame), lambda: ref.decompress(frame, 65536)),
                       ("compress 64 KiB", lambda: c.compress(data), lambda: ref.compress(data, level=3))):
    g = med(gpu); h = med(cpu)
    print("%s: GPU path median %.0f us (p10 %.0f), reference on one core median %.0f us (p10 %.0f)" % (name, g[0], g[1], h[0], h[1]), flush=True)
