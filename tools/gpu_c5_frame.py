"""BASELINE config 5 in small: ONE frame of many 128 KiB blocks through stream_reader / decompress, block-parallel path vs
the lane-per-frame path (ZB200_BLOCK_PATH=0), the unmodified reference on one core beside it.
  MB=256 python tools/gpu_c5_frame.py"""
import os, sys, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, corpus
from oracle import RefZstd
import python_zstandard_b200 as zstd
from python_zstandard_b200 import _native
mb = int(os.environ.get("MB", "256"))
t = corpus.text_corpus(8 << 20)
data = np.tile(t, (mb << 20) // len(t) + 1)[:mb << 20].tobytes()
ref = RefZstd()
t0 = time.perf_counter(); frame = ref.compress(data, level=3, checksum=os.environ.get("CK", "0") == "1"); tc = time.perf_counter() - t0
t0 = time.perf_counter(); back = ref.decompress(frame, len(data)); td = time.perf_counter() - t0
assert back == data
print("frame: %d MiB -> %.1f MiB, %d blocks; reference (1 thread): compress %.2f GB/s, decompress %.2f GB/s" % (
    mb, len(frame) / 2**20, (len(data) + 131071) // 131072, len(data) / tc / 1e9, len(data) / td / 1e9), flush=True)
d = zstd.ZstdDecompressor(max_window_size=1 << 31)
ctx = _native.Context.get(0)
for it in range(3):
    t0 = time.perf_counter(); out = d.decompress(frame); dt = time.perf_counter() - t0
    assert out == data
    print("decompress(): %.3f s = %.2f GB/s" % (dt, len(data) / dt / 1e9), flush=True)
ctx.profile(True); out = d.decompress(frame); pr = ctx.profile_read(); ctx.profile(False)
print("kernels (ms):", {k: round(v[0], 2) for k, v in pr.items()}, " pointer-doubling rounds:", ctx.L.zb200_last_chase_rounds(ctx.h), flush=True)
t0 = time.perf_counter()
with d.stream_reader(io.BytesIO(frame)) as r:
    n = 0
    while True:
        c = r.read(1 << 20)
        if not c: break
        n += len(c)
dt = time.perf_counter() - t0
assert n == len(data)
print("stream_reader: %.3f s = %.2f GB/s" % (dt, len(data) / dt / 1e9), flush=True)
# one-shot compress() of the same large buffer: one frame of independent 128 KiB blocks
c = zstd.ZstdCompressor(level=3)
for it in range(2):
    t0 = time.perf_counter(); ours = c.compress(data); dt = time.perf_counter() - t0
print("compress(): %.3f s = %.2f GB/s, %.1f MiB (reference level 3: %.1f MiB, %+.2f %%)" % (
    dt, len(data) / dt / 1e9, len(ours) / 2**20, len(frame) / 2**20, 100.0 * (len(ours) / len(frame) - 1)), flush=True)
assert ref.decompress(ours, len(data)) == data
t0 = time.perf_counter(); back = d.decompress(ours); dt = time.perf_counter() - t0
assert back == data
print("decompress() of our own frame: %.3f s = %.2f GB/s" % (dt, len(data) / dt / 1e9), flush=True)
