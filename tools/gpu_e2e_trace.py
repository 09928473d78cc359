import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, corpus
from oracle import RefZstd
import python_zstandard_b200 as zstd
from python_zstandard_b200 import decompressor as D
n = 262144
ref = RefZstd()
blob, off, ln = corpus.text_segments(n, 4096)
cblob, clens = ref.batch(True, blob, off, ln, level=3, threads=os.cpu_count())
coff = np.concatenate([[0], np.cumsum(clens)[:-1]]).astype(np.uint64)
segs = np.stack([coff, clens], axis=1).astype(np.uint64)
pin = zstd.PinnedBuffer(len(cblob)); np.frombuffer(pin, dtype=np.uint8)[:] = cblob
bws = zstd.BufferWithSegments(pin, segs.tobytes())
d = zstd.ZstdDecompressor()
orig = D.ZstdDecompressor._launch
log = []
def traced(self, ctx, base_ptr, segs_, n_, ssz, flags=0):
    t0 = time.perf_counter(); r = orig(self, ctx, base_ptr, segs_, n_, ssz, flags); t1 = time.perf_counter()
    log.append((t0, t1, n_)); return r
D.ZstdDecompressor._launch = traced
for it in range(4):
    log.clear()
    T0 = time.perf_counter(); r = d.multi_decompress_to_buffer(bws); T1 = time.perf_counter()
    x = r[n - 1].tobytes(); T2 = time.perf_counter()
    print("iter %d total %.1f ms (+tobytes %.2f)" % (it, (T1 - T0) * 1e3, (T2 - T1) * 1e3))
    for a, b, k in sorted(log): print("   job %6d frames: start %.2f end %.2f (%.2f ms)" % (k, (a - T0) * 1e3, (b - T0) * 1e3, (b - a) * 1e3))
    del r
