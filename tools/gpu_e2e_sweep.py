"""End-to-end (pinned host in -> pinned host out) timing of the two batch calls over pipeline depth and
sub-batch size; prints best-of-N wall time per configuration and a per-job timeline for the default one.
Run on the GPU box:  python tools/gpu_e2e_sweep.py [decode|compress|both]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, corpus
from oracle import RefZstd
import python_zstandard_b200 as zstd
from python_zstandard_b200 import decompressor as D, compressor as K

what = sys.argv[1] if len(sys.argv) > 1 else "both"
ref = RefZstd()


def best(fn, reps=6):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), sorted(ts)[len(ts) // 2]


if what in ("decode", "both"):
    n = 262144
    blob, off, ln = corpus.text_segments(n, 4096)
    cblob, clens = ref.batch(True, blob, off, ln, level=3, threads=os.cpu_count())
    coff = np.concatenate([[0], np.cumsum(clens)[:-1]]).astype(np.uint64)
    segs = np.stack([coff, clens], axis=1).astype(np.uint64)
    pin = zstd.PinnedBuffer(len(cblob)); np.frombuffer(pin, dtype=np.uint8)[:] = cblob
    bws = zstd.BufferWithSegments(pin, segs.tobytes())
    d = zstd.ZstdDecompressor()
    U = n * 4096 / 1e6

    def go():
        r = d.multi_decompress_to_buffer(bws); x = r[n - 1].tobytes(); del r

    for depth, mb in ((3, 24), (4, 24), (5, 24), (6, 24), (4, 16), (5, 16), (6, 12), (4, 32), (8, 12)):
        D.ZstdDecompressor.PIPELINE_DEPTH = depth
        D.ZstdDecompressor.SUB_BATCH_INPUT_BYTES = mb << 20
        go(); go()
        b, m = best(go)
        print("decode depth %d sub %2d MiB: best %.2f ms (%.1f GB/s) median %.2f ms" % (depth, mb, b, U / b, m), flush=True)
    D.ZstdDecompressor.PIPELINE_DEPTH = 4
    D.ZstdDecompressor.SUB_BATCH_INPUT_BYTES = 24 << 20
    orig = D.ZstdDecompressor._launch
    log = []

    def traced(self, ctx, base_ptr, segs_, n_, ssz, flags=0):
        t0 = time.perf_counter(); r = orig(self, ctx, base_ptr, segs_, n_, ssz, flags); t1 = time.perf_counter()
        log.append((t0, t1, n_)); return r
    D.ZstdDecompressor._launch = traced
    go()
    log.clear()
    T0 = time.perf_counter(); r = d.multi_decompress_to_buffer(bws); T1 = time.perf_counter()
    print("decode trace (depth 4, 24 MiB): total %.2f ms" % ((T1 - T0) * 1e3))
    for a, b_, k in sorted(log):
        print("   job %6d frames: start %.2f end %.2f (%.2f ms)" % (k, (a - T0) * 1e3, (b_ - T0) * 1e3, (b_ - a) * 1e3))
    D.ZstdDecompressor._launch = orig
    del r

if what in ("compress", "both"):
    cn = int(os.environ.get("CN", "2048"))
    cin, coff_in, cln_in = corpus.silesia_mix(cn, 131072)
    csegs = np.stack([coff_in, cln_in], axis=1).astype(np.uint64)
    cpin = zstd.PinnedBuffer(len(cin)); np.frombuffer(cpin, dtype=np.uint8)[:] = cin
    cbws = zstd.BufferWithSegments(cpin, csegs.tobytes())
    c = zstd.ZstdCompressor(level=3)
    U = len(cin) / 1e6

    def cgo():
        rr = c.multi_compress_to_buffer(cbws); x = rr[cn - 1].tobytes(); del rr

    for depth, mb in ((1, 1024), (3, 64), (3, 32), (4, 32), (4, 16), (6, 16), (8, 8)):
        K.ZstdCompressor.PIPELINE_DEPTH = depth
        K.ZstdCompressor.SUB_BATCH_INPUT_BYTES = mb << 20
        cgo(); cgo()
        b, m = best(cgo, 4)
        print("compress depth %d sub %3d MiB: best %.2f ms (%.1f GB/s) median %.2f ms" % (depth, mb, b, U / b, m), flush=True)
