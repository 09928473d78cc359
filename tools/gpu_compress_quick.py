"""Quick compress timing on the GPU box: N x SIZE Silesia-mix (or text) segments, device-resident kernel time,
round trip through the reference decoder, size vs the reference's level 3, optional end-to-end timing.
  N=2048 SIZE=131072 MIX=1 E2E=1 python tools/gpu_compress_quick.py"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, corpus, torch
from oracle import RefZstd
import python_zstandard_b200 as zstd
from python_zstandard_b200 import _native

n = int(os.environ.get("N", "2048")); size = int(os.environ.get("SIZE", "131072"))
mix = os.environ.get("MIX", "1") == "1"
blob, off, ln = (corpus.silesia_mix if mix else corpus.text_segments)(n, size)
segs = np.stack([off, ln], axis=1).astype(np.uint64)
ref = RefZstd()
L = _native.lib()
ctx = _native.Context.get(0)
d_in = torch.empty(len(blob) + 256, dtype=torch.uint8, device="cuda"); d_in[:len(blob)].copy_(torch.from_numpy(blob))
d_segs = torch.from_numpy(segs.view(np.int64).copy()).cuda()
p = zstd.compressor.CParams(3, 0, 1, 0)


def step():
    r = C.c_void_p()
    ctx.check(L.zb200_compress_batch(ctx.h, d_in.data_ptr(), d_segs.data_ptr(), n, C.byref(p), None,
                                     _native.SRC_DEVICE | _native.DST_DEVICE, C.byref(r)), "compress")
    return r


r0 = step()
csz = int(L.zb200_result_size(r0))
comp = np.empty(csz, dtype=np.uint8)
ctx.check(L.zb200_memcpy_d2h(ctx.h, comp.ctypes.data, L.zb200_result_data(r0), csz), "d2h")
cs = np.ctypeslib.as_array(C.cast(L.zb200_result_segments(r0), C.POINTER(C.c_uint64)), shape=(n, 2)).copy()
L.zb200_result_free(r0)
back, _ = ref.batch(False, comp, np.ascontiguousarray(cs[:, 0]), np.ascontiguousarray(cs[:, 1]), threads=os.cpu_count())
ok = np.array_equal(back, blob)
refc, reflens = ref.batch(True, blob, off, ln, level=3, threads=os.cpu_count())
L.zb200_result_free(step())
ctx.profile(True)
reps = 5
for _ in range(reps):
    L.zb200_result_free(step())
pr = ctx.profile_read()
ctx.profile(False)
tot = sum(v[0] for v in pr.values()) / reps
print("roundtrip %s  size %d vs level-3 %d (%+.2f%%)  kernels %s  total %.3f ms -> %.2f GB/s" % (
    "OK" if ok else "FAILED", csz, int(reflens.sum()), 100.0 * (csz / float(reflens.sum()) - 1),
    {k: round(v[0] / reps, 3) for k, v in pr.items()}, tot, len(blob) / tot / 1e6), flush=True)

if os.environ.get("E2E", "0") == "1":
    from python_zstandard_b200 import compressor as K
    pin = zstd.PinnedBuffer(len(blob)); np.frombuffer(pin, dtype=np.uint8)[:] = blob
    bws = zstd.BufferWithSegments(pin, segs.tobytes())
    c = zstd.ZstdCompressor(level=3)
    for depth, mb in ((1, 1 << 20), (2, 128), (4, 32)):
        K.ZstdCompressor.PIPELINE_DEPTH = depth; K.ZstdCompressor.SUB_BATCH_INPUT_BYTES = mb << 20
        ts = []
        for it in range(5):
            t0 = time.perf_counter(); rr = c.multi_compress_to_buffer(bws); x = rr[n - 1].tobytes(); ts.append((time.perf_counter() - t0) * 1e3); del rr
        print("e2e depth %d sub %d MiB: best %.2f ms (%.2f GB/s)" % (depth, mb, min(ts[1:]), len(blob) / min(ts[1:]) / 1e6), flush=True)
