"""Config 4 (dictionary path) timing on the GPU box: N ~1 KiB JSON-like records with a trained dictionary,
compress + decompress through the public API (host buffers) and the CPU reference beside it.
  N=262144 python tools/gpu_c4_dict.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, corpus
from oracle import RefZstd
import python_zstandard_b200 as zstd
from python_zstandard_b200 import _native

n = int(os.environ.get("N", "262144"))
ref = RefZstd()
recs = corpus.json_records(n + 2000)
dct = ref.train_dictionary(112640, recs[:2000])
recs = recs[2000:]
ln = np.array([len(r) for r in recs], dtype=np.uint64)
off = np.concatenate([[0], np.cumsum(ln)[:-1]]).astype(np.uint64)
blob = np.frombuffer(b"".join(recs), dtype=np.uint8)
U = int(ln.sum())
cores = os.cpu_count()
print("records %d, %.1f MB, mean %.0f B, dict %d B" % (n, U / 1e6, U / n, len(dct)), flush=True)

# CPU reference (oracle/_ref through the batch orchestration), best of 3
rc, rl = ref.batch(True, blob, off, ln, level=3, threads=cores, dict_data=dct)
ro = np.concatenate([[0], np.cumsum(rl)[:-1]]).astype(np.uint64)
tc = td = 1e9
for _ in range(3):
    t0 = time.perf_counter(); ref.batch(True, blob, off, ln, level=3, threads=cores, dict_data=dct, gather=False); tc = min(tc, time.perf_counter() - t0)
    t0 = time.perf_counter(); ref.batch(False, rc, ro, rl.astype(np.uint64), dst_len=ln, threads=cores, dict_data=dct, gather=False); td = min(td, time.perf_counter() - t0)
print("CPU reference (%d threads): compress %.2f GB/s, decompress %.2f GB/s, ratio %.2f" % (cores, U / tc / 1e9, U / td / 1e9, U / float(rl.sum())), flush=True)

d = zstd.ZstdCompressionDict(dct)
pin = zstd.PinnedBuffer(len(blob)); np.frombuffer(pin, dtype=np.uint8)[:] = blob
bws = zstd.BufferWithSegments(pin, np.stack([off, ln], axis=1).astype(np.uint64).tobytes())
cctx = zstd.ZstdCompressor(level=3, dict_data=d)
dctx = zstd.ZstdDecompressor(dict_data=d)
ctx = _native.Context.get(0)
res = cctx.multi_compress_to_buffer(bws)
csz = res.size()
tg = 1e9
for _ in range(3):
    t0 = time.perf_counter(); r2 = cctx.multi_compress_to_buffer(bws); tg = min(tg, time.perf_counter() - t0); del r2
ctx.profile(True); r2 = cctx.multi_compress_to_buffer(bws); pk = ctx.profile_read(); ctx.profile(False); del r2
print("GPU compress e2e %.2f GB/s (ratio %.2f, %+.1f%% vs reference), kernels %s" % (
    U / tg / 1e9, U / csz, 100.0 * (csz / float(rl.sum()) - 1), {k: round(v[0], 3) for k, v in pk.items()}), flush=True)

import ctypes as C
L = _native.lib(); L.zb_encode_phase_read.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_uint64 * 16)(); L.zb_encode_phase_read(buf, 1)
r2 = cctx.multi_compress_to_buffer(bws); del r2
L.zb_encode_phase_read(buf, 1)
names = ["setup/rle", "A hash links", "C parse", "D compaction+gather", "E tables", "E literals", "E chains", "E seq pack", "F assemble"]
print("compress phases, cycles per record:", {nm: int(buf[i] / n) for i, nm in enumerate(names)}, flush=True)

# decode: the reference's frames (host pinned) through the public API
pin2 = zstd.PinnedBuffer(len(rc)); np.frombuffer(pin2, dtype=np.uint8)[:] = rc
fbws = zstd.BufferWithSegments(pin2, np.stack([ro, rl.astype(np.uint64)], axis=1).astype(np.uint64).tobytes())
out = dctx.multi_decompress_to_buffer(fbws)
ok = all(out[i].tobytes() == recs[i] for i in range(0, n, max(1, n // 997)))
tgd = 1e9
for _ in range(3):
    t0 = time.perf_counter(); o2 = dctx.multi_decompress_to_buffer(fbws); tgd = min(tgd, time.perf_counter() - t0); del o2
ctx.profile(True); o2 = dctx.multi_decompress_to_buffer(fbws); pk = ctx.profile_read(); ctx.profile(False); del o2
print("GPU decompress e2e %.2f GB/s (sampled equal: %s), kernels of context 0 %s" % (U / tgd / 1e9, ok, {k: round(v[0], 3) for k, v in pk.items()}), flush=True)
# our own frames decode too
out2 = dctx.multi_decompress_to_buffer(res)
print("our dictionary frames decode on the GPU:", all(out2[i].tobytes() == recs[i] for i in range(0, n, max(1, n // 997))))
