"""Config-3-style check: compress 128 KiB Silesia-mix segments on the GPU, then time the decode of (a) our frames
and (b) the reference's level-3 frames of the same input (device resident)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, corpus
from oracle import RefZstd
import python_zstandard_b200 as zstd
from python_zstandard_b200 import _native
n = int(os.environ.get("N", "2048"))
ref = RefZstd()
blob, off, ln = corpus.silesia_mix(n, 131072)
ctx = _native.Context.get(0); L = ctx.L
def decode_timed(cblob, coff, clens, label):
    L.zb_entropy_phase_read.argtypes = [C.c_void_p, C.c_int]; L.zb_entropy_phase_read((C.c_uint64 * 8)(), 1)
    segs = np.stack([coff, clens], axis=1).astype(np.uint64)
    d_src = torch.empty(len(cblob) + 256, dtype=torch.uint8, device="cuda"); d_src[:len(cblob)].copy_(torch.from_numpy(cblob))
    d_segs = torch.from_numpy(segs.view(np.int64).copy()).cuda()
    def step():
        res = C.c_void_p(); ctx.check(L.zb200_decompress_batch(ctx.h, d_src.data_ptr(), d_segs.data_ptr(), n, None, None, 3, C.byref(res)), "dec"); return res
    r = step(); out = np.empty(len(blob), dtype=np.uint8); L.zb200_memcpy_d2h(ctx.h, out.ctypes.data, L.zb200_result_data(r), len(blob)); L.zb200_result_free(r)
    ok = np.array_equal(out, blob)
    for _ in range(2): L.zb200_result_free(step())
    ctx.profile(True)
    for _ in range(3): L.zb200_result_free(step())
    p = ctx.profile_read(); ctx.profile(False)
    L.zb_entropy_phase_read.argtypes = [C.c_void_p, C.c_int]; buf = (C.c_uint64 * 8)(); L.zb_entropy_phase_read(buf, 1)
    names = ['other/loop', 'A block header', 'B literals hdr+weights', 'huffman table+streams', 'C seq header+ncount', 'D tables+sequences']
    print('   phase cycles (sum over warps):', {nm: int(buf[i]) for i, nm in enumerate(names)})
    tot = sum(v[0] / v[1] for v in p.values())
    print(label, "equal", ok, {k: round(v[0] / v[1], 3) for k, v in p.items()}, "-> %.1f GB/s" % (len(blob) / tot / 1e6))
rc, rl = ref.batch(True, blob, off, ln, level=3, threads=os.cpu_count())
ro = np.concatenate([[0], np.cumsum(rl)[:-1]]).astype(np.uint64)
decode_timed(rc, ro, rl.astype(np.uint64), "reference-made 128 KiB frames:")
# (a) our own frames of the same segments: ~10 blocks each, so the batch takes the block-parallel path (ZB200_BLOCK_PATH=0 for the other)
segs_in = np.stack([off, ln], axis=1).astype(np.uint64)
coll = zstd.ZstdCompressor(level=3).multi_compress_to_buffer(zstd.BufferWithSegments(blob.tobytes(), segs_in.tobytes()))
ours = [coll[i].tobytes() for i in range(n)]
ol = np.array([len(x) for x in ours], dtype=np.uint64); oo = np.concatenate([[0], np.cumsum(ol)[:-1]]).astype(np.uint64)
decode_timed(np.frombuffer(b"".join(ours), dtype=np.uint8), oo, ol, "our own 128 KiB frames (ZB200_BLOCK_PATH=%s):" % os.environ.get("ZB200_BLOCK_PATH", "auto"))
