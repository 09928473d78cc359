"""Per-phase clock64 cycles of zb_compress_smem (tuning build: ZB200_PHASE_TIMERS=1 python -m python_zstandard_b200.build,
loaded with ZB200_LIB=.../libzb200_timers.so).  Device-resident input, so no phase waits for an upload.
  N=1184 SIZE=131072 MIX=0 python tools/gpu_phase_encode2.py"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, corpus, torch
import python_zstandard_b200 as zstd
from python_zstandard_b200 import _native
n = int(os.environ.get("N", "1184")); size = int(os.environ.get("SIZE", "131072"))
mix = os.environ.get("MIX", "0") == "1"
blob, off, ln = (corpus.silesia_mix if mix else corpus.text_segments)(n, size)
segs = np.stack([off, ln], axis=1).astype(np.uint64)
L = _native.lib()
ctx = _native.Context.get(0)
d_in = torch.empty(len(blob) + 256, dtype=torch.uint8, device="cuda"); d_in[:len(blob)].copy_(torch.from_numpy(blob))
d_segs = torch.from_numpy(segs.view(np.int64).copy()).cuda()
p = zstd.compressor.CParams(3, 0, 1, 0)
def step():
    r = C.c_void_p()
    ctx.check(L.zb200_compress_batch(ctx.h, d_in.data_ptr(), d_segs.data_ptr(), n, C.byref(p), None, _native.SRC_DEVICE | _native.DST_DEVICE, C.byref(r)), "compress")
    csz = int(L.zb200_result_size(r)); L.zb200_result_free(r); return csz
L.zb_encode2_phase_read.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_uint64 * 16)()
step(); L.zb_encode2_phase_read(buf, 1)
ctx.profile(True); csz = step(); pr = ctx.profile_read(); ctx.profile(False)
print({k: round(v[0], 3) for k, v in pr.items()}, "ratio %.3f" % (len(blob) / csz))
L.zb_encode2_phase_read(buf, 1)
names = {0: "load/trivial", 1: "mf prologue H0+L0", 8: "mf passes", 9: "stitch", 2: "gather+hist", 3: "tables+chains+Huffman", 7: "  (FSE tables, w0)", 13: "  (chains, w0)", 15: "  (Huffman, w3)",
         11: "sub-blocks+lit count", 10: "-", 12: "seq bits+layout", 4: "headers+lit pack", 5: "seq pack", 6: "end"}
tot = sum(buf[i] for i in (0, 1, 8, 9, 2, 3, 11, 10, 12, 4, 5, 6))
for i in (0, 1, 8, 9, 2, 3, 7, 13, 15, 11, 12, 4, 5, 6):
    print("%-22s %10.0f cycles/block  %5.1f%%" % (names[i], buf[i] / n, 100.0 * buf[i] / max(tot, 1)))
print("total cycles/block %.0f (%s, %d x %d)" % (tot / n, "mix" if mix else "text", n, size))
