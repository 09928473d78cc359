"""Per-phase clock64 cycles of zb_compress_smem (tuning build: ZB200_PHASE_TIMERS=1 python -m python_zstandard_b200.build).
  N=1184 SIZE=131072 MIX=0 python tools/gpu_phase_encode2.py"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, corpus
import python_zstandard_b200 as zb
from python_zstandard_b200 import _native
n = int(os.environ.get("N", "1184")); size = int(os.environ.get("SIZE", "131072"))
mix = os.environ.get("MIX", "0") == "1"
blob, off, ln = (corpus.silesia_mix if mix else corpus.text_segments)(n, size)
segs = np.stack([off, ln], axis=1).astype(np.uint64)
bws = zb.BufferWithSegments(blob, segs.tobytes())
c = zb.ZstdCompressor()
L = _native.lib()
L.zb_encode2_phase_read.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_uint64 * 16)()
c.multi_compress_to_buffer(bws)
L.zb_encode2_phase_read(buf, 1)
ctx = _native.Context.get(0); ctx.profile(True)
res = c.multi_compress_to_buffer(bws)
print(ctx.profile_read())
L.zb_encode2_phase_read(buf, 1)
names = ["load/trivial", "match finding", "stitch+gather", "tables", "literals", "sequences", "assemble"]
tot = sum(buf[i] for i in range(7))
for i, nm in enumerate(names):
    print("%-16s %10.0f cycles/block  %5.1f%%" % (nm, buf[i] / n, 100.0 * buf[i] / max(tot, 1)))
print("total cycles/block", tot / n, "(%s, %d x %d)" % ("mix" if mix else "text", n, size))
