"""Driver for ncu captures of the block compressor: device-resident input (no overlapped upload: under ncu's kernel
replay an upload on another stream cannot make progress), three calls.  N, SIZE, MIX as in gpu_compress_quick.py."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, corpus, torch
import python_zstandard_b200 as zstd
from python_zstandard_b200 import _native
n = int(os.environ.get("N", "296")); size = int(os.environ.get("SIZE", "131072"))
mix = os.environ.get("MIX", "0") == "1"
blob, off, ln = (corpus.silesia_mix if mix else corpus.text_segments)(n, size)
segs = np.stack([off, ln], axis=1).astype(np.uint64)
L = _native.lib(); ctx = _native.Context.get(0)
d_in = torch.empty(len(blob) + 256, dtype=torch.uint8, device="cuda"); d_in[:len(blob)].copy_(torch.from_numpy(blob))
d_segs = torch.from_numpy(segs.view(np.int64).copy()).cuda()
p = zstd.compressor.CParams(int(os.environ.get("LEVEL", "3")), 0, 1, 0)
for _ in range(3):
    r = C.c_void_p()
    ctx.check(L.zb200_compress_batch(ctx.h, d_in.data_ptr(), d_segs.data_ptr(), n, C.byref(p), None, _native.SRC_DEVICE | _native.DST_DEVICE, C.byref(r)), "compress")
    csz = int(L.zb200_result_size(r)); L.zb200_result_free(r)
print("ok", csz)
