import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, corpus
import python_zstandard_b200 as zb
n = int(os.environ.get("N", "592")); size = int(os.environ.get("SIZE", "131072"))
blob, off, ln = corpus.text_segments(n, size)
segs = np.stack([off, ln], axis=1).astype(np.uint64)
bws = zb.BufferWithSegments(blob, segs.tobytes())
c = zb.ZstdCompressor()
for _ in range(3):
    res = c.multi_compress_to_buffer(bws)
print("ok", res.size())
