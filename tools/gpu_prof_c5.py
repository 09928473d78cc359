"""ncu target: ONE 256 MiB reference-made frame through decompress() (block-parallel path with the pointer-jumping
execute stage).  ncu --set full -k regex:zb_chase -c 6 ... python tools/gpu_prof_c5.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, corpus
from oracle import RefZstd
import python_zstandard_b200 as zstd
mb = int(os.environ.get("MB", "256"))
t = corpus.text_corpus(8 << 20)
data = np.tile(t, (mb << 20) // len(t) + 1)[:mb << 20].tobytes()
frame = RefZstd().compress(data, level=3)
d = zstd.ZstdDecompressor(max_window_size=1 << 31)
assert d.decompress(frame) == data
print("ok")
