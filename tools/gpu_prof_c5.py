"""Driver for ncu captures of the block-parallel decode path: one frame of MB MiB, three calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, corpus
from oracle import RefZstd
import python_zstandard_b200 as zstd
mb = int(os.environ.get("MB", "32"))
t = corpus.text_corpus(8 << 20)
data = np.tile(t, (mb << 20) // len(t) + 1)[:mb << 20].tobytes()
frame = RefZstd().compress(data, level=3)
d = zstd.ZstdDecompressor(max_window_size=1 << 31)
for _ in range(3):
    assert d.decompress(frame) == data
print("ok")
