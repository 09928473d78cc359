"""Development check: decode reference-made frames on the GPU and compare with the originals."""
import os, sys, time, struct
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import corpus
from oracle import RefZstd, Oracle
import python_zstandard_b200 as zb

r = RefZstd()
d = zb.ZstdDecompressor()
text = corpus.text_corpus().tobytes()
rng = np.random.default_rng(5)
cases = {
    "empty": b"", "one": b"a", "foo12": b"foo" * 12, "x64": b"x" * 64, "text1k": text[:1000], "text4k": text[5000:9096],
    "text64k": text[:65536], "text300k": text[:300000], "rand5k": rng.integers(0, 256, 5000, dtype=np.uint8).tobytes(),
    "zeros100k": b"\0" * 100000, "abcd": bytes(rng.choice(list(b"abcd"), 20000).astype(np.uint8)),
    "text1m": text[:1 << 20],
}
bad = 0
for lvl in (1, 3, 5, 9, 19, -5):
    for ck in (False, True):
        names = [k for k in cases if len(cases[k]) > 0]
        frames = [r.compress(cases[k], level=lvl, checksum=ck) for k in names]
        try:
            out = d.multi_decompress_to_buffer(frames)
        except Exception as e:
            print("FAIL level", lvl, "ck", ck, type(e).__name__, e); bad += 1
            for k, f in zip(names, frames):
                try:
                    o = d.multi_decompress_to_buffer([f])
                    if o[0].tobytes() != cases[k]: print("   mismatch", k)
                except Exception as e2:
                    print("   ", k, e2)
            continue
        for i, k in enumerate(names):
            if out[i].tobytes() != cases[k]:
                got = out[i].tobytes(); exp = cases[k]
                first = next((j for j in range(min(len(got), len(exp))) if got[j] != exp[j]), -1)
                print("MISMATCH level", lvl, "ck", ck, k, len(got), len(exp), "first diff", first); bad += 1
print("small cases bad =", bad)

# batch of 4 KiB text frames
n = int(os.environ.get("N", "65536"))
blob, off, ln = corpus.text_segments(n, 4096, unique=8192)
t = time.time(); cblob, clens = r.batch(True, blob, off, ln, threads=os.cpu_count()); print("ref compress s", time.time() - t, "ratio", len(blob) / len(cblob))
coff = np.concatenate([[0], np.cumsum(clens)[:-1]]).astype(np.uint64)
segs = np.stack([coff, clens], axis=1).astype(np.uint64)
bws = zb.BufferWithSegments(cblob, segs.tobytes())
from python_zstandard_b200 import _native
ctx = _native.Context.get(0)
ctx.profile(True)
for it in range(3):
    t = time.time(); out = d.multi_decompress_to_buffer(bws); dt = time.time() - t
    print("e2e decompress %.1f ms  %.2f GB/s" % (dt * 1e3, len(blob) / dt / 1e9))
print(ctx.profile_read())
got = np.frombuffer(out[0]._parent._data, dtype=np.uint8) if False else None
res = out._buffers[0]
ok = np.array_equal(np.frombuffer(res._data, dtype=np.uint8), blob)
print("batch equal:", ok)
if not ok:
    g = np.frombuffer(res._data, dtype=np.uint8)
    diff = np.nonzero(g != blob)[0]
    print("n diff", len(diff), "first", diff[:10], "frames", np.unique(diff // 4096)[:20])
