"""Development check: compress on the GPU, decode with the reference + oracle, compare sizes with level 3."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import corpus
from oracle import RefZstd, Oracle
import python_zstandard_b200 as zb

r = RefZstd(); orc = Oracle()
c = zb.ZstdCompressor()
text = corpus.text_corpus().tobytes()
rng = np.random.default_rng(5)
cases = {
    "empty": b"", "one": b"a", "foo12": b"foo" * 12, "x64": b"x" * 64, "text100": text[:100], "text1k": text[:1000], "text4k": text[5000:9096],
    "text64k": text[:65536], "text128k": text[:131072], "text300k": text[:300000], "rand5k": rng.integers(0, 256, 5000, dtype=np.uint8).tobytes(),
    "zeros100k": b"\0" * 100000, "abcd": bytes(rng.choice(list(b"abcd"), 20000).astype(np.uint8)), "bin": corpus.binary_blob(100000).tobytes(),
    "text1m": text[:1 << 20],
}
bad = 0
for ck in (False, True):
    for name, d in cases.items():
        try:
            f = zb.ZstdCompressor(write_checksum=ck).compress(d)
        except Exception as e:
            print("COMPRESS FAIL", name, e); bad += 1; continue
        try:
            out = r.decompress(f, len(d))
            ok = out == d
        except Exception as e:
            ok = False; print("   ref decode error", name, e)
        try:
            out2 = orc.decompress(f, len(d)); ok2 = out2 == d
        except Exception as e:
            ok2 = False; print("   oracle decode error", name, e)
        l3 = len(r.compress(d, level=3, checksum=ck))
        print("%-10s ck=%d n=%7d ours=%7d ref3=%7d (%+.1f%%) %s %s" % (name, ck, len(d), len(f), l3, 100.0 * (len(f) / max(l3, 1) - 1), "OK" if ok else "BAD", "OK" if ok2 else "BAD-oracle"))
        bad += (not ok) + (not ok2)
print("bad =", bad)
if bad: sys.exit(1)
# batches
for label, (blob, off, ln) in {"text4k x8192": corpus.text_segments(8192, 4096), "mix128k x256": corpus.silesia_mix(256, 131072), "text128k x256": corpus.text_segments(256, 131072)}.items():
    segs = np.stack([off, ln], axis=1).astype(np.uint64)
    bws = zb.BufferWithSegments(blob, segs.tobytes())
    from python_zstandard_b200 import _native
    ctx = _native.Context.get(0); ctx.profile(True)
    for it in range(2):
        t = time.time(); res = c.multi_compress_to_buffer(bws); dt = time.time() - t
    prof = ctx.profile_read(); ctx.profile(False)
    cb = res._buffers[0]
    csegs = np.frombuffer(cb._segments, dtype=np.uint64).reshape(-1, 2)
    cblob = np.frombuffer(cb._data, dtype=np.uint8)
    rb, rl = r.batch(False, cblob, np.ascontiguousarray(csegs[:, 0]), np.ascontiguousarray(csegs[:, 1]), threads=os.cpu_count())
    ok = np.array_equal(rb, blob)
    refc, _ = r.batch(True, blob, off, ln, level=3, threads=os.cpu_count())
    k = prof.get("zb_compress_blocks", (0, 1))
    print("%s: roundtrip %s  ours %d ref3 %d (%+.2f%%) ratio %.3f  e2e %.1f ms (%.2f GB/s)  kernel %.2f ms (%.1f GB/s)" % (
        label, ok, len(cblob), len(refc), 100.0 * (len(cblob) / len(refc) - 1), len(blob) / len(cblob), dt * 1e3, len(blob) / dt / 1e9,
        k[0] / k[1], len(blob) / (k[0] / k[1] * 1e-3) / 1e9))
    # and our own decoder
    d = zb.ZstdDecompressor().multi_decompress_to_buffer(res)
    print("   own decode equal:", np.array_equal(np.frombuffer(d._buffers[0]._data, dtype=np.uint8), blob))
