"""Dependency chains of the copy-execute stage on level-3 text: per 128 KiB block, the longest chain of matches that copy\nfrom one another (the reason zb_execute_big is bound at ~0.3 ms per block and zb_chase_* exists).  CPU only."""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np, corpus
from oracle import Oracle, RefZstd
ref = RefZstd(); orc = Oracle()
t = corpus.text_corpus(8 << 20)
data = bytes(t[1<<20:(1<<20)+(1<<20)])
frame = ref.compress(data, level=3)
out, lits, seqs, bn = orc.trace(frame, len(data))
assert out == data
depth = np.zeros(len(data), dtype=np.int32)      # chain depth inside the block (earlier blocks = 0)
depx = np.zeros(len(data), dtype=np.int32)       # depth counted only on chains rooted in the PREVIOUS blocks' bytes
pos = 0; si = 0
for b, ns in enumerate(bn):
    bstart = pos
    md = 0; mdx = 0; hist = []
    for (ll, ml, off) in seqs[si:si+ns]:
        pos += ll
        s = pos - off
        span = min(ml, off)
        src = slice(max(s, bstart), s + span)
        d = 1 + (int(depth[src].max()) if s + span > bstart and src.stop > src.start else 0)
        ext = s < bstart
        dx = 0
        if s + span > bstart and src.stop > src.start: dx = int(depx[src].max())
        if ext: dx = max(dx, 0) + 1 if True else dx
        elif dx > 0: dx += 1
        depth[pos:pos+ml] = d; depx[pos:pos+ml] = dx
        md = max(md, d); mdx = max(mdx, dx)
        hist.append(d)
        pos += ml
    si += ns
    # block end
    nxt = min(len(data), bstart + 131072)
    h = np.bincount(hist)
    print("block %d: nseq %d  max in-block depth %d  max depth of chains rooted outside %d  mean depth %.1f" % (b, ns, md, mdx, np.mean(hist)))
    pos = nxt
