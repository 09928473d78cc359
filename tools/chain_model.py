"""A cost model of the in-order execute of ONE frame (0.3 us per in-tile hop, 2 us per hop through global memory): completion\ntime of every block for three publication granularities.  CPU only; see DESIGN.md section 4."""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np, corpus
from oracle import Oracle, RefZstd
ref = RefZstd(); orc = Oracle()
t = corpus.text_corpus(8 << 20)
N = 6 << 20
data = bytes(t[:N])
frame = ref.compress(data, level=3)
import ctypes as C
class Seq(C.Structure):
    _fields_ = [("ll", C.c_uint32), ("ml", C.c_uint32), ("off", C.c_uint32)]
class Trace(C.Structure):
    _fields_ = [("seqs", C.POINTER(Seq)), ("seq_cap", C.c_size_t), ("n_seqs", C.c_size_t),
                ("lits", C.c_void_p), ("lit_cap", C.c_size_t), ("n_lits", C.c_size_t),
                ("block_nseq", C.POINTER(C.c_uint32)), ("block_nlit", C.POINTER(C.c_uint32)),
                ("block_cap", C.c_size_t), ("n_blocks", C.c_size_t)]
seq_cap = 1 << 21
sq = (Seq * seq_cap)(); lt = C.create_string_buffer(len(data) + 1); bcap = 4096
bnn = (C.c_uint32 * bcap)(); bll = (C.c_uint32 * bcap)()
tr = Trace(sq, seq_cap, 0, C.cast(lt, C.c_void_p), len(data) + 1, 0, bnn, bll, bcap, 0)
ob = C.create_string_buffer(len(data)); err = C.c_int(0); used = C.c_size_t(0)
n = orc.L.zo_decompress_frame(ob, len(data), frame, len(frame), None, 0, C.byref(used), C.byref(tr), C.byref(err))
assert n == len(data) and ob.raw == data
A = np.frombuffer(sq, dtype=np.uint32, count=3 * tr.n_seqs).reshape(-1, 3)
seqs = A.tolist(); bn = [bnn[i] for i in range(tr.n_blocks)]; bl = [bll[i] for i in range(tr.n_blocks)]
bsz = []; si_ = 0
for ns, nl in zip(bn, bl):
    a = A[si_:si_+ns]; bsz.append(int(a[:,1].sum()) + nl); si_ += ns
assert sum(bsz) == len(data), (sum(bsz), len(data))
print("blocks", len(bn), "sizes min/max", min(bsz), max(bsz))

HOP_IN, HOP_EXT = 0.3, 2.0
def run(gran):
    # gran: publication granularity in bytes (0 = whole block)
    T = np.zeros(len(data), dtype=np.float32)
    pos = 0; si = 0; res = []
    for b, ns in enumerate(bn):
        bstart = pos; bend = bstart + bsz[b]
        for (ll, ml, off) in seqs[si:si+ns]:
            pos += ll
            s = pos - off; span = min(ml, off)
            tt = 0.0
            if s < bstart:   # external part
                e = min(s + span, bstart)
                tt = float(T[s:e].max()) + HOP_EXT
            if s + span > bstart:
                a = max(s, bstart)
                if s + span > a:
                    assert 0 <= a < s+span <= len(data), (b, pos, ll, ml, off, s, span, a, bstart)
                    tt = max(tt, float(T[a:s+span].max()) + HOP_IN)
            T[pos:pos+ml] = tt
            pos += ml
        si += ns
        pos = bend
        # publication: bytes become visible to other blocks when the prefix up to their granule is complete
        blk = T[bstart:bend]
        if gran == 0:
            blk[:] = blk.max()
        else:
            pm = np.maximum.accumulate(blk)          # prefix-complete time
            idx = np.minimum((np.arange(len(blk)) // gran + 1) * gran - 1, len(blk) - 1)
            blk[:] = pm[idx]
        res.append(float(blk.max()))
    return res
for g in (0, 65536, 4096):
    r = run(g)
    print("granularity %6d: completion time of block 8: %.0f us, block 24: %.0f, last (%d): %.0f us -> %.2f us per block" % (g, r[8], r[24], len(r)-1, r[-1], (r[-1]-r[8])/(len(r)-1-8)))
