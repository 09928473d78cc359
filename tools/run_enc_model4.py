"""CPU ratio model of the round-2 match finder (tools/enc_model4.c): sizes vs the reference's level 3."""
import ctypes as C, numpy as np, sys, os, subprocess
sys.path.insert(0, '/root/repo')
import corpus
from oracle import RefZstd
subprocess.check_call("gcc -O2 -fPIC -shared -I/root/reference/zstd -o /tmp/enc_model4.so tools/enc_model4.c -Loracle/_ref -lzstd_ref -Wl,-rpath,/root/repo/oracle/_ref", shell=True, cwd='/root/repo')
r = RefZstd()
M = C.CDLL('/tmp/enc_model4.so')
class P(C.Structure):
    _fields_ = [(k, C.c_int) for k in "SC near_bits tag_bits far_bits far_mml R cap lazy U rep_probe backext maxdist near_mml exact_step step small_mask".split()]
M.model4_compress.restype = C.c_size_t
M.model4_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(P), C.POINTER(C.c_size_t)]
BASE = dict(SC=1024, near_bits=9, tag_bits=6, far_bits=14, far_mml=5, R=1024, cap=15, lazy=1, U=132, rep_probe=0, backext=1, maxdist=65535, near_mml=4, exact_step=0, step=32, small_mask=0)
def variants():
    yield "base", {}
    for a in sys.argv[1:]:
        d = {}
        for kv in a.split(","):
            k, v = kv.split("="); d[k] = int(v)
        yield a, d
sets = {}
b, o, l = corpus.text_segments(48, 131072); sets["text128K"] = (b, o, l)
b, o, l = corpus.silesia_mix(80, 131072); sets["mix128K"] = (b, o, l)
b, o, l = corpus.text_segments(512, 4096); sets["text4K"] = (b, o, l)
refs = {}
for k, (b, o, l) in sets.items():
    _, rl = r.batch(True, b, o, l, level=3, threads=8); refs[k] = int(rl.sum())
for name, d in variants():
    pd = dict(BASE); pd.update(d); p = P(**pd)
    res = []
    for k, (b, o, l) in sets.items():
        tot = 0; nst = 0
        for oo, ll in zip(o, l):
            seg = b[int(oo):int(oo) + int(ll)]
            ns = C.c_size_t()
            tot += M.model4_compress(seg.ctypes.data, len(seg), 131072, C.byref(p), C.byref(ns)); nst += ns.value
        res.append("%s %+.2f%% (%d seq)" % (k, 100.0 * (tot / refs[k] - 1), nst // len(o)))
    print("%-50s %s" % (name, " | ".join(res)), flush=True)
