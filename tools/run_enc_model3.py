"""CPU ratio model of the match finder with an optional second (8-byte hash) table: sizes vs the reference's level 3.
Build first:  gcc -O2 -fPIC -shared -I/root/reference/zstd -o /tmp/enc_model3.so tools/enc_model3.c -Loracle/_ref -lzstd_ref -Wl,-rpath,$PWD/oracle/_ref"""
import ctypes as C, numpy as np, sys, os
sys.path.insert(0,'/root/repo')
import corpus
from oracle import RefZstd
r=RefZstd()
M=C.CDLL('/tmp/enc_model3.so')
M.model2_compress.restype=C.c_size_t
M.model2_compress.argtypes=[C.c_void_p,C.c_size_t,C.c_uint32,C.c_uint32,C.c_int,C.c_int,C.c_uint32,C.c_int,C.c_int,C.POINTER(C.c_size_t)]
def run(name, blob, off, ln):
    _, rl = r.batch(True, blob, off, ln, level=3, threads=8)
    ref = int(rl.sum())
    out=[]
    for label, hlog, mml, llog in (("hash4 2^14 (built)",14,4,0),("hash5 2^14",14,5,0),("hash4 2^13 + hash8 2^13",13,4,13),("hash4 2^14 + hash8 2^13",14,4,13),("hash4 2^14 + hash8 2^14",14,4,14),("hash4 2^15",15,4,0)):
        M.model3_set_long(llog)
        tot=0
        for o,l in zip(off,ln):
            seg=blob[int(o):int(o)+int(l)]
            ns=C.c_size_t()
            tot+=M.model2_compress(seg.ctypes.data, len(seg), 131072, 1024, hlog, mml, 32, 1, 65535, C.byref(ns))
        out.append("%s: %+.2f%%"%(label, 100.0*(tot/ref-1)))
    print(name, "| ".join(out))
b,o,l=corpus.text_segments(64,131072); run("text 128K:", b,o,l)
b,o,l=corpus.silesia_mix(96,131072); run("mix 128K:", b,o,l)
b,o,l=corpus.text_segments(1024,4096); run("text 4K:", b,o,l)
