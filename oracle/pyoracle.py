"""ctypes front-ends of the checker libraries.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

``Oracle``  -> oracle/libzo.so      (plain-C restatement, zstd_oracle.c)
``RefZstd`` -> oracle/_ref/*.so     (unmodified reference codec + batch driver)
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ZO = os.path.join(_HERE, "libzo.so")
_REF = os.path.join(_HERE, "_ref", "libzstd_ref.so")
_RB = os.path.join(_HERE, "_ref", "libref_batch.so")

# ZSTD_cParameter values (zstd/zstd.h:343-475)
ZSTD_c_compressionLevel = 100
ZSTD_c_contentSizeFlag = 200
ZSTD_c_checksumFlag = 201
ZSTD_c_dictIDFlag = 202


def build(need_ref=False):
    """Compile the checker.  The reference codec is only (re)built when its sources
    are present (this container); the GPU box uses the prebuilt files."""
    if not os.path.exists(_ZO) or os.path.getmtime(_ZO) < os.path.getmtime(os.path.join(_HERE, "zstd_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libzo.so"])
    if os.path.exists("/root/reference/zstd/zstd.c"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    if need_ref and not have_ref():
        raise RuntimeError("oracle/_ref is not built and /root/reference is absent")


def have_ref():
    return os.path.exists(_REF) and os.path.exists(_RB)


class Oracle:
    """Plain-C restatement of the reference decoder (libzo.so)."""

    def __init__(self):
        build()
        L = C.CDLL(_ZO)
        L.zo_decompress.restype = C.c_size_t
        L.zo_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_int)]
        L.zo_decompress_frame.restype = C.c_size_t
        L.zo_decompress_frame.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                          C.POINTER(C.c_size_t), C.c_void_p, C.POINTER(C.c_int)]
        L.zo_error_name.restype = C.c_char_p
        L.zo_error_name.argtypes = [C.c_int]
        L.zo_xxh64.restype = C.c_uint64
        L.zo_xxh64.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64]
        L.zo_find_frame_compressed_size.restype = C.c_size_t
        L.zo_find_frame_compressed_size.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int)]
        self.L = L

    class Error(Exception):
        pass

    def decompress(self, frame, max_out, dict_data=b""):
        frame = bytes(frame)
        out = C.create_string_buffer(max(max_out, 1))
        err = C.c_int(0)
        n = self.L.zo_decompress(out, max_out, frame, len(frame), dict_data or None, len(dict_data), C.byref(err))
        if err.value:
            raise Oracle.Error(self.L.zo_error_name(err.value).decode())
        return out.raw[:n]

    def trace(self, frame, max_out, dict_data=b"", seq_cap=1 << 20):
        """Decode one frame and return (bytes, literals, [(ll, ml, off)], block_nseq)."""
        class Seq(C.Structure):
            _fields_ = [("ll", C.c_uint32), ("ml", C.c_uint32), ("off", C.c_uint32)]

        class Trace(C.Structure):
            _fields_ = [("seqs", C.POINTER(Seq)), ("seq_cap", C.c_size_t), ("n_seqs", C.c_size_t),
                        ("lits", C.c_void_p), ("lit_cap", C.c_size_t), ("n_lits", C.c_size_t),
                        ("block_nseq", C.POINTER(C.c_uint32)), ("block_nlit", C.POINTER(C.c_uint32)),
                        ("block_cap", C.c_size_t), ("n_blocks", C.c_size_t)]
        frame = bytes(frame)
        seqs = (Seq * seq_cap)()
        lits = C.create_string_buffer(max_out + 1)
        bcap = max_out // 1 + 16
        bn = (C.c_uint32 * bcap)()
        bl = (C.c_uint32 * bcap)()
        tr = Trace(seqs, seq_cap, 0, C.cast(lits, C.c_void_p), max_out + 1, 0, bn, bl, bcap, 0)
        out = C.create_string_buffer(max(max_out, 1))
        err = C.c_int(0)
        used = C.c_size_t(0)
        n = self.L.zo_decompress_frame(out, max_out, frame, len(frame), dict_data or None, len(dict_data),
                                       C.byref(used), C.byref(tr), C.byref(err))
        if err.value:
            raise Oracle.Error(self.L.zo_error_name(err.value).decode())
        return (out.raw[:n], lits.raw[:tr.n_lits],
                [(seqs[i].ll, seqs[i].ml, seqs[i].off) for i in range(tr.n_seqs)],
                [bn[i] for i in range(tr.n_blocks)])

    def xxh64(self, data, seed=0):
        return self.L.zo_xxh64(bytes(data), len(data), seed)

    def frame_compressed_size(self, data):
        err = C.c_int(0)
        n = self.L.zo_find_frame_compressed_size(bytes(data), len(data), C.byref(err))
        if err.value:
            raise Oracle.Error(self.L.zo_error_name(err.value).decode())
        return n


class RefZstd:
    """The unmodified reference codec (vendored zstd 1.5.7) through ctypes."""

    class Error(Exception):
        pass

    def __init__(self):
        build()
        if not have_ref():
            raise RuntimeError("oracle/_ref not built (run `make -C oracle ref` where /root/reference exists)")
        Z = C.CDLL(_REF, mode=C.RTLD_GLOBAL)
        for name, res, args in [
            ("ZSTD_createCCtx", C.c_void_p, []),
            ("ZSTD_freeCCtx", C.c_size_t, [C.c_void_p]),
            ("ZSTD_CCtx_setParameter", C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
            ("ZSTD_CCtx_loadDictionary", C.c_size_t, [C.c_void_p, C.c_char_p, C.c_size_t]),
            ("ZSTD_compress2", C.c_size_t, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]),
            ("ZSTD_compressBound", C.c_size_t, [C.c_size_t]),
            ("ZSTD_createDCtx", C.c_void_p, []),
            ("ZSTD_freeDCtx", C.c_size_t, [C.c_void_p]),
            ("ZSTD_DCtx_loadDictionary", C.c_size_t, [C.c_void_p, C.c_char_p, C.c_size_t]),
            ("ZSTD_decompressDCtx", C.c_size_t, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]),
            ("ZSTD_isError", C.c_uint, [C.c_size_t]),
            ("ZSTD_getErrorName", C.c_char_p, [C.c_size_t]),
            ("ZSTD_versionNumber", C.c_uint, []),
            ("ZDICT_trainFromBuffer", C.c_size_t, [C.c_void_p, C.c_size_t, C.c_char_p, C.POINTER(C.c_size_t), C.c_uint]),
            ("ZDICT_isError", C.c_uint, [C.c_size_t]),
        ]:
            f = getattr(Z, name)
            f.restype = res
            f.argtypes = args
        self.Z = Z
        B = C.CDLL(_RB)
        B.rb_run.restype = C.c_void_p
        B.rb_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_void_p, C.c_size_t, C.c_int]
        B.rb_error.restype = C.c_int
        B.rb_error.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_char_p)]
        B.rb_total.restype = C.c_uint64
        B.rb_total.argtypes = [C.c_void_p]
        B.rb_gather.restype = None
        B.rb_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        B.rb_free.restype = None
        B.rb_free.argtypes = [C.c_void_p]
        self.B = B
        assert Z.ZSTD_versionNumber() == 10507, Z.ZSTD_versionNumber()

    # -- one-shot -------------------------------------------------------
    def compress(self, data, level=3, dict_data=b"", checksum=False, content_size=True, dict_id=True):
        data = bytes(data)
        Z = self.Z
        c = Z.ZSTD_createCCtx()
        try:
            Z.ZSTD_CCtx_setParameter(c, ZSTD_c_compressionLevel, level)
            Z.ZSTD_CCtx_setParameter(c, ZSTD_c_checksumFlag, int(checksum))
            Z.ZSTD_CCtx_setParameter(c, ZSTD_c_contentSizeFlag, int(content_size))
            Z.ZSTD_CCtx_setParameter(c, ZSTD_c_dictIDFlag, int(dict_id))
            if dict_data:
                Z.ZSTD_CCtx_loadDictionary(c, dict_data, len(dict_data))
            cap = Z.ZSTD_compressBound(len(data))
            out = C.create_string_buffer(cap)
            n = Z.ZSTD_compress2(c, out, cap, data, len(data))
            if Z.ZSTD_isError(n):
                raise RefZstd.Error(Z.ZSTD_getErrorName(n).decode())
            return out.raw[:n]
        finally:
            Z.ZSTD_freeCCtx(c)

    def decompress(self, frame, max_out, dict_data=b""):
        frame = bytes(frame)
        Z = self.Z
        d = Z.ZSTD_createDCtx()
        try:
            if dict_data:
                Z.ZSTD_DCtx_loadDictionary(d, dict_data, len(dict_data))
            out = C.create_string_buffer(max(max_out, 1))
            n = Z.ZSTD_decompressDCtx(d, out, max_out, frame, len(frame))
            if Z.ZSTD_isError(n):
                raise RefZstd.Error(Z.ZSTD_getErrorName(n).decode())
            return out.raw[:n]
        finally:
            Z.ZSTD_freeDCtx(d)

    def train_dictionary(self, dict_size, samples):
        blob = b"".join(samples)
        sizes = (C.c_size_t * len(samples))(*[len(s) for s in samples])
        out = C.create_string_buffer(dict_size)
        n = self.Z.ZDICT_trainFromBuffer(out, dict_size, blob, sizes, len(samples))
        if self.Z.ZDICT_isError(n):
            raise RefZstd.Error("dictionary training failed")
        return out.raw[:n]

    # -- batch path (numpy arrays: src uint8, off/len uint64) -------------
    def batch(self, compress, src, off, ln, dst_len=None, level=3, threads=1, dict_data=b"", checksum=False,
              gather=True):
        """Run the reference batch orchestration.  Returns (blob uint8, lens uint64) or the
        total output byte count when gather=False (timing runs)."""
        import numpy as np
        assert src.dtype == np.uint8 and off.dtype == np.uint64 and ln.dtype == np.uint64
        n = len(off)
        h = self.B.rb_run(int(compress), level, int(checksum), dict_data or None, len(dict_data),
                          src.ctypes.data, off.ctypes.data, ln.ctypes.data,
                          dst_len.ctypes.data if dst_len is not None else None, n, threads)
        try:
            item = C.c_size_t(0)
            msg = C.c_char_p()
            e = self.B.rb_error(h, C.byref(item), C.byref(msg))
            if e:
                raise RefZstd.Error("item %d: %s" % (item.value, msg.value.decode()))
            total = self.B.rb_total(h)
            if not gather:
                return total
            blob = np.empty(total, dtype=np.uint8)
            lens = np.empty(n, dtype=np.uint64)
            self.B.rb_gather(h, blob.ctypes.data, lens.ctypes.data)
            return blob, lens
        finally:
            self.B.rb_free(h)
