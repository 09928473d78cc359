"""ZstdCompressionDict -- the dictionary object handed to (de)compressors.

Mirrors c-ext/compressiondict.c:164-348 for the *use* of a dictionary (bytes,
dict_id(), as_bytes(), len()).  Training (ZDICT_*, c-ext/compressiondict.c:13-146)
is out of scope: train with the reference and pass the bytes in.
"""
import struct
import threading

from . import _native
from .errors import ZstdError

DICT_TYPE_AUTO = 0
DICT_TYPE_RAWCONTENT = 1
DICT_TYPE_FULLDICT = 2
_DICT_MAGIC = 0xEC30A437


class ZstdCompressionDict:
    def __init__(self, data, dict_type=DICT_TYPE_AUTO, k=0, d=0):
        if dict_type not in (DICT_TYPE_AUTO, DICT_TYPE_RAWCONTENT, DICT_TYPE_FULLDICT):
            raise ValueError("invalid dictionary load mode: %d; must use DICT_TYPE_* constants" % dict_type)
        self._data = bytes(memoryview(data))
        self._dict_type = dict_type
        self.k = k
        self.d = d
        self._ddicts = {}       # device index -> native handle
        self._lock = threading.Lock()
        is_full = len(self._data) >= 8 and struct.unpack_from("<I", self._data)[0] == _DICT_MAGIC
        if dict_type == DICT_TYPE_FULLDICT and not is_full:
            raise ZstdError("dictionary is not a full zstd dictionary")
        self._raw = dict_type == DICT_TYPE_RAWCONTENT or not is_full

    def __len__(self):
        return len(self._data)

    def dict_id(self):
        if self._raw:
            return 0
        return struct.unpack_from("<I", self._data, 4)[0]

    def as_bytes(self):
        return self._data

    def precompute_compress(self, level=0, compression_params=None):
        if level and compression_params:
            raise ValueError("must only specify one of level or compression_params")
        if not level and not compression_params:
            raise ValueError("must specify one of level or compression_params")
        # digests are built lazily per device on first use

    # -- device digest (ensure_ddict, c-ext/compressiondict.c:148-162)
    def _ddict(self, ctx):
        """The device digest of this dictionary on ctx's device: one per device, created under the dictionary's lock and
        the context's lock (several pipeline workers may ask at once), freed when the dictionary object goes away."""
        h = self._ddicts.get(ctx.device)
        if h is not None:
            return h
        with self._lock:
            h = self._ddicts.get(ctx.device)
            if h is None:
                import ctypes as C
                import weakref
                h = C.c_void_p()
                data = self._data
                if self._raw and len(data) >= 8 and struct.unpack_from("<I", data)[0] == _DICT_MAGIC:
                    raise ZstdError("raw-content dictionaries starting with the dictionary magic are not supported")
                with ctx.lock:
                    rc = ctx.L.zb200_ddict_create(ctx.h, data, len(data), C.byref(h))
                if rc != 0:
                    raise ZstdError("unable to load dictionary: %s" % ctx.last_error())
                self._ddicts[ctx.device] = h
                weakref.finalize(self, ctx.L.zb200_ddict_free, h)
        return h
