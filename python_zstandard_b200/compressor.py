"""ZstdCompressor -- the compression half of the reference API that sits on the batch path.

Mirrors c-ext/compressor.c: constructor :88-261 (same keyword arguments, defaults and error
texts), compress() :509-574 and multi_compress_to_buffer() :1341-1504.  The frames come from the
CUDA block compressor in libzb200 (zb_encode.cu); they are RFC 8878 zstd and decode with any zstd
decoder, but they are not byte-identical to CPU zstd's -- the parse is this project's own.
With dict_data, the last 32 KiB of the dictionary content act as match history in front of every frame, the frame
starts from the dictionary's repcodes, and its entropy tables are reused where that is cheaper ("repeat" FSE tables,
"treeless" literals), like the reference's ZSTD_compress_insertDictionary (zstd/zstd.c:28015-28180).
"""
import ctypes as C

import numpy as np

from . import _native
from .buffers import BufferWithSegments, BufferWithSegmentsCollection, DeviceBufferWithSegments
from .decompressor import _devices
from .dictionary import ZstdCompressionDict
from .errors import ZstdError

MAX_COMPRESSION_LEVEL = 22


class CParams(C.Structure):
    _fields_ = [("level", C.c_int32), ("write_checksum", C.c_uint32), ("write_content_size", C.c_uint32),
                ("dict_id", C.c_uint32), ("window_log", C.c_uint32), ("reserved", C.c_uint32 * 3)]


class ZstdCompressionParameters:
    """The subset of c-ext/compressionparams.c that reaches this backend: level, the three frame flags and window_log.
    The match-finder knobs (hash_log, chain_log, search_log, min_match, target_length, strategy) and the LDM / job knobs
    configure CPU zstd's strategies, which this backend does not have: anything but their defaults is rejected loudly."""

    _FIELDS = ("format", "compression_level", "window_log", "hash_log", "chain_log", "search_log", "min_match",
               "target_length", "strategy", "write_content_size", "write_checksum", "write_dict_id", "job_size",
               "overlap_log", "force_max_window", "enable_ldm", "ldm_hash_log", "ldm_min_match",
               "ldm_bucket_size_log", "ldm_hash_rate_log", "threads")

    def __init__(self, **kw):
        for k in kw:
            if k not in self._FIELDS:
                raise TypeError("'%s' is an invalid keyword argument" % k)
        self.format = kw.get("format", 0)
        self.compression_level = kw.get("compression_level", 0)
        self.write_content_size = kw.get("write_content_size", 1)
        self.write_checksum = kw.get("write_checksum", 0)
        self.write_dict_id = kw.get("write_dict_id", 0)
        self.threads = kw.get("threads", 0)
        wl = kw.get("window_log", 0) or 0
        if wl and not (10 <= wl <= 31):
            raise ValueError("window_log out of range")        # ZSTD_c_windowLog bounds, zstd/zstd.c:23003
        self.window_log = wl
        for k in ("hash_log", "chain_log", "search_log", "min_match", "target_length", "strategy",
                  "job_size", "overlap_log", "force_max_window", "enable_ldm", "ldm_hash_log", "ldm_min_match",
                  "ldm_bucket_size_log", "ldm_hash_rate_log"):
            v = kw.get(k, 0)
            if v not in (0, -1, None):
                raise ZstdError("compression parameter %s is not supported by the B200 backend" % k)
            setattr(self, k, 0)

    # window_log of ZSTD_getCParams(level, source_size, dict_size): the W column of the four level tables
    # (zstd/zstd.c:30650-30756: any size / <= 256 KB / <= 128 KB / <= 16 KB; row 0 = negative levels, rows 1..22 = levels)
    _LEVEL_WINDOW_LOG = (
        (19, 19, 20, 21, 21, 21, 21, 21, 21, 22, 22, 22, 22, 22, 22, 22, 22, 23, 23, 23, 25, 26, 27),
        (18,) * 23, (17,) * 23, (14,) * 23)

    @classmethod
    def _level_window_log(cls, level, source_size, dict_size):
        """ZSTD_getCParams(...).windowLog: table row by size class (ZSTD_getCParamRowSize / ZSTD_getCParams_internal,
        zstd/zstd.c:30823-30871), then the downsizing of ZSTD_adjustCParams_internal (:24498-24524).  source_size 0 means
        unknown, as in the reference's from_level (c-ext/compressionparams.c:497-560 -> ZSTD_getCParams :30876)."""
        unknown = source_size == 0
        if unknown:
            rsize = (1 << 64) - 1 if dict_size == 0 else dict_size + 499        # (UNKNOWN + dictSize + 500 wraps in the reference)
        else:
            rsize = source_size + dict_size
        table = (rsize <= 256 << 10) + (rsize <= 128 << 10) + (rsize <= 16 << 10)
        row = 3 if level == 0 else (0 if level < 0 else min(level, 22))
        wlog = cls._LEVEL_WINDOW_LOG[table][row]
        if not unknown and source_size <= (1 << 30) and dict_size <= (1 << 30):
            tsize = (source_size + dict_size) & 0xFFFFFFFF
            src_log = 6 if tsize < 64 else (tsize - 1).bit_length()
            wlog = min(wlog, src_log)
        return max(wlog, 10)

    @classmethod
    def from_level(cls, level, source_size=0, dict_size=0, **kwargs):
        """ZstdCompressionParameters.from_level (c-ext/compressionparams.c:497-560): the parameters ZSTD_getCParams picks for
        (level, source_size, dict_size), unless given.  window_log is the one that reaches this backend (frame header, block
        size); the match-finder columns of the level tables describe CPU strategies this backend does not have."""
        if kwargs.get("window_log") in (None, 0, -1):
            kwargs["window_log"] = cls._level_window_log(level, source_size, dict_size)
        return cls(compression_level=level, **kwargs)


class ZstdCompressor:
    def __init__(self, level=3, dict_data=None, compression_params=None, write_checksum=None,
                 write_content_size=None, write_dict_id=None, threads=0):
        if level > MAX_COMPRESSION_LEVEL:
            raise ValueError("level must be less than %d" % (MAX_COMPRESSION_LEVEL + 1))
        if dict_data is not None and not isinstance(dict_data, ZstdCompressionDict):
            raise TypeError("dict_data must be zstd.ZstdCompressionDict")
        if compression_params is not None and not isinstance(compression_params, ZstdCompressionParameters):
            raise TypeError("compression_params must be zstd.ZstdCompressionParameters")
        if compression_params is not None:
            if write_checksum is not None:
                raise ValueError("cannot define compression_params and write_checksum")
            if write_content_size is not None:
                raise ValueError("cannot define compression_params and write_content_size")
            if write_dict_id is not None:
                raise ValueError("cannot define compression_params and write_dict_id")
            if threads:
                raise ValueError("cannot define compression_params and threads")
            self._level = compression_params.compression_level or 3
            self._checksum = bool(compression_params.write_checksum)
            self._content_size = bool(compression_params.write_content_size)
            self._write_dict_id = bool(compression_params.write_dict_id)
            self._window_log = compression_params.window_log
        else:
            self._window_log = 0
            self._level = level
            # defaults: content size on, checksum off, dict id on (c-ext/compressor.c:213-227)
            self._checksum = bool(write_checksum) if write_checksum is not None else False
            self._content_size = bool(write_content_size) if write_content_size is not None else True
            self._write_dict_id = bool(write_dict_id) if write_dict_id is not None else True
        self._dict_data = dict_data
        self._threads = threads

    def _params(self):
        did = self._dict_data.dict_id() if (self._dict_data is not None and self._write_dict_id) else 0
        return CParams(self._level, int(self._checksum), int(self._content_size), did, self._window_log)

    def _dict(self, ctx):
        return self._dict_data._ddict(ctx) if self._dict_data is not None else None

    def memory_size(self):
        return 0

    def frame_progression(self):
        return (0, 0, 0)

    # ------------------------------------------------------------------ one-shot
    def compress(self, data):
        view = memoryview(data)
        buf = np.frombuffer(view, dtype=np.uint8) if view.nbytes else np.zeros(0, dtype=np.uint8)
        ctx = _native.Context.get(_native.default_device())
        L = ctx.L
        seg = np.array([[0, len(buf)]], dtype=np.uint64)
        res = C.c_void_p()
        p = self._params()
        dd = self._dict(ctx)            # (takes ctx.lock itself when it has to build the digest)
        with ctx.lock:
            rc = L.zb200_compress_batch(ctx.h, buf.ctypes.data if len(buf) else None, seg.ctypes.data, 1, C.byref(p),
                                        dd, 0, C.byref(res))
        ctx.check(rc, "zb200_compress_batch")
        try:
            n = L.zb200_result_size(res)
            return C.string_at(L.zb200_result_data(res), n)
        finally:
            L.zb200_result_free(res)

    # ------------------------------------------------------------------ batch
    def multi_compress_to_buffer(self, data, threads=0):
        if isinstance(data, DeviceBufferWithSegments):
            return self._run_device(data)
        if isinstance(data, BufferWithSegments):
            sources = [data]
        elif isinstance(data, BufferWithSegmentsCollection):
            sources = data._buffers
        elif isinstance(data, list):
            sources = None
        else:
            raise TypeError("argument must be list of BufferWithSegments")
        L = _native.lib()
        results = []
        if sources is not None:
            count = sum(len(b) for b in sources)
            total = 0
            for b in sources:
                segs = np.frombuffer(b._segments, dtype=np.uint64).reshape(-1, 2)
                total += int(segs[:, 1].sum()) if len(segs) else 0
            if count == 0:
                raise ValueError("no source elements found")
            if total == 0:
                raise ValueError("source elements are empty")
            for b in sources:
                if len(b) == 0:
                    continue
                segs = np.frombuffer(b._segments, dtype=np.uint64).reshape(-1, 2)
                dat = np.frombuffer(b._data, dtype=np.uint8) if b.size else np.zeros(1, dtype=np.uint8)
                results.extend(self._run(dat.ctypes.data, segs, threads, keep=(dat,)))
            return BufferWithSegmentsCollection(*results)
        views = []
        for i, item in enumerate(data):
            try:
                v = memoryview(item)
            except TypeError:
                raise TypeError("item %d not a bytes like object" % i)
            if not v.contiguous:
                raise TypeError("item %d not a bytes like object" % i)
            views.append(v)
        if not views:
            raise ValueError("no source elements found")
        if sum(v.nbytes for v in views) == 0:
            raise ValueError("source elements are empty")
        lengths = np.array([v.nbytes for v in views], dtype=np.uint64)
        from .decompressor import ZstdDecompressor
        devs = _devices(threads)
        parts = ZstdDecompressor._split(None, lengths, len(devs))
        p = self._params()
        for di, (lo, hi) in enumerate(parts):
            ctx = _native.Context.get(devs[di])
            k = hi - lo
            arrs = [np.frombuffer(v, dtype=np.uint8) if v.nbytes else np.zeros(0, dtype=np.uint8) for v in views[lo:hi]]
            ptrs = (C.c_void_p * k)(*[a.ctypes.data if len(a) else None for a in arrs])
            lens = (C.c_size_t * k)(*[len(a) for a in arrs])
            res = C.c_void_p()
            dd = self._dict(ctx)
            with ctx.lock:
                rc = L.zb200_compress_batch_ptrs(ctx.h, ptrs, lens, k, C.byref(p), dd, 0, C.byref(res))
            ctx.check(rc, "zb200_compress_batch_ptrs")
            results.append(BufferWithSegments._from_result(ctx, res))
        return BufferWithSegmentsCollection(*results)

    def _run_device(self, data):
        """Device-resident segments -> device-resident frames."""
        n = len(data)
        if n == 0:
            raise ValueError("no source elements found")
        segs = np.frombuffer(data._segments, dtype=np.uint64).reshape(-1, 2)
        if int(segs[:, 1].sum()) == 0:
            raise ValueError("source elements are empty")
        ctx = _native.Context.get(data.device)
        L = ctx.L
        dd = self._dict(ctx)
        p = self._params()
        res = C.c_void_p()
        with ctx.lock:
            rc = L.zb200_compress_batch(ctx.h, data._ptr, data._segments, n, C.byref(p), dd,
                                        _native.SRC_DEVICE | _native.DST_DEVICE | _native.SEGS_HOST, C.byref(res))
        ctx.check(rc, "zb200_compress_batch")
        return DeviceBufferWithSegments._from_result(ctx, res)

    # one call keeps the whole device busy for milliseconds per 128 KiB block, so sub-batches only pay once a
    # call is large enough that its upload is worth hiding (measured: 256 MiB in one piece 20.5 ms, in 4 pieces 29 ms)
    PIPELINE_DEPTH = 2
    SUB_BATCH_INPUT_BYTES = 1 << 30

    def _run(self, base_ptr, segs, threads, keep=()):
        """Device partition as in the reference (contiguous ranges by bytes), then sub-batches of each range
        kept in flight on extra contexts so that uploads, kernels and downloads overlap."""
        from .decompressor import ZstdDecompressor, _executor
        L = _native.lib()
        devs = _devices(threads)
        lens = np.ascontiguousarray(segs[:, 1])
        parts = ZstdDecompressor._split(None, lens, len(devs))
        p = self._params()
        jobs = []
        for di, (lo, hi) in enumerate(parts):
            nbytes = int(lens[lo:hi].sum())
            k = max(1, min((hi - lo) // 64 or 1, nbytes // self.SUB_BATCH_INPUT_BYTES))
            if k < 2:
                jobs.append((devs[di], 0, lo, hi))
                continue
            for i, (a, c) in enumerate(ZstdDecompressor._split(None, lens[lo:hi], k)):
                jobs.append((devs[di], i % self.PIPELINE_DEPTH, lo + a, lo + c))

        def run(job):
            dev, slot, lo, hi = job
            ctx = _native.Context.get(dev, slot)
            sub = np.ascontiguousarray(segs[lo:hi])
            return BufferWithSegments._from_result(ctx, self._launch(ctx, base_ptr, sub, hi - lo, p))

        if len(jobs) == 1:
            return [run(jobs[0])]
        return list(_executor(self.PIPELINE_DEPTH * len(parts)).map(run, jobs))

    def _launch(self, ctx, base_ptr, sub, n, p):
        res = C.c_void_p()
        dd = self._dict(ctx)
        with ctx.lock:
            rc = ctx.L.zb200_compress_batch(ctx.h, base_ptr, sub.ctypes.data, n, C.byref(p), dd, 0, C.byref(res))
        ctx.check(rc, "zb200_compress_batch")
        return res

    # ------------------------------------------------------------------ out of scope
    def _unsupported(self, *a, **k):
        raise NotImplementedError("streaming compression objects are outside the B200 batch path")

    stream_reader = stream_writer = read_to_iter = copy_stream = chunker = _unsupported

    def compressobj(self, size=-1):
        """c-ext/compressor.c:576-640; the frame is written in one piece at flush() (streams.py)."""
        from .streams import ZstdCompressionObj
        return ZstdCompressionObj(self, size)
