"""ZstdDecompressor -- the decompression half of the reference API that sits on the batch path.

Mirrors c-ext/decompressor.c: constructor :17-126, decompress() :263-395 and
multi_decompress_to_buffer() :1460-1711 (same arguments, same error types and texts).
The work itself runs as CUDA kernels through libzb200 (include/zb200.h); there is no CPU path.
Streaming objects (stream_reader, copy_stream, ...) are out of scope for this tier.
"""
import ctypes as C
import os

import numpy as np

from . import _native
from .buffers import BufferWithSegments, BufferWithSegmentsCollection, SEGMENT_SIZE
from .dictionary import ZstdCompressionDict
from .errors import ZstdError

FORMAT_ZSTD1 = 0
FORMAT_ZSTD1_MAGICLESS = 1
_UNKNOWN = (1 << 64) - 1
_SIZES_ARE_CAPACITY = 4


_EXECUTORS = {}


def _executor(workers):
    """Persistent worker threads (ctypes calls release the GIL, so sub-batches overlap on the device)."""
    ex = _EXECUTORS.get(workers)
    if ex is None:
        from concurrent.futures import ThreadPoolExecutor
        ex = _EXECUTORS[workers] = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="zb200")
    return ex


def _devices(threads):
    """`threads` of the reference becomes a hint for how many GPUs to spread a batch over:
    <= 1 -> the default device only (see set_device); -1 -> every visible device (cpu_count() in the
    reference).  Returns the list of device indices to use."""
    n = _native.device_count()
    base = _native.default_device()
    if threads is None or threads == 0 or threads == 1:
        return [base]
    k = max(1, n) if threads < 0 else max(1, min(threads, n))
    return [(base + i) % max(n, 1) for i in range(k)]


class ZstdDecompressor:
    def __init__(self, dict_data=None, max_window_size=0, format=FORMAT_ZSTD1):
        if dict_data is not None and not isinstance(dict_data, ZstdCompressionDict):
            raise TypeError("dict_data must be a ZstdCompressionDict")
        if format != FORMAT_ZSTD1:
            raise ZstdError("only FORMAT_ZSTD1 is supported by the B200 backend")
        if max_window_size < 0 or (max_window_size and max_window_size > (1 << 31)):
            raise ZstdError("unable to set max window size: Parameter is out of bound")
        self._dict_data = dict_data
        self._max_window_size = max_window_size
        # enforced by the scan kernel like ZSTD_DCtx_setMaxWindowSize (c-ext/decompressor.c:22-24, zstd/zstd.c:45452)
        self._dparams = _native.DParams(max_window_size, (C.c_uint32 * 2)(0, 0))
        self._format = format
        self._ctx = None

    def _context(self, device=None):
        return _native.Context.get(_native.default_device() if device is None else device)

    def memory_size(self):
        return 0

    # ------------------------------------------------------------------ one-shot
    def decompress(self, data, max_output_size=0, read_across_frames=False, allow_extra_data=True):
        if read_across_frames:
            raise ZstdError("ZstdDecompressor.read_across_frames=True is not yet implemented")
        view = memoryview(data)
        buf = np.frombuffer(view, dtype=np.uint8) if len(view) else np.zeros(0, dtype=np.uint8)
        L = _native.lib()
        info = _native.FrameInfo()
        L.zb200_frame_info(buf.ctypes.data if len(buf) else None, len(buf), C.byref(info))
        if info.status != 0 or len(buf) < 5:
            raise ZstdError("error determining content size from frame header")
        size = info.content_size
        if size == 0:
            return b""
        flags = 0
        if size == _UNKNOWN:
            if not max_output_size:
                raise ZstdError("could not determine content size in frame header")
            size = max_output_size
            flags = _SIZES_ARE_CAPACITY
        ctx = self._context()
        seg = np.array([[0, len(buf)]], dtype=np.uint64)
        sizes = np.array([size], dtype=np.uint64)
        res = C.c_void_p()
        dd = self._dict_data._ddict(ctx) if self._dict_data is not None else None
        with ctx.lock:
            rc = L.zb200_decompress_batch_ex(ctx.h, buf.ctypes.data, seg.ctypes.data, 1, sizes.ctypes.data, dd,
                                             C.byref(self._dparams), flags, C.byref(res))
        ctx.check(rc, "zb200_decompress_batch")
        try:
            item, code = C.c_size_t(), C.c_int()
            got, exp = C.c_uint64(), C.c_uint64()
            if L.zb200_result_first_error(res, C.byref(item), C.byref(code), C.byref(got), C.byref(exp)):
                if code.value == 201:
                    raise ZstdError("decompression error: did not decompress full frame")
                raise ZstdError("decompression error: %s" % L.zb200_error_string(code.value).decode())
            seg0 = _native.Segment.from_address(L.zb200_result_segments(res))
            out = C.string_at(L.zb200_result_data(res) + seg0.offset, seg0.length) if seg0.length else b""
        finally:
            L.zb200_result_free(res)
        if not allow_extra_data:
            used = self._frame_size(buf)
            if used is not None and used < len(buf):
                raise ZstdError("compressed input contains %d bytes of unused data, which is disallowed"
                                % (len(buf) - used))
        return out

    @staticmethod
    def _frame_size(buf):
        """Compressed size of the first frame (header + block chain + checksum); host-side walk of
        the 3-byte block headers only (ZSTD_findFrameCompressedSize, zstd/zstd.c:43905)."""
        L = _native.lib()
        info = _native.FrameInfo()
        L.zb200_frame_info(buf.ctypes.data, len(buf), C.byref(info))
        if info.status:
            return None
        pos = info.header_size
        n = len(buf)
        while True:
            if pos + 3 > n:
                return None
            bh = int(buf[pos]) | (int(buf[pos + 1]) << 8) | (int(buf[pos + 2]) << 16)
            pos += 3
            btype = (bh >> 1) & 3
            pos += 1 if btype == 1 else (bh >> 3)
            if bh & 1:
                break
        return pos + (4 if info.has_checksum else 0)

    # ------------------------------------------------------------------ batch
    def multi_decompress_to_buffer(self, frames, decompressed_sizes=None, threads=0):
        sizes = None
        if decompressed_sizes is not None:
            sizes = np.frombuffer(memoryview(decompressed_sizes), dtype=np.uint8)
        L = _native.lib()

        if isinstance(frames, BufferWithSegments):
            sources = [frames]
        elif isinstance(frames, BufferWithSegmentsCollection):
            sources = frames._buffers
        elif isinstance(frames, list):
            sources = None
        else:
            raise TypeError("argument must be list of BufferWithSegments")

        if sources is not None:
            count = sum(len(b) for b in sources)
            self._check_sizes(sizes, count)
            if count == 0:
                raise ValueError("no source elements found")
            results = []
            first = 0
            for b in sources:
                n = len(b)
                sz = sizes[first * 8:(first + n) * 8] if sizes is not None else None
                results.extend(self._run_contiguous(b, n, sz, first, threads))
                first += n
            return BufferWithSegmentsCollection(*results)

        # list of bytes-like objects (c-ext/decompressor.c:1601-1685)
        count = len(frames)
        self._check_sizes(sizes, count)
        views = []
        for i, item in enumerate(frames):
            try:
                v = memoryview(item)
            except TypeError:
                raise TypeError("item %d not a bytes like object" % i)
            if not v.contiguous:
                raise TypeError("item %d not a bytes like object" % i)
            views.append(v)
        if count == 0:
            raise ValueError("no source elements found")
        return BufferWithSegmentsCollection(*self._run_list(views, sizes, threads))

    @staticmethod
    def _check_sizes(sizes, count):
        if sizes is not None and len(sizes) != count * 8:
            raise ValueError("decompressed_sizes size mismatch; expected %d, got %d" % (count * 8, len(sizes)))

    def _raise_item_error(self, L, res, base):
        item, code = C.c_size_t(), C.c_int()
        got, exp = C.c_uint64(), C.c_uint64()
        if not L.zb200_result_first_error(res, C.byref(item), C.byref(code), C.byref(got), C.byref(exp)):
            return
        L.zb200_result_free(res)
        idx = base + item.value
        if code.value == 200:
            raise ValueError("could not determine decompressed size of item %d" % idx)
        if code.value == 201:
            raise ZstdError("error decompressing item %d: decompressed %d bytes; expected %d"
                            % (idx, got.value, exp.value))
        raise ZstdError("error decompressing item %d: %s" % (idx, L.zb200_error_string(code.value).decode()))

    @staticmethod
    def _split(self, lengths, parts):
        from .sharding import split_ranges
        return split_ranges(lengths, parts)

    def _launch(self, ctx, base_ptr, segs, n, sizes_arr, flags=0):
        L = ctx.L
        res = C.c_void_p()
        dd = self._dict_data._ddict(ctx) if self._dict_data is not None else None
        with ctx.lock:
            rc = L.zb200_decompress_batch_ex(ctx.h, base_ptr, segs.ctypes.data, n,
                                             sizes_arr.ctypes.data if sizes_arr is not None else None, dd,
                                             C.byref(self._dparams), flags, C.byref(res))
        ctx.check(rc, "zb200_decompress_batch")
        return res

    # sub-batches in flight per device: while one copies its output to the host, the next runs its kernels
    # and a third uploads its input (PCIe is full duplex; the copies dominate the end-to-end time)
    PIPELINE_DEPTH = int(os.environ.get("ZB200_PIPELINE_DEPTH", "4"))
    SUB_BATCH_INPUT_BYTES = int(os.environ.get("ZB200_SUB_BATCH_MB", "24")) << 20

    def _run_contiguous(self, b, n, sizes_bytes, first_index, threads):
        if n == 0:
            return []
        L = _native.lib()
        segs = np.frombuffer(b._segments, dtype=np.uint64).reshape(-1, 2)
        lens = np.ascontiguousarray(segs[:, 1])
        data = np.frombuffer(b._data, dtype=np.uint8) if b.size else np.zeros(1, dtype=np.uint8)
        sizes_arr = np.frombuffer(sizes_bytes, dtype=np.uint64) if sizes_bytes is not None else None
        devs = _devices(threads)
        parts = self._split(None, lens, len(devs))
        depth = self.PIPELINE_DEPTH
        jobs = []           # (device, slot, lo, hi) in output order
        for di, (lo, hi) in enumerate(parts):
            dev = devs[di]
            nbytes = int(lens[lo:hi].sum())
            k = max(1, min((hi - lo) // 256 or 1, nbytes // self.SUB_BATCH_INPUT_BYTES))
            if k < 2:
                jobs.append((dev, 0, lo, hi))
                continue
            for i, (a, c) in enumerate(self._split(None, lens[lo:hi], k)):
                jobs.append((dev, i % depth, lo + a, lo + c))

        def run(job):
            # everything per sub-batch happens on the worker (launch, error lookup, wrapping the result) so that
            # it overlaps the other sub-batches' device work; errors are returned, the lowest item wins below
            dev, slot, lo, hi = job
            ctx = _native.Context.get(dev, slot)
            sub = np.ascontiguousarray(segs[lo:hi])
            ssz = np.ascontiguousarray(sizes_arr[lo:hi]) if sizes_arr is not None else None
            res = self._launch(ctx, data.ctypes.data, sub, hi - lo, ssz)
            try:
                self._raise_item_error(L, res, first_index + lo)
                return BufferWithSegments._from_result(ctx, res)
            except Exception as e:
                return e

        if len(jobs) == 1:
            handles = [run(jobs[0])]
        else:
            handles = list(_executor(depth * len(parts)).map(run, jobs))
        for h in handles:
            if isinstance(h, Exception):
                raise h
        return handles

    def _run_list(self, views, sizes, threads):
        L = _native.lib()
        n = len(views)
        lengths = np.array([v.nbytes for v in views], dtype=np.uint64)
        sizes_arr = np.frombuffer(sizes, dtype=np.uint64) if sizes is not None else None
        devs = _devices(threads)
        parts = self._split(None, lengths, len(devs))
        out = []
        for di, (lo, hi) in enumerate(parts):
            ctx = self._context(devs[di])
            k = hi - lo
            arrs = [np.frombuffer(v, dtype=np.uint8) if v.nbytes else np.zeros(0, dtype=np.uint8) for v in views[lo:hi]]
            ptrs = (C.c_void_p * k)(*[a.ctypes.data if len(a) else None for a in arrs])
            lens = (C.c_size_t * k)(*[len(a) for a in arrs])
            ssz = np.ascontiguousarray(sizes_arr[lo:hi]) if sizes_arr is not None else None
            res = C.c_void_p()
            dd = self._dict_data._ddict(ctx) if self._dict_data is not None else None
            with ctx.lock:
                rc = L.zb200_decompress_batch_ptrs_ex(ctx.h, ptrs, lens, k, ssz.ctypes.data if ssz is not None else None,
                                                      dd, C.byref(self._dparams), 0, C.byref(res))
            ctx.check(rc, "zb200_decompress_batch_ptrs")
            self._raise_item_error(L, res, lo)
            out.append(BufferWithSegments._from_result(ctx, res))
        return out

    # ------------------------------------------------------------------ out of scope
    def _unsupported(self, *a, **k):
        raise NotImplementedError("streaming decompression objects are outside the B200 batch path")

    stream_writer = read_to_iter = copy_stream = _unsupported
    decompress_content_dict_chain = _unsupported

    # ------------------------------------------------------------------ SURVEY.md section 8f rows 1 and 3
    def decompressobj(self, write_size=131072, read_across_frames=False):
        """c-ext/decompressor.c:397-455; a whole frame is the unit of device work (streams.py)."""
        from .streams import ZstdDecompressionObj
        return ZstdDecompressionObj(self, write_size, read_across_frames)

    def stream_reader(self, source, read_size=131075, read_across_frames=False, closefd=True):
        """c-ext/decompressor.c:509-566, c-ext/decompressionreader.c:177-318."""
        from .streams import ZstdDecompressionReader
        return ZstdDecompressionReader(self, source, read_size, read_across_frames, closefd)
