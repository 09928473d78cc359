"""Buffer types of the batch API.

Mirrors c-ext/bufferutil.c of the reference (BufferWithSegments :39-245,
BufferSegments :247-300, BufferSegment :302-370, BufferWithSegmentsCollection :372-520):
same constructor arguments, same validation, same error texts.  A segment table is
an array of native-endian ``{u64 offset; u64 length}`` (c-ext/python-zstandard.h:307-313),
which is also exactly ``zb200_segment`` of the C ABI, so tables go to the device as they are.
"""
import ctypes as C
import struct
import weakref

_SEG = struct.Struct("=QQ")
SEGMENT_SIZE = _SEG.size   # 16


class BufferSegment:
    """One segment of a BufferWithSegments; keeps the parent alive (bufferutil.c:181-186)."""

    __slots__ = ("_parent", "_view", "_offset")

    def __init__(self, parent, view, offset):
        self._parent = parent
        self._view = view
        self._offset = offset

    @property
    def offset(self):
        return self._offset

    def __len__(self):
        return len(self._view)

    def tobytes(self):
        return self._view.tobytes()

    def __buffer__(self, flags):
        return memoryview(self._view)

    def __release_buffer__(self, view):
        pass


class BufferSegments:
    """Read-only view of a segment table (bufferutil.c:247-300)."""

    __slots__ = ("_parent", "_raw")

    def __init__(self, parent, raw):
        self._parent = parent
        self._raw = raw

    def __buffer__(self, flags):
        return memoryview(self._raw)

    def __release_buffer__(self, view):
        pass


class BufferWithSegments:
    """A contiguous buffer plus the (offset, length) table of the items inside it."""

    def __init__(self, data, segments):
        view = memoryview(data)
        if view.ndim != 1 or not view.contiguous:
            raise TypeError("data must be a contiguous buffer")
        view = view.cast("B") if view.format != "B" else view
        seg = bytes(memoryview(segments))          # copied, like the reference (bufferutil.c:78-88)
        if len(seg) % SEGMENT_SIZE:
            raise ValueError("segments array size is not a multiple of %d" % SEGMENT_SIZE)
        n = len(seg) // SEGMENT_SIZE
        size = len(view)
        for i in range(n):
            off, length = _SEG.unpack_from(seg, i * SEGMENT_SIZE)
            if off + length > size:
                raise ValueError("offset within segments array references memory outside buffer")
        self._data = view
        self._segments = seg
        self._count = n
        self._owner = None
        self._ptr = None           # raw address when the memory belongs to libzb200 (pinned host memory)

    # -- construction from shim-owned memory (BufferWithSegments_FromMemory, bufferutil.c:107-148)
    @classmethod
    def _from_result(cls, ctx, handle):
        L = ctx.L
        size = L.zb200_result_size(handle)
        n = L.zb200_result_count(handle)
        ptr = L.zb200_result_data(handle)
        segp = L.zb200_result_segments(handle)
        self = cls.__new__(cls)
        arr = (C.c_ubyte * size).from_address(ptr) if size else (C.c_ubyte * 0)()
        self._data = memoryview(arr).cast("B")
        self._segments = C.string_at(segp, n * SEGMENT_SIZE) if n else b""
        self._count = n
        self._ptr = ptr
        self._owner = weakref.finalize(self, L.zb200_result_free, handle)
        return self

    @property
    def size(self):
        return len(self._data)

    def __len__(self):
        return self._count

    def _segment(self, i):
        return _SEG.unpack_from(self._segments, i * SEGMENT_SIZE)

    def __getitem__(self, i):
        if i < 0:
            raise IndexError("offset must be non-negative")
        if i >= self._count:
            raise IndexError("offset must be less than %d" % self._count)
        off, length = self._segment(i)
        return BufferSegment(self, self._data[off:off + length], off)

    def segments(self):
        return BufferSegments(self, self._segments)

    def tobytes(self):
        return self._data.tobytes()

    def __buffer__(self, flags):
        return memoryview(self._data)

    def __release_buffer__(self, view):
        pass


class BufferWithSegmentsCollection:
    """Several BufferWithSegments addressed as one flat sequence (bufferutil.c:372-520)."""

    def __init__(self, *args):
        if not args:
            raise ValueError("must pass at least 1 argument")
        for a in args:
            if not isinstance(a, BufferWithSegments):
                raise TypeError("arguments must be BufferWithSegments instances")
            if len(a) == 0 or a.size == 0:
                raise ValueError("ZstdBufferWithSegments cannot be empty")
        self._buffers = list(args)
        self._first = []
        total = 0
        for a in args:
            total += len(a)
            self._first.append(total)      # running end index, like firstElements (bufferutil.c:420-428)
        self._count = total

    def __len__(self):
        return self._count

    def __getitem__(self, i):
        if i < 0:
            raise IndexError("offset must be non-negative")
        if i >= self._count:
            raise IndexError("offset must be less than %d" % self._count)
        start = 0
        for buf, end in zip(self._buffers, self._first):
            if i < end:
                return buf[i - start]
            start = end
        raise IndexError("offset must be less than %d" % self._count)   # pragma: no cover

    def size(self):
        total = 0
        for buf in self._buffers:
            for i in range(len(buf)):
                total += buf._segment(i)[1]
        return total
