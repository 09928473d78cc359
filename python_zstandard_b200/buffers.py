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


class DeviceBufferSegment:
    """One item of a DeviceBufferWithSegments: a view of device memory (``__cuda_array_interface__``), ``tobytes()`` copies
    it to the host."""

    __slots__ = ("_parent", "_offset", "_length")

    def __init__(self, parent, offset, length):
        self._parent, self._offset, self._length = parent, offset, length

    @property
    def offset(self):
        return self._offset

    def __len__(self):
        return self._length

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self._length,), "typestr": "|u1", "data": (self._parent._ptr + self._offset, False), "version": 3}

    def tobytes(self):
        return self._parent._download(self._offset, self._length)


class DeviceBufferWithSegments:
    """BufferWithSegments whose bytes stay in DEVICE memory: the batch calls take it without a
    host->device copy and -- given one -- return one, so the user-visible rate is the kernels', not PCIe's.

    ``data``: any object with ``__cuda_array_interface__`` (a torch CUDA tensor, a cupy array, another of these), 1-D bytes;
    it must be ready (no kernel of another stream still writing it) when a batch call reads it.  ``segments``: the same
    ``{u64 offset; u64 length}`` table as BufferWithSegments takes (host memory).  The object exposes
    ``__cuda_array_interface__`` itself: ``torch.as_tensor(buf, device="cuda")`` is a zero-copy view."""

    def __init__(self, data, segments):
        from . import _native
        cai = getattr(data, "__cuda_array_interface__", None)
        if cai is None:
            raise TypeError("data must expose __cuda_array_interface__ (a CUDA tensor or array)")
        shape = cai["shape"]
        if len(shape) != 1 or cai.get("strides") not in (None, (np_itemsize(cai["typestr"]),)):
            raise TypeError("data must be a contiguous 1-D device array")
        size = int(shape[0]) * np_itemsize(cai["typestr"])
        ptr = int(cai["data"][0]) if size else 0
        seg = bytes(memoryview(segments))
        if len(seg) % SEGMENT_SIZE:
            raise ValueError("segments array size is not a multiple of %d" % SEGMENT_SIZE)
        n = len(seg) // SEGMENT_SIZE
        for i in range(n):
            off, length = _SEG.unpack_from(seg, i * SEGMENT_SIZE)
            if off + length > size:
                raise ValueError("offset within segments array references memory outside buffer")
        dev = _native.lib().zb200_pointer_device(ptr) if size else 0
        if dev < 0:
            raise TypeError("data does not live in device memory")
        self._keep = data            # keeps the caller's allocation alive
        self._ptr, self._size, self._segments, self._count, self._device = ptr, size, seg, n, dev
        self._owner = None

    @classmethod
    def _from_result(cls, ctx, handle):
        """A device-resident result of libzb200 (zb200_result with ZB200_DST_DEVICE): the result owns its allocation."""
        L = ctx.L
        self = cls.__new__(cls)
        n = L.zb200_result_count(handle)
        self._keep = None
        self._ptr = L.zb200_result_data(handle) or 0
        self._size = L.zb200_result_size(handle)
        self._segments = C.string_at(L.zb200_result_segments(handle), n * SEGMENT_SIZE) if n else b""
        self._count = n
        self._device = ctx.device
        self._ctx = ctx
        self._owner = weakref.finalize(self, L.zb200_result_free, handle)
        return self

    @property
    def device(self):
        return self._device

    @property
    def size(self):
        return self._size

    def __len__(self):
        return self._count

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self._size,), "typestr": "|u1", "data": (self._ptr, False), "version": 3}

    def segments(self):
        return BufferSegments(self, self._segments)

    def __getitem__(self, i):
        if i < 0:
            raise IndexError("offset must be non-negative")
        if i >= self._count:
            raise IndexError("offset must be less than %d" % self._count)
        off, length = _SEG.unpack_from(self._segments, i * SEGMENT_SIZE)
        return DeviceBufferSegment(self, off, length)

    def _download(self, offset, length):
        from . import _native
        if not length:
            return b""
        ctx = _native.Context.get(self._device)
        out = bytearray(length)
        dst = (C.c_ubyte * length).from_buffer(out)
        with ctx.lock:
            rc = ctx.L.zb200_memcpy_d2h(ctx.h, C.addressof(dst), self._ptr + offset, length)
        del dst
        ctx.check(rc, "zb200_memcpy_d2h")
        return bytes(out)

    def tobytes(self):
        return self._download(0, self._size)

    def to_host(self):
        """The same items as an ordinary BufferWithSegments (one device->host copy)."""
        return BufferWithSegments(self.tobytes(), self._segments)


def np_itemsize(typestr):
    return int(typestr[2:]) if len(typestr) > 2 else 1


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
