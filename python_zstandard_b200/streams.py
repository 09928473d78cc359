"""Streaming-shaped objects over the one-shot GPU paths.

The reference drives ZSTD_decompressStream / ZSTD_compressStream2 chunk by chunk
(c-ext/decompressionreader.c:177-318, c-ext/decompressobj.c:28-150, c-ext/compressobj.c:21-210).  Here a whole frame
is the unit of device work: the objects below buffer input until a frame is complete (its end is found by walking the
3-byte block headers on the host, zstd/zstd.c:43905 ZSTD_findFrameCompressedSize), hand the frame to the GPU decoder /
encoder, and serve the result in the sizes the caller asks for.  Same observable results as the reference's objects
(bytes, unused_data, eof, tell()), different timing of when output becomes available.
"""
import io

import numpy as np

from .errors import ZstdError

COMPRESSOBJ_FLUSH_FINISH = 0
COMPRESSOBJ_FLUSH_BLOCK = 1
DECOMPRESSION_RECOMMENDED_INPUT_SIZE = 131075
DECOMPRESSION_RECOMMENDED_OUTPUT_SIZE = 131072
COMPRESSION_RECOMMENDED_INPUT_SIZE = 131072
COMPRESSION_RECOMMENDED_OUTPUT_SIZE = 131591


class _FrameWalker:
    """Incremental search for the end of the zstd frame at the start of a growing buffer."""

    def __init__(self):
        self.pos = None          # next block header (None until the frame header is complete)
        self.blocks = 0
        self.has_checksum = False
        self.content_size = None
        self.end = None          # frame size once known

    def feed(self, buf):
        """buf: bytes-like holding the frame from its first byte.  Returns the frame's size or None (need more)."""
        from . import _native
        import ctypes as C
        if self.end is not None:
            return self.end if len(buf) >= self.end else None
        n = len(buf)
        if self.pos is None:
            if n < 5:
                return None
            head = bytes(buf[:18])
            info = _native.FrameInfo()
            _native.lib().zb200_frame_info(head, len(head), C.byref(info))
            if info.status == 72 and n < 18:       # header not complete yet
                return None
            if info.status:
                raise ZstdError("zstd decompress error: %s" % _native.lib().zb200_error_string(info.status).decode())
            self.pos = info.header_size
            self.has_checksum = bool(info.has_checksum)
            self.content_size = None if info.content_size == (1 << 64) - 1 else info.content_size
        while True:
            if self.pos + 3 > n:
                return None
            bh = buf[self.pos] | (buf[self.pos + 1] << 8) | (buf[self.pos + 2] << 16)
            btype = (bh >> 1) & 3
            size = 1 if btype == 1 else (bh >> 3)
            nxt = self.pos + 3 + size
            if nxt > n:
                return None
            self.pos = nxt
            self.blocks += 1
            if bh & 1:
                self.end = self.pos + (4 if self.has_checksum else 0)
                return self.end if n >= self.end else None


def _decode_frame(dctx, frame, walker):
    """One complete frame on the GPU -> its output, still in the native result (frames without a content size get their
    block count x 128 KiB as capacity)."""
    if walker.content_size is not None:
        return dctx._decompress_frame(frame)
    return dctx._decompress_frame(frame, max_output_size=max(1, walker.blocks) * 131072)


def _take_frame(dctx, buf, end, walker):
    """Decode buf[:end] (a bytearray) without copying it first, then drop it from the buffer."""
    mv = memoryview(buf)[:end]
    try:
        out = _decode_frame(dctx, mv, walker)
    finally:
        try:
            mv.release()
        except BufferError:          # (an exception's traceback still holds the view)
            pass
    del buf[:end]
    return out


class ZstdDecompressionObj:
    """ZstdDecompressor.decompressobj() (c-ext/decompressobj.c)."""

    def __init__(self, dctx, write_size=DECOMPRESSION_RECOMMENDED_OUTPUT_SIZE, read_across_frames=False):
        if write_size < 1:
            raise ValueError("write_size must be positive")
        self._dctx = dctx
        self._across = read_across_frames
        self._buf = bytearray()
        self._walker = _FrameWalker()
        self._finished = False
        self._unused = b""

    def decompress(self, data):
        if self._finished:
            raise ZstdError("cannot use a decompressobj multiple times")
        self._buf += bytes(memoryview(data))
        out = []
        while self._buf:
            end = self._walker.feed(self._buf)
            if end is None:
                break
            fr = _take_frame(self._dctx, self._buf, end, self._walker)
            out.append(fr.tobytes())
            fr.close()
            self._walker = _FrameWalker()
            if not self._across:
                self._finished = True
                self._unused = bytes(self._buf)
                self._buf = bytearray()
                break
        return b"".join(out)

    def flush(self, length=0):
        return b""

    @property
    def unused_data(self):
        return self._unused

    @property
    def unconsumed_tail(self):
        return b""

    @property
    def eof(self):
        return self._finished


class ZstdDecompressionReader(io.RawIOBase):
    """ZstdDecompressor.stream_reader(source) (c-ext/decompressionreader.c): read()/readinto()/read1()/readall()/tell()."""

    def __init__(self, dctx, source, read_size=DECOMPRESSION_RECOMMENDED_INPUT_SIZE, read_across_frames=False, closefd=True):
        super().__init__()
        self._dctx = dctx
        self._source = source
        self._read_size = max(1, read_size)
        self._across = read_across_frames
        self._closefd = closefd
        self._in = bytearray()
        self._src_done = False
        self._src_view = None
        self._src_pos = 0
        if not hasattr(source, "read"):
            self._src_view = memoryview(source).cast("B")
        self._walker = _FrameWalker()
        self._out = None             # the current frame's output (_FrameOutput), served from where the device copied it
        self._out_len = 0
        self._out_pos = 0
        self._finished = False
        self._returned = 0
        self._entered = False

    def __enter__(self):
        if self._entered:
            raise ValueError("cannot __enter__ multiple times")
        if self.closed:
            raise ValueError("stream is closed")
        self._entered = True
        return self

    def __exit__(self, *a):
        self._entered = False
        self.close()
        return False

    def readable(self):
        return True

    def writable(self):
        return False

    def seekable(self):
        return False

    def write(self, data):
        raise io.UnsupportedOperation()

    def tell(self):
        return self._returned

    def close(self):
        if self.closed:
            return
        super().close()
        self._drop_frame()
        if self._closefd and hasattr(self._source, "close"):
            self._source.close()

    def _drop_frame(self):
        if self._out is not None:
            self._out.close()
        self._out, self._out_len, self._out_pos = None, 0, 0

    def _more_input(self):
        if self._src_done:
            return False
        if self._src_view is not None:
            chunk = self._src_view[self._src_pos:self._src_pos + max(self._read_size, 1 << 20)]
            self._src_pos += len(chunk)
        else:
            chunk = self._source.read(self._read_size)
        if not chunk:
            self._src_done = True
            return False
        self._in += chunk
        return True

    def _next_frame(self):
        """Decode the next complete frame into self._out.  False at the end of the input."""
        if self._finished:
            return False
        while True:
            end = self._walker.feed(self._in) if self._in else None
            if end is not None:
                break
            if not self._more_input():
                if self._in:
                    raise ZstdError("zstd decompress error: Src size is incorrect")      # input ends inside a frame
                self._finished = True
                return False
        self._drop_frame()
        self._out = _take_frame(self._dctx, self._in, end, self._walker)
        self._out_len = len(self._out)
        self._walker = _FrameWalker()
        if not self._across:
            self._finished = True
        return True

    def readinto(self, b):
        if self.closed:
            raise ValueError("stream is closed")
        mv = memoryview(b).cast("B")
        got = 0
        while got < len(mv):
            if self._out_pos >= self._out_len:
                if not self._next_frame():
                    break
                continue
            k = min(len(mv) - got, self._out_len - self._out_pos)
            self._out.copy_into(mv[got:got + k], self._out_pos, k)
            self._out_pos += k
            got += k
        self._returned += got
        return got

    def read(self, size=-1):
        if self.closed:
            raise ValueError("stream is closed")
        if size < -1:
            raise ValueError("cannot read negative amounts less than -1")
        if size == -1:
            return self.readall()
        if self._out_pos >= self._out_len and not self._next_frame():
            return b""
        if self._out_len - self._out_pos >= size:           # the common case: one copy, pinned result -> bytes
            data = self._out.tobytes(self._out_pos, self._out_pos + size)
            self._out_pos += size
            self._returned += size
            return data
        buf = bytearray(size)
        n = self.readinto(buf)
        return bytes(buf[:n])

    def read1(self, size=-1):
        if self.closed:
            raise ValueError("stream is closed")
        if self._out_pos >= self._out_len and not self._next_frame():
            return b""
        avail = self._out_len - self._out_pos
        k = avail if size < 0 else min(size, avail)
        data = self._out.tobytes(self._out_pos, self._out_pos + k)
        self._out_pos += k
        self._returned += k
        return data

    def readinto1(self, b):
        data = self.read1(len(memoryview(b).cast("B")))
        memoryview(b).cast("B")[:len(data)] = data
        return len(data)

    def readall(self):
        if self.closed:
            raise ValueError("stream is closed")
        parts = []
        while True:
            if self._out_pos < self._out_len:
                parts.append(self._out.tobytes(self._out_pos))
                self._returned += self._out_len - self._out_pos
                self._out_pos = self._out_len
            if not self._next_frame():
                break
        return b"".join(parts)

    def __iter__(self):
        raise io.UnsupportedOperation()

    def __next__(self):
        raise io.UnsupportedOperation()

    next = __next__

    def readline(self, size=-1):
        raise io.UnsupportedOperation()

    def readlines(self, hint=-1):
        raise io.UnsupportedOperation()


class ZstdCompressionObj:
    """ZstdCompressor.compressobj() (c-ext/compressobj.c): compress() buffers, flush() writes the frame."""

    def __init__(self, cctx, size=-1):
        self._cctx = cctx
        self._size = size
        self._parts = []
        self._finished = False

    def compress(self, data):
        if self._finished:
            raise ZstdError("cannot call compress() after compressor finished")
        self._parts.append(bytes(memoryview(data)))
        return b""

    def flush(self, flush_mode=COMPRESSOBJ_FLUSH_FINISH):
        if flush_mode not in (COMPRESSOBJ_FLUSH_FINISH, COMPRESSOBJ_FLUSH_BLOCK):
            raise ValueError("flush mode not recognized")
        if self._finished:
            raise ZstdError("compressor object already finished")
        if flush_mode == COMPRESSOBJ_FLUSH_BLOCK:
            raise NotImplementedError("COMPRESSOBJ_FLUSH_BLOCK: the B200 backend writes a frame in one piece")
        self._finished = True
        data = b"".join(self._parts)
        self._parts = []
        if self._size >= 0 and self._size != len(data):
            raise ZstdError("error ending compression stream: Src size is incorrect")
        return self._cctx.compress(data)
