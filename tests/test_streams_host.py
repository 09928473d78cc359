"""Host logic of the streaming-shaped objects (python_zstandard_b200/streams.py) that needs no device: the frame walker that
finds the end of a frame in a growing buffer (what the reference's ZSTD_decompressStream discovers block by block,
zstd/zstd.c:45307-45560) against ZSTD_findFrameCompressedSize of the unmodified reference."""
import ctypes as C

import numpy as np
import pytest

import corpus
from oracle import RefZstd, have_ref
from python_zstandard_b200 import streams
from python_zstandard_b200.errors import ZstdError

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref is built from /root/reference")


def test_frame_walker_finds_the_frame_end_in_a_growing_buffer():
    ref = RefZstd()
    Z = ref.Z
    Z.ZSTD_findFrameCompressedSize.restype = C.c_size_t
    Z.ZSTD_findFrameCompressedSize.argtypes = [C.c_char_p, C.c_size_t]
    rng = np.random.default_rng(2)
    text = corpus.text_corpus(1 << 20)
    cases = [b"", b"x", bytes(1000), text[:70000].tobytes(), text[1000:1000 + 400000].tobytes(),
             rng.integers(0, 256, 200000).astype(np.uint8).tobytes()]
    for data in cases:
        for checksum in (False, True):
            for content_size in (False, True):
                frame = ref.compress(data, level=3, checksum=checksum, content_size=content_size)
                want = Z.ZSTD_findFrameCompressedSize(frame, len(frame))
                assert want == len(frame)
                trailing = frame + b"\x28\xb5\x2f\xfd" + bytes(9)              # the next frame's first bytes follow
                for step in (1, 7, 4096, 131075):
                    w = streams._FrameWalker(); buf = bytearray(); end = None; pos = 0
                    while end is None and pos < len(trailing):
                        buf += trailing[pos:pos + step]; pos += step
                        end = w.feed(buf)
                        if step == 1 and len(buf) > 64 and end is None:
                            step = 997                                        # (byte-wise only through the header)
                    assert end == want, (len(data), checksum, content_size, step)
                    assert w.content_size == (len(data) if content_size else None)
                # a truncated frame never reports an end
                w = streams._FrameWalker()
                assert w.feed(bytearray(frame[:-1])) is None
    with pytest.raises(ZstdError):
        streams._FrameWalker().feed(bytearray(b"not a zstd frame at all....."))


class _HostFrame:
    """Stands in for decompressor._FrameOutput (the device result) in the tests below."""

    def __init__(self, data):
        self._d = data
        self.closed = False

    def __len__(self):
        return len(self._d)

    def tobytes(self, start=0, stop=None):
        return self._d[start:len(self._d) if stop is None else stop]

    def copy_into(self, mv, start, n):
        mv[:n] = self._d[start:start + n]

    def close(self):
        self.closed = True


class _HostDctx:
    """The two calls the stream objects make on a ZstdDecompressor, answered by the reference codec: the objects' own logic
    (buffering, frame boundaries, reads across frames, tell / eof / unused_data) runs without a device."""

    def __init__(self, ref):
        self.ref = ref
        self.frames = []

    def _decompress_frame(self, frame, max_output_size=0, allow_extra_data=True):
        f = _HostFrame(self.ref.decompress(bytes(frame), max_output_size or (1 << 24)))
        self.frames.append(f)
        return f


def test_reader_and_decompressobj_logic_over_several_frames():
    import io
    ref = RefZstd()
    text = corpus.text_corpus(1 << 20)
    parts = [text[:50000].tobytes(), b"", text[60000:60000 + 300000].tobytes(), b"tail"]
    blob = b"".join(ref.compress(p, level=3, content_size=(i != 2)) for i, p in enumerate(parts))
    whole = b"".join(parts)
    # stream_reader across frames, odd read sizes, from a file object and from bytes
    for source in (io.BytesIO(blob), blob):
        dctx = _HostDctx(ref)
        r = streams.ZstdDecompressionReader(dctx, source, read_size=4097, read_across_frames=True)
        got = []
        while True:
            c = r.read(33333)
            if not c:
                break
            got.append(c)
        assert b"".join(got) == whole and r.tell() == len(whole)
        assert all(f.closed for f in dctx.frames[:-1])                        # finished frames give their result back
        r.close()
        assert dctx.frames[-1].closed
    # one frame only (the default): the reader stops at its end
    r = streams.ZstdDecompressionReader(_HostDctx(ref), io.BytesIO(blob))
    assert r.readall() == parts[0] and r.read(10) == b""
    # readinto / read1 / context manager
    with streams.ZstdDecompressionReader(_HostDctx(ref), blob, read_across_frames=True) as r:
        buf = bytearray(70000)
        assert r.readinto(buf) == 70000 and bytes(buf) == whole[:70000]
        assert r.read1(10) == whole[70000:70010]
        assert r.readall() == whole[70010:]
    with pytest.raises(ValueError):
        r.read(1)                                                            # closed
    # input that ends inside a frame
    with pytest.raises(ZstdError, match="Src size is incorrect"):
        streams.ZstdDecompressionReader(_HostDctx(ref), io.BytesIO(blob[:len(blob) - 3]), read_across_frames=True).readall()
    # decompressobj: fed in small pieces; one frame, then unused_data
    o = streams.ZstdDecompressionObj(_HostDctx(ref))
    out, i = b"", 0
    while not o.eof:
        out += o.decompress(blob[i:i + 5000]); i += 5000
    assert out == parts[0] and o.unused_data == blob[len(ref.compress(parts[0], level=3)):i]
    with pytest.raises(ZstdError, match="cannot use a decompressobj multiple times"):
        o.decompress(b"x")
    o = streams.ZstdDecompressionObj(_HostDctx(ref), read_across_frames=True)
    assert b"".join(o.decompress(blob[i:i + 7777]) for i in range(0, len(blob), 7777)) == whole


def test_compressobj_logic():
    ref = RefZstd()

    class Cctx:
        def compress(self, data):
            return ref.compress(data, level=3)
    text = corpus.text_corpus(1 << 20)[:200000].tobytes()
    o = streams.ZstdCompressionObj(Cctx())
    assert [o.compress(text[i:i + 30000]) for i in range(0, len(text), 30000)] == [b""] * 7
    frame = o.flush()
    assert ref.decompress(frame, len(text)) == text
    with pytest.raises(ZstdError, match="cannot call compress\\(\\) after compressor finished"):
        o.compress(b"more")
    with pytest.raises(ZstdError, match="compressor object already finished"):
        o.flush()
    o = streams.ZstdCompressionObj(Cctx(), size=10)
    o.compress(b"12345")
    with pytest.raises(ZstdError, match="Src size is incorrect"):
        o.flush()
    with pytest.raises(ValueError, match="flush mode not recognized"):
        streams.ZstdCompressionObj(Cctx()).flush(flush_mode=7)
    with pytest.raises(NotImplementedError):
        streams.ZstdCompressionObj(Cctx()).flush(flush_mode=streams.COMPRESSOBJ_FLUSH_BLOCK)
