"""stream_reader / decompressobj / compressobj over the GPU one-shot paths: same bytes
as the reference's objects produce (scenarios of the reference's tests/test_decompressor_stream_reader.py,
test_decompressor_decompressobj.py and test_compressor_compressobj.py that do not depend on chunk timing)."""
import io

import numpy as np
import pytest

import corpus
import python_zstandard_b200 as zstd
from oracle import RefZstd, have_ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_ref(), reason="oracle/_ref is built from /root/reference")]


@pytest.fixture(scope="module")
def ref():
    return RefZstd()


def test_stream_reader_large_frame_of_128k_blocks(ref):
    """BASELINE config 5 in small: one frame of many 128 KiB blocks (cross-block matches, repeat tables, treeless literals),
    read through stream_reader in odd sizes; also without a content size in the header."""
    data = corpus.text_corpus(8 << 20).tobytes()[:6 * 1024 * 1024 + 12345]
    for content_size in (True, False):
        frame = ref.compress(data, level=3, content_size=content_size, checksum=True)
        with zstd.ZstdDecompressor(max_window_size=1 << 27).stream_reader(io.BytesIO(frame), read_size=70001) as r:
            parts = []
            while True:
                chunk = r.read(100003)
                if not chunk:
                    break
                parts.append(chunk)
            assert r.tell() == len(data)
        assert b"".join(parts) == data
    r = zstd.ZstdDecompressor().stream_reader(frame)                   # bytes source, readall, readinto
    assert r.readall() == data
    r = zstd.ZstdDecompressor().stream_reader(frame)
    buf = bytearray(1000)
    assert r.readinto(buf) == 1000 and bytes(buf) == data[:1000]
    assert r.read1(50) == data[1000:1050]
    with pytest.raises(io.UnsupportedOperation):
        r.write(b"x")


def test_stream_reader_frames_and_errors(ref):
    a, b = b"foo" * 1000, corpus.text_corpus(1 << 20).tobytes()[:70000]
    both = ref.compress(a) + ref.compress(b)
    assert zstd.ZstdDecompressor().stream_reader(both).read(-1) == a                       # first frame only
    assert zstd.ZstdDecompressor().stream_reader(both, read_across_frames=True).readall() == a + b
    with pytest.raises(zstd.ZstdError, match="Src size is incorrect"):
        zstd.ZstdDecompressor().stream_reader(ref.compress(b)[:-5]).readall()
    r = zstd.ZstdDecompressor().stream_reader(both)
    r.close()
    with pytest.raises(ValueError, match="stream is closed"):
        r.read(1)


def test_decompressobj(ref):
    data = corpus.text_corpus(1 << 20).tobytes()[:300000]
    frame = ref.compress(data)
    d = zstd.ZstdDecompressor().decompressobj()
    out = b"".join(d.decompress(frame[i:i + 8191]) for i in range(0, len(frame), 8191))
    assert out == data and d.eof and d.unused_data == b"" and d.flush() == b""
    with pytest.raises(zstd.ZstdError, match="cannot use a decompressobj multiple times"):
        d.decompress(b"x")
    d = zstd.ZstdDecompressor().decompressobj()
    assert d.decompress(frame + b"trailing") == data and d.unused_data == b"trailing"
    d = zstd.ZstdDecompressor().decompressobj(read_across_frames=True)
    assert d.decompress(frame + ref.compress(b"second")) == data + b"second" and not d.eof


def test_compressobj(ref):
    data = corpus.text_corpus(1 << 20).tobytes()[:200000]
    c = zstd.ZstdCompressor(level=3, write_checksum=True).compressobj()
    assert c.compress(data[:1000]) == b"" and c.compress(data[1000:]) == b""
    frame = c.flush()
    assert ref.decompress(frame, len(data)) == data
    with pytest.raises(zstd.ZstdError, match="cannot call compress\\(\\) after compressor finished"):
        c.compress(b"x")
    with pytest.raises(zstd.ZstdError, match="Src size is incorrect"):
        c2 = zstd.ZstdCompressor().compressobj(size=5)
        c2.compress(b"abc")
        c2.flush()
