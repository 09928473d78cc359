"""The fifteen scenarios of the reference's own batch-API tests, restated against this package.

/root/reference/tests/test_compressor_multi_compress_to_buffer.py (6 tests) and
test_decompressor_multi_decompress_to_buffer.py (9 tests) exercise ZstdCompressor.multi_compress_to_buffer and
ZstdDecompressor.multi_decompress_to_buffer through the public API only.  /root/reference does not exist on the GPU
box, so the scenarios are written out here: the same inputs, the same calls, the same assertions (types, counts, sizes --
44 and 47 bytes for the two checksummed pairs --, error types and texts), with `python_zstandard_b200` where the
reference's tests say `zstandard`.  As there, frames are made by the package's own ZstdCompressor.
"""
import struct

import pytest

import python_zstandard_b200 as zstd

pytestmark = pytest.mark.gpu


def seg_table(lengths):
    out, pos = [], 0
    for n in lengths:
        out += [pos, n]
        pos += n
    return struct.pack("=" + "Q" * len(out), *out)


# ---------------------------------------------------------------- compressor (reference file :12-131)
def test_c_invalid_inputs():
    cctx = zstd.ZstdCompressor()
    with pytest.raises(TypeError):
        cctx.multi_compress_to_buffer(True)
    with pytest.raises(TypeError):
        cctx.multi_compress_to_buffer((1, 2))
    with pytest.raises(TypeError, match="item 0 not a bytes like object"):
        cctx.multi_compress_to_buffer(["foo"])


def test_c_empty_input():
    cctx = zstd.ZstdCompressor()
    with pytest.raises(ValueError, match="no source elements found"):
        cctx.multi_compress_to_buffer([])
    with pytest.raises(ValueError, match="source elements are empty"):
        cctx.multi_compress_to_buffer([b"", b"", b""])


def test_c_list_input():
    cctx = zstd.ZstdCompressor(write_checksum=True)
    original = [b"foo" * 12, b"bar" * 6]
    frames = [cctx.compress(c) for c in original]
    b = cctx.multi_compress_to_buffer(original)
    assert isinstance(b, zstd.BufferWithSegmentsCollection)
    assert len(b) == 2
    assert b.size() == 44
    assert b[0].tobytes() == frames[0]
    assert b[1].tobytes() == frames[1]


def test_c_buffer_with_segments_input():
    cctx = zstd.ZstdCompressor(write_checksum=True)
    original = [b"foo" * 4, b"bar" * 6]
    frames = [cctx.compress(c) for c in original]
    segments = zstd.BufferWithSegments(b"".join(original), seg_table(map(len, original)))
    result = cctx.multi_compress_to_buffer(segments)
    assert len(result) == 2
    assert result.size() == 47
    assert result[0].tobytes() == frames[0]
    assert result[1].tobytes() == frames[1]


def test_c_buffer_with_segments_collection_input():
    cctx = zstd.ZstdCompressor(write_checksum=True)
    original = [b"foo1", b"foo2" * 2, b"foo3" * 3, b"foo4" * 4, b"foo5" * 5]
    frames = [cctx.compress(c) for c in original]
    b1 = zstd.BufferWithSegments(b"".join(original[:2]), seg_table(map(len, original[:2])))
    b2 = zstd.BufferWithSegments(b"".join(original[2:]), seg_table(map(len, original[2:])))
    result = cctx.multi_compress_to_buffer(zstd.BufferWithSegmentsCollection(b1, b2))
    assert len(result) == len(frames)
    for i, frame in enumerate(frames):
        assert result[i].tobytes() == frame


def test_c_multiple_threads():
    refcctx = zstd.ZstdCompressor(write_checksum=True)
    reference = [refcctx.compress(b"x" * 64), refcctx.compress(b"y" * 64)]
    cctx = zstd.ZstdCompressor(write_checksum=True)
    frames = [b"x" * 64] * 256 + [b"y" * 64] * 256
    result = cctx.multi_compress_to_buffer(frames, threads=-1)
    assert len(result) == 512
    for i in range(512):
        assert result[i].tobytes() == reference[0 if i < 256 else 1]


# ---------------------------------------------------------------- decompressor (reference file :17-227)
def test_d_invalid_inputs():
    dctx = zstd.ZstdDecompressor()
    with pytest.raises(TypeError):
        dctx.multi_decompress_to_buffer(True)
    with pytest.raises(TypeError):
        dctx.multi_decompress_to_buffer((1, 2))
    with pytest.raises(TypeError, match="item 0 not a bytes like object"):
        dctx.multi_decompress_to_buffer(["foo"])
    with pytest.raises(ValueError, match="could not determine decompressed size of item 0"):
        dctx.multi_decompress_to_buffer([b"foobarbaz"])


def test_d_list_input():
    cctx = zstd.ZstdCompressor()
    original = [b"foo" * 4, b"bar" * 6]
    frames = [cctx.compress(d) for d in original]
    result = zstd.ZstdDecompressor().multi_decompress_to_buffer(frames)
    assert len(result) == len(frames)
    assert result.size() == sum(map(len, original))
    for i, data in enumerate(original):
        assert result[i].tobytes() == data
    assert result[0].offset == 0
    assert len(result[0]) == 12
    assert len(result[1]) == 18


def test_d_list_input_frame_sizes():
    cctx = zstd.ZstdCompressor()
    original = [b"foo" * 4, b"bar" * 6, b"baz" * 8]
    frames = [cctx.compress(d) for d in original]
    sizes = struct.pack("=" + "Q" * len(original), *map(len, original))
    result = zstd.ZstdDecompressor().multi_decompress_to_buffer(frames, decompressed_sizes=sizes)
    assert len(result) == len(frames)
    assert result.size() == sum(map(len, original))
    for i, data in enumerate(original):
        assert result[i].tobytes() == data


def test_d_buffer_with_segments_input():
    cctx = zstd.ZstdCompressor()
    original = [b"foo" * 4, b"bar" * 6]
    frames = [cctx.compress(d) for d in original]
    b = zstd.BufferWithSegments(b"".join(frames), seg_table(map(len, frames)))
    result = zstd.ZstdDecompressor().multi_decompress_to_buffer(b)
    assert len(result) == len(frames)
    assert result[0].offset == 0
    assert len(result[0]) == 12
    assert len(result[1]) == 18


def test_d_buffer_with_segments_sizes():
    cctx = zstd.ZstdCompressor(write_content_size=False)
    original = [b"foo" * 4, b"bar" * 6, b"baz" * 8]
    frames = [cctx.compress(d) for d in original]
    sizes = struct.pack("=" + "Q" * len(original), *map(len, original))
    b = zstd.BufferWithSegments(b"".join(frames), seg_table(map(len, frames)))
    result = zstd.ZstdDecompressor().multi_decompress_to_buffer(b, decompressed_sizes=sizes)
    assert len(result) == len(frames)
    assert result.size() == sum(map(len, original))
    for i, data in enumerate(original):
        assert result[i].tobytes() == data


def test_d_buffer_with_segments_collection_input():
    cctx = zstd.ZstdCompressor()
    original = [b"foo0" * 2, b"foo1" * 3, b"foo2" * 4, b"foo3" * 5, b"foo4" * 6]
    frames = cctx.multi_compress_to_buffer(original)
    decompressed = zstd.ZstdDecompressor().multi_decompress_to_buffer(frames, threads=3)       # round trip of the collection
    assert len(decompressed) == len(original)
    for i, data in enumerate(original):
        assert decompressed[i].tobytes() == data
    fb = [frames[i].tobytes() for i in range(5)]                                                  # and a manual mode
    b1 = zstd.BufferWithSegments(b"".join(fb[:2]), seg_table(map(len, fb[:2])))
    b2 = zstd.BufferWithSegments(b"".join(fb[2:]), seg_table(map(len, fb[2:])))
    decompressed = zstd.ZstdDecompressor().multi_decompress_to_buffer(zstd.BufferWithSegmentsCollection(b1, b2))
    assert len(decompressed) == 5
    for i in range(5):
        assert decompressed[i].tobytes() == original[i]


def test_d_dict():
    """The reference trains its dictionary with zstd.train_dictionary (out of scope here: SURVEY.md row 17); a raw-content
    dictionary made of the samples exercises the same path: compress with dict_data, batch-decompress with dict_data."""
    samples = []
    for i in range(128):        # tests/common.py:generate_samples of the reference
        samples.append(b"foo" * 64)
        samples.append(b"bar" * 64)
        samples.append(b"foobar" * 64)
        samples.append(b"baz" * 64)
        samples.append(b"foobaz" * 64)
        samples.append(b"bazfoo" * 64)
    d = zstd.ZstdCompressionDict(b"".join(samples[:12]), dict_type=zstd.DICT_TYPE_RAWCONTENT)
    cctx = zstd.ZstdCompressor(dict_data=d, level=1)
    frames = [cctx.compress(s) for s in samples]
    result = zstd.ZstdDecompressor(dict_data=d).multi_decompress_to_buffer(frames)
    assert [result[i].tobytes() for i in range(len(result))] == samples


def test_d_multiple_threads():
    cctx = zstd.ZstdCompressor()
    frames = [cctx.compress(b"x" * 64)] * 256 + [cctx.compress(b"y" * 64)] * 256
    result = zstd.ZstdDecompressor().multi_decompress_to_buffer(frames, threads=-1)
    assert len(result) == len(frames)
    assert result.size() == 2 * 64 * 256
    assert result[0].tobytes() == b"x" * 64
    assert result[256].tobytes() == b"y" * 64


def test_d_item_failure():
    cctx = zstd.ZstdCompressor()
    frames = [cctx.compress(b"x" * 128), cctx.compress(b"y" * 128)]
    frames[1] = frames[1][0:15] + b"extra" + frames[1][15:]
    pat = "error decompressing item 1: (Data corruption detected|Destination buffer is too small)"
    with pytest.raises(zstd.ZstdError, match=pat):
        zstd.ZstdDecompressor().multi_decompress_to_buffer(frames)
    with pytest.raises(zstd.ZstdError, match=pat):
        zstd.ZstdDecompressor().multi_decompress_to_buffer(frames, threads=2)
