"""Compression parity: frames made by the CUDA compressor must decode bit-exact through the
reference's own decoder (oracle/_ref) and the plain-C oracle, and stay within a stated size
margin of the reference's level 3 (the parse is ours, so the bytes differ by design)."""
import struct

import numpy as np
import pytest

import corpus
import python_zstandard_b200 as zstd
from oracle import Oracle, RefZstd, have_ref

pytestmark = pytest.mark.gpu

# stated size margins vs the reference at level 3 (single-block inputs; measured: +0.3% / +1.1% / +4.2%)
MARGIN_4K_TEXT = 1.02
MARGIN_128K_MIX = 1.03
MARGIN_128K_TEXT = 1.06


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.fixture(scope="module")
def ref():
    if not have_ref():
        pytest.skip("oracle/_ref not present")
    return RefZstd()


def _cases():
    text = corpus.text_corpus().tobytes()
    rng = np.random.default_rng(5)
    return {
        "empty": b"", "one": b"a", "foo12": b"foo" * 12, "x64": b"x" * 64, "text100": text[:100], "text1k": text[:1000],
        "text4k": text[5000:9096], "text64k": text[:65536], "text128k": text[:131072], "text300k": text[:300000],
        "rand5k": rng.integers(0, 256, 5000, dtype=np.uint8).tobytes(), "zeros100k": b"\0" * 100000,
        "abcd": bytes(rng.choice(list(b"abcd"), 20000).astype(np.uint8)), "bin": corpus.binary_blob(100000).tobytes(),
        "two_symbols": bytes(rng.choice(list(b"ab"), 3000).astype(np.uint8)), "block_edge": text[:131072 + 1],
    }


def test_one_shot_roundtrip_through_oracle(oracle):
    for ck in (False, True):
        for cs in (True, False):
            c = zstd.ZstdCompressor(write_checksum=ck, write_content_size=cs)
            for name, data in _cases().items():
                frame = c.compress(data)
                assert oracle.decompress(frame, len(data)) == data, (name, ck, cs)
                if cs:
                    assert zstd.frame_content_size(frame) == len(data)
                else:
                    assert zstd.frame_content_size(frame) == -1


def test_one_shot_roundtrip_through_reference_decoder(ref):
    c = zstd.ZstdCompressor(write_checksum=True)
    for name, data in _cases().items():
        frame = c.compress(data)
        assert ref.decompress(frame, len(data)) == data, name


def test_tiny_inputs_match_reference_sizes(ref):
    # the reference's exact-size expectations (tests/test_compressor_multi_compress_to_buffer.py:45,64):
    # 'foo'*12 + 'bar'*6 with checksums -> 44 bytes, 'foo'*4 + 'bar'*6 -> 47 bytes
    cctx = zstd.ZstdCompressor(write_checksum=True)
    b = cctx.multi_compress_to_buffer([b"foo" * 12, b"bar" * 6])
    assert isinstance(b, zstd.BufferWithSegmentsCollection)
    assert len(b) == 2 and b.size() == 44
    assert b[0].tobytes() == cctx.compress(b"foo" * 12) and b[1].tobytes() == cctx.compress(b"bar" * 6)
    offsets = struct.pack("=QQQQ", 0, 12, 12, 18)
    r = cctx.multi_compress_to_buffer(zstd.BufferWithSegments(b"foo" * 4 + b"bar" * 6, offsets))
    assert len(r) == 2 and r.size() == 47
    # empty-frame known answers (tests/test_compressor_compress.py:19,28)
    assert zstd.ZstdCompressor(write_content_size=False).compress(b"") == bytes.fromhex("28b52ffd0000010000")
    assert zstd.ZstdCompressor().compress(b"") == bytes.fromhex("28b52ffd2000010000")
    assert zstd.ZstdCompressor().compress(b"foo") == bytes.fromhex("28b52ffd2003190000666f6f")


def test_argument_errors():
    cctx = zstd.ZstdCompressor()
    with pytest.raises(TypeError):
        cctx.multi_compress_to_buffer(True)
    with pytest.raises(TypeError):
        cctx.multi_compress_to_buffer((1, 2))
    with pytest.raises(TypeError, match="item 0 not a bytes like object"):
        cctx.multi_compress_to_buffer(["foo"])
    with pytest.raises(ValueError, match="no source elements found"):
        cctx.multi_compress_to_buffer([])
    with pytest.raises(ValueError, match="source elements are empty"):
        cctx.multi_compress_to_buffer([b"", b"", b""])
    with pytest.raises(ValueError, match="level must be less than 23"):
        zstd.ZstdCompressor(level=23)


def test_collection_input_and_many_items(oracle):
    cctx = zstd.ZstdCompressor(write_checksum=True)
    original = [b"foo1", b"foo2" * 2, b"foo3" * 3, b"foo4" * 4, b"foo5" * 5]
    b1 = zstd.BufferWithSegments(original[0] + original[1], struct.pack("=QQQQ", 0, 4, 4, 8))
    b2 = zstd.BufferWithSegments(b"".join(original[2:]), struct.pack("=QQQQQQ", 0, 12, 12, 16, 28, 20))
    result = cctx.multi_compress_to_buffer(zstd.BufferWithSegmentsCollection(b1, b2))
    assert len(result) == 5
    for i, d in enumerate(original):
        assert result[i].tobytes() == cctx.compress(d)
        assert oracle.decompress(result[i].tobytes(), len(d)) == d
    frames = [b"x" * 64] * 256 + [b"y" * 64] * 256
    result = cctx.multi_compress_to_buffer(frames, threads=-1)
    assert len(result) == 512
    assert all(result[i].tobytes() == result[0].tobytes() for i in range(256))
    assert all(result[i].tobytes() == result[256].tobytes() for i in range(256, 512))
    out = zstd.ZstdDecompressor().multi_decompress_to_buffer(result)
    assert out[0].tobytes() == b"x" * 64 and out[511].tobytes() == b"y" * 64


def _batch_check(ref, blob, off, ln, margin):
    segs = np.stack([off, ln], axis=1).astype(np.uint64)
    res = zstd.ZstdCompressor().multi_compress_to_buffer(zstd.BufferWithSegments(blob, segs.tobytes()))
    cb = res._buffers[0]
    csegs = np.frombuffer(cb._segments, dtype=np.uint64).reshape(-1, 2)
    cblob = np.frombuffer(cb._data, dtype=np.uint8)
    # the reference decoder regenerates the input bit-exact
    rb, rl = ref.batch(False, cblob, np.ascontiguousarray(csegs[:, 0]), np.ascontiguousarray(csegs[:, 1]), threads=8)
    assert np.array_equal(rb, blob)
    # size within the stated margin of the reference's level 3
    refc, _ = ref.batch(True, blob, off, ln, level=3, threads=8)
    assert len(cblob) <= len(refc) * margin, (len(cblob), len(refc))
    # and our own decoder agrees
    d = zstd.ZstdDecompressor().multi_decompress_to_buffer(res)
    assert np.array_equal(np.frombuffer(d._buffers[0]._data, dtype=np.uint8), blob)
    return len(cblob) / len(refc)


def test_batch_4k_text(ref):
    blob, off, ln = corpus.text_segments(4096, 4096)
    _batch_check(ref, blob, off, ln, MARGIN_4K_TEXT)


def test_batch_128k_mix(ref):
    blob, off, ln = corpus.silesia_mix(96, 131072)
    _batch_check(ref, blob, off, ln, MARGIN_128K_MIX)


def test_batch_128k_text(ref):
    blob, off, ln = corpus.text_segments(64, 131072)
    _batch_check(ref, blob, off, ln, MARGIN_128K_TEXT)


def test_ragged_sizes(ref, oracle):
    rng = np.random.default_rng(9)
    text = corpus.text_corpus().tobytes()
    sizes = [int(x) for x in rng.integers(1, 20000, 200)] + [131072, 131073, 262144 + 5, 7, 31, 32, 33]
    items = [text[i * 37: i * 37 + s] for i, s in enumerate(sizes)]
    res = zstd.ZstdCompressor().multi_compress_to_buffer(items)
    assert len(res) == len(items)
    for i, d in enumerate(items):
        assert oracle.decompress(res[i].tobytes(), len(d)) == d, i


def test_dictionary_compression(ref, oracle):
    """Config 4: records compressed against a trained dictionary.  The frames carry the dictionary id,
    decode bit-exact through the reference decoder (with the dictionary) and are far smaller than without
    it; the size margin against the reference's dictionary compression is stated (the dictionary serves as match
    history, start repcodes and entropy tables; measured +0.3 % on the config-4 workload, asserted <= 1.10x)."""
    import os
    from tests import helpers
    dct = open(os.path.join(helpers.GOLDEN, "dict.bin"), "rb").read()
    recs = corpus.json_records(700)[500:]
    d = zstd.ZstdCompressionDict(dct)
    cctx = zstd.ZstdCompressor(dict_data=d)
    res = cctx.multi_compress_to_buffer(recs)
    assert len(res) == len(recs)
    ours = plain = refsz = 0
    nodict = zstd.ZstdCompressor().multi_compress_to_buffer(recs)
    for i, r in enumerate(recs):
        f = res[i].tobytes()
        assert ref.decompress(f, len(r), dct) == r, i
        assert oracle.decompress(f, len(r), dct) == r, i
        ours += len(f)
        plain += len(nodict[i])
        refsz += len(ref.compress(r, level=3, dict_data=dct))
    assert ours < plain * 0.75                     # the dictionary must pay off
    assert ours <= refsz * 1.10, (ours, refsz)     # stated margin vs the reference with the same dictionary (measured +0.3 %)
    out = zstd.ZstdDecompressor(dict_data=d).multi_decompress_to_buffer(res)
    assert [out[i].tobytes() for i in range(len(recs))] == recs
    # dict id is recorded unless disabled
    info_frame = res[0].tobytes()
    assert zstd.ZstdCompressor(dict_data=d, write_dict_id=False).compress(recs[0])[4] & 3 == 0
    assert info_frame[4] & 3 != 0
    # a bigger input whose match crosses from the dictionary into the block
    big = dct[-3000:] + recs[0] * 3
    f = cctx.compress(big)
    assert ref.decompress(f, len(big), dct) == big


def test_window_log_is_honoured(ref):
    """ZstdCompressionParameters(window_log=W) (c-ext/compressionparams.c:46, ZSTD_c_windowLog): frames above 2^W declare a
    2^W window (the reference's frame-parameter reader says so), blocks are cut to the window, and a decoder limited to
    that window accepts the frame."""
    data = corpus.text_corpus(1 << 20).tobytes()[:300000]
    for wl in (10, 13, 16, 18):
        params = zstd.ZstdCompressionParameters(compression_level=3, window_log=wl, write_content_size=1)
        frame = zstd.ZstdCompressor(compression_params=params).compress(data)
        assert ref.decompress(frame, len(data)) == data
        info = zstd.get_frame_parameters(frame) if hasattr(zstd, "get_frame_parameters") else None
        import ctypes as C
        from python_zstandard_b200 import _native
        fi = _native.FrameInfo()
        _native.lib().zb200_frame_info(frame, len(frame), C.byref(fi))
        assert fi.window_size == 1 << wl and fi.content_size == len(data)
        nofcs = zstd.ZstdCompressor(compression_params=zstd.ZstdCompressionParameters(window_log=wl, write_content_size=0)).compress(data)
        limited = zstd.ZstdDecompressor(max_window_size=1 << wl)
        assert limited.decompress(nofcs, max_output_size=len(data)) == data
    with pytest.raises(ValueError):
        zstd.ZstdCompressionParameters(window_log=9)


def test_config4_records_at_size(ref):
    """BASELINE config 4 at a size that fills the machine: 65,536 ~1 KiB JSON-like records (16,384 distinct) with a trained
    dictionary through zb_compress_recs; the unmodified reference decoder regenerates every record with the dictionary,
    the total stays within +1 % of the reference's level 3 with the same dictionary, our own decoder agrees."""
    import os
    n_unique, reps = 16384, 4
    recs = corpus.json_records(n_unique + 2000)
    dct = ref.train_dictionary(112640, recs[:2000])
    recs = recs[2000:]
    one = np.frombuffer(b"".join(recs), dtype=np.uint8)
    ln1 = np.array([len(r) for r in recs], dtype=np.uint64)
    blob = np.tile(one, reps); ln = np.tile(ln1, reps)
    off = np.concatenate([[0], np.cumsum(ln)[:-1]]).astype(np.uint64)
    d = zstd.ZstdCompressionDict(dct)
    bws = zstd.BufferWithSegments(blob.tobytes(), np.stack([off, ln], axis=1).astype(np.uint64).tobytes())
    res = zstd.ZstdCompressor(level=3, dict_data=d).multi_compress_to_buffer(bws)
    assert len(res) == n_unique * reps
    datas, segs, base = [], [], 0
    for b_ in res._buffers:
        d_ = np.frombuffer(b_.tobytes(), dtype=np.uint8)
        g_ = np.frombuffer(b_._segments, dtype=np.uint64).reshape(-1, 2).copy()
        g_[:, 0] += np.uint64(base)
        datas.append(d_); segs.append(g_); base += len(d_)
    seg = np.concatenate(segs)
    back, _ = ref.batch(False, np.concatenate(datas), np.ascontiguousarray(seg[:, 0]), np.ascontiguousarray(seg[:, 1]), dst_len=ln,
                        threads=os.cpu_count(), dict_data=dct)
    assert np.array_equal(back, blob)
    _, rl = ref.batch(True, blob, off, ln, level=3, threads=os.cpu_count(), dict_data=dct)
    assert res.size() <= float(rl.sum()) * 1.01, (res.size(), int(rl.sum()))
    out = zstd.ZstdDecompressor(dict_data=d).multi_decompress_to_buffer(res)
    for i in range(0, len(res), 997):
        assert out[i].tobytes() == recs[i % n_unique]
