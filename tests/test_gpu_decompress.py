"""Parity tests proper: the CUDA path (through the Python mirror -> C ABI -> kernels) against the
oracle and the committed golden vectors.  Bit-exact (byte work: tolerance = 0)."""
import struct

import numpy as np
import pytest

import corpus
import python_zstandard_b200 as zstd
from oracle import Oracle, RefZstd, have_ref
from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.fixture(scope="module")
def ref():
    if not have_ref():
        pytest.skip("oracle/_ref not present")
    return RefZstd()


def test_known_answer_frames():
    d = zstd.ZstdDecompressor()
    assert d.decompress(helpers.KAT_EMPTY_FCS) == b""
    assert d.decompress(helpers.KAT_FOO) == b"foo"
    assert d.decompress(helpers.KAT_LARGE, max_output_size=131073) == b"f" * 131072 + b"o"
    out = d.multi_decompress_to_buffer([helpers.KAT_FOO, helpers.KAT_LARGE],
                                       decompressed_sizes=struct.pack("=QQ", 3, 131073))
    assert out[0].tobytes() == b"foo" and out[1].tobytes() == b"f" * 131072 + b"o"


def test_golden_vectors(oracle):
    vecs = helpers.golden_vectors()
    plain = [v for v in vecs if not v[3] and "nocs" not in v[0]]
    out = zstd.ZstdDecompressor().multi_decompress_to_buffer([v[1] for v in plain])
    assert len(out) == len(plain)
    assert out.size() == sum(len(v[2]) for v in plain)
    for i, (name, frame, raw, _) in enumerate(plain):
        assert out[i].tobytes() == raw == oracle.decompress(frame, len(raw)), name
    # frame without content size: sizes must be supplied
    nocs = [v for v in vecs if "nocs" in v[0]][0]
    with pytest.raises(ValueError, match="could not determine decompressed size of item 0"):
        zstd.ZstdDecompressor().multi_decompress_to_buffer([nocs[1]])
    out = zstd.ZstdDecompressor().multi_decompress_to_buffer([nocs[1]], decompressed_sizes=struct.pack("=Q", len(nocs[2])))
    assert out[0].tobytes() == nocs[2]
    assert zstd.ZstdDecompressor().decompress(nocs[1], max_output_size=1 << 16) == nocs[2]


def test_dictionary_frames(oracle):
    vecs = [v for v in helpers.golden_vectors() if v[3]]
    d = zstd.ZstdCompressionDict(vecs[0][3])
    assert d.dict_id() != 0
    out = zstd.ZstdDecompressor(dict_data=d).multi_decompress_to_buffer([v[1] for v in vecs])
    for i, (name, frame, raw, dct) in enumerate(vecs):
        assert out[i].tobytes() == raw == oracle.decompress(frame, len(raw), dct), name


def test_input_kinds_and_sizes(ref):
    # mirrors reference tests/test_decompressor_multi_decompress_to_buffer.py:36-177
    original = [b"foo" * 4, b"bar" * 6, b"baz" * 8]
    frames = [ref.compress(d) for d in original]
    dctx = zstd.ZstdDecompressor()
    result = dctx.multi_decompress_to_buffer(frames)
    assert len(result) == 3 and result.size() == sum(map(len, original))
    assert [result[i].tobytes() for i in range(3)] == original
    assert result[0].offset == 0 and len(result[0]) == 12 and len(result[1]) == 18
    sizes = struct.pack("=QQQ", *map(len, original))
    result = dctx.multi_decompress_to_buffer(frames, decompressed_sizes=sizes)
    assert [result[i].tobytes() for i in range(3)] == original
    offs, pos = [], 0
    for f in frames:
        offs += [pos, len(f)]
        pos += len(f)
    b = zstd.BufferWithSegments(b"".join(frames), struct.pack("=6Q", *offs))
    result = dctx.multi_decompress_to_buffer(b)
    assert [result[i].tobytes() for i in range(3)] == original
    nofcs = [ref.compress(d, content_size=False) for d in original]
    offs, pos = [], 0
    for f in nofcs:
        offs += [pos, len(f)]
        pos += len(f)
    b = zstd.BufferWithSegments(b"".join(nofcs), struct.pack("=6Q", *offs))
    result = dctx.multi_decompress_to_buffer(b, decompressed_sizes=sizes)
    assert [result[i].tobytes() for i in range(3)] == original
    c = zstd.BufferWithSegmentsCollection(zstd.BufferWithSegments(frames[0], struct.pack("=QQ", 0, len(frames[0]))),
                                          zstd.BufferWithSegments(frames[1] + frames[2], struct.pack("=QQQQ", 0, len(frames[1]), len(frames[1]), len(frames[2]))))
    result = dctx.multi_decompress_to_buffer(c, threads=3)
    assert [result[i].tobytes() for i in range(3)] == original
    with pytest.raises(ValueError, match="decompressed_sizes size mismatch; expected 24, got 16"):
        dctx.multi_decompress_to_buffer(frames, decompressed_sizes=sizes[:16])
    with pytest.raises(zstd.ZstdError, match="error decompressing item 1: decompressed 18 bytes; expected 19"):
        dctx.multi_decompress_to_buffer(frames, decompressed_sizes=struct.pack("=QQQ", 12, 19, 24))


def test_item_failure(ref):
    # mirrors reference tests/test_decompressor_multi_decompress_to_buffer.py:209-227
    frames = [ref.compress(b"x" * 128), ref.compress(b"y" * 128)]
    frames[1] = frames[1][0:15] + b"extra" + frames[1][15:]
    with pytest.raises(zstd.ZstdError, match="error decompressing item 1: (Data corruption detected|Destination buffer is too small)"):
        zstd.ZstdDecompressor().multi_decompress_to_buffer(frames)
    with pytest.raises(TypeError, match="item 0 not a bytes like object"):
        zstd.ZstdDecompressor().multi_decompress_to_buffer(["foo"])
    with pytest.raises(TypeError):
        zstd.ZstdDecompressor().multi_decompress_to_buffer((1, 2))
    with pytest.raises(ValueError, match="could not determine decompressed size of item 0"):
        zstd.ZstdDecompressor().multi_decompress_to_buffer([b"foobarbaz"])


def test_mutated_frames_never_accept_what_the_reference_rejects(ref, oracle):
    """Bit-flipped frames: whatever the reference rejects we reject; whatever we accept the
    reference accepts with identical bytes; and the CUDA path agrees with the oracle exactly.
    (We follow the reference's portable Huffman body, which insists that every stream is consumed
    exactly -- zstd/zstd.c:39958-39959; its fast loop :40100-40140 lets some such frames through,
    so "reference accepts, we reject" is allowed and documented in DESIGN.md.)"""
    rng = np.random.default_rng(3)
    text = corpus.text_corpus().tobytes()
    base = ref.compress(text[3000:7096], level=3)
    d = zstd.ZstdDecompressor()
    sizes = struct.pack("=Q", 4096)
    both = 0
    for k in range(200):
        f = bytearray(base)
        pos = int(rng.integers(4, len(f)))
        f[pos] ^= 1 << int(rng.integers(0, 8))
        f = bytes(f)
        try:
            exp = ref.decompress(f, 4096)
            exp = exp if len(exp) == 4096 else None
        except RefZstd.Error:
            exp = None
        try:
            orc = oracle.decompress(f, 4096)
            orc = orc if len(orc) == 4096 else None
        except Oracle.Error:
            orc = None
        try:
            got = d.multi_decompress_to_buffer([f], decompressed_sizes=sizes)[0].tobytes()
        except (zstd.ZstdError, ValueError):
            got = None
        assert got == orc, k
        if exp is None:
            assert got is None, k
        elif got is not None:
            assert got == exp, k
            both += 1
    assert both > 50


def test_levels_and_shapes(ref, oracle):
    text = corpus.text_corpus().tobytes()
    rng = np.random.default_rng(5)
    cases = [b"a", b"foo" * 12, b"x" * 64, text[:1000], text[5000:9096], text[:65536], text[:300000],
             rng.integers(0, 256, 5000, dtype=np.uint8).tobytes(), b"\0" * 100000,
             bytes(rng.choice(list(b"abcd"), 20000).astype(np.uint8)), text[: 1 << 20]]
    d = zstd.ZstdDecompressor()
    for level in (1, 3, 5, 9, 19, -5):
        for ck in (False, True):
            frames = [ref.compress(c, level=level, checksum=ck) for c in cases]
            out = d.multi_decompress_to_buffer(frames)
            for i, c in enumerate(cases):
                assert out[i].tobytes() == c, (level, ck, i)


def test_batch_4k_frames_equals_reference(ref):
    n = 16384
    blob, off, ln = corpus.text_segments(n, 4096, unique=4096)
    cblob, clens = ref.batch(True, blob, off, ln, threads=8)
    coff = np.concatenate([[0], np.cumsum(clens)[:-1]]).astype(np.uint64)
    segs = np.stack([coff, clens], axis=1).astype(np.uint64)
    out = zstd.ZstdDecompressor().multi_decompress_to_buffer(zstd.BufferWithSegments(cblob, segs.tobytes()))
    assert len(out) == n
    got = np.frombuffer(out._buffers[0]._data, dtype=np.uint8)
    assert np.array_equal(got, blob)
    # and the reference decoder agrees on a sample (the oracle of record)
    rblob, rlens = ref.batch(False, cblob, coff, clens, threads=8)
    assert np.array_equal(rblob, blob)


def test_mixed_128k_segments(ref):
    n = 96
    blob, off, ln = corpus.silesia_mix(n, 131072)
    cblob, clens = ref.batch(True, blob, off, ln, threads=8)
    coff = np.concatenate([[0], np.cumsum(clens)[:-1]]).astype(np.uint64)
    segs = np.stack([coff, clens], axis=1).astype(np.uint64)
    out = zstd.ZstdDecompressor().multi_decompress_to_buffer(zstd.BufferWithSegments(cblob, segs.tobytes()))
    got = np.frombuffer(out._buffers[0]._data, dtype=np.uint8)
    assert np.array_equal(got, blob)


def test_content_checksum_is_verified(oracle):
    """A frame whose stored XXH64 does not match its content is rejected (zstd/zstd.c:44271-44277)."""
    vec = [v for v in helpers.golden_vectors() if v[0] == "text4k_l3_ck"][0]
    frame, raw = vec[1], vec[2]
    d = zstd.ZstdDecompressor()
    assert d.multi_decompress_to_buffer([frame])[0].tobytes() == raw
    bad = frame[:-1] + bytes([frame[-1] ^ 0x40])
    with pytest.raises(zstd.ZstdError, match="error decompressing item 1: Restored data doesn't match checksum"):
        d.multi_decompress_to_buffer([frame, bad])
    with pytest.raises(zstd.ZstdError, match="decompression error: Restored data doesn't match checksum"):
        d.decompress(bad)
    with pytest.raises(Oracle.Error, match="checksum"):
        oracle.decompress(bad, len(raw))
    # a multi-block checksummed frame still verifies
    big = [v for v in helpers.golden_vectors() if v[0] == "text150k_l3_multiblock"][0]
    assert d.decompress(big[1]) == big[2]


def test_size_mismatch_is_reported_for_checksummed_frames_too(ref):
    """c-ext/decompressor.c:1151-1162: every item whose regenerated size differs from decompressed_sizes[i] fails, with the
    size it did regenerate in the message -- frames with a content checksum included."""
    data = [b"a" * 1000, corpus.text_corpus(1 << 16).tobytes()[:5000]]
    for checksum in (False, True):
        frames = [ref.compress(d, checksum=checksum) for d in data]
        with pytest.raises(zstd.ZstdError, match="error decompressing item 1: decompressed 5000 bytes; expected 5001"):
            zstd.ZstdDecompressor().multi_decompress_to_buffer(frames, decompressed_sizes=struct.pack("=QQ", 1000, 5001))
        nofcs = [ref.compress(d, checksum=checksum, content_size=False) for d in data]
        with pytest.raises(zstd.ZstdError, match="error decompressing item 0: decompressed 1000 bytes; expected 1024"):
            zstd.ZstdDecompressor().multi_decompress_to_buffer(nofcs, decompressed_sizes=struct.pack("=QQ", 1024, 5000))
        out = zstd.ZstdDecompressor().multi_decompress_to_buffer(nofcs, decompressed_sizes=struct.pack("=QQ", 1000, 5000))
        assert [out[i].tobytes() for i in range(2)] == data


def test_max_window_size_is_enforced(ref):
    """ZstdDecompressor(max_window_size=...) -> ZSTD_DCtx_setMaxWindowSize (c-ext/decompressor.c:22-24).  As in the reference's
    streaming decoder (zstd/zstd.c:45406-45453) the limit binds for frames without a content size in their header; the
    message is the reference's (zstd/zstd.c:3585)."""
    data = corpus.text_corpus(1 << 20).tobytes()[:300000]
    with_fcs, without = ref.compress(data), ref.compress(data, content_size=False)
    small = zstd.ZstdDecompressor(max_window_size=1 << 12)
    assert small.multi_decompress_to_buffer([with_fcs])[0].tobytes() == data            # single-pass path: no window buffer needed
    with pytest.raises(zstd.ZstdError, match="error decompressing item 0: Frame requires too much memory for decoding"):
        small.multi_decompress_to_buffer([without], decompressed_sizes=struct.pack("=Q", len(data)))
    with pytest.raises(zstd.ZstdError, match="decompression error: Frame requires too much memory for decoding"):
        small.decompress(without, max_output_size=len(data))
    roomy = zstd.ZstdDecompressor(max_window_size=1 << 22)
    assert roomy.multi_decompress_to_buffer([without], decompressed_sizes=struct.pack("=Q", len(data)))[0].tobytes() == data
    assert zstd.ZstdDecompressor().decompress(without, max_output_size=len(data)) == data


def test_device_resident_results_survive_the_next_call(ref):
    """ZB200_DST_DEVICE results own their device allocation: a later call on the same context leaves them alone."""
    import ctypes as C
    from python_zstandard_b200 import _native
    L = _native.lib(); ctx = _native.Context.get(0)
    a, b = [corpus.text_corpus(1 << 20).tobytes()[i * 50000:(i + 1) * 50000] for i in range(2)]
    def run(data):
        f = np.frombuffer(ref.compress(data), dtype=np.uint8).copy()
        seg = np.array([[0, len(f)]], dtype=np.uint64)
        r = C.c_void_p()
        ctx.check(L.zb200_decompress_batch(ctx.h, f.ctypes.data, seg.ctypes.data, 1, None, None, _native.DST_DEVICE, C.byref(r)), "decompress")
        return r
    ra = run(a); rb = run(b)
    for r, want in ((ra, a), (rb, b)):
        out = np.empty(len(want), dtype=np.uint8)
        ctx.check(L.zb200_memcpy_d2h(ctx.h, out.ctypes.data, L.zb200_result_data(r), len(want)), "d2h")
        assert out.tobytes() == want
        L.zb200_result_free(r)


def test_c_abi_one_batch_over_several_contexts(ref):
    """zb200_decompress_batch_multi / zb200_compress_batch_multi: the reference's `threads` partition inside the C ABI.
    Three contexts on device 0 stand in for three devices (bench.py --gpus N and tools run the real thing): the ranges
    come back in item order, every item as the single-context call returns it; a dictionary is digested per context;
    a damaged item is reported inside its range's result."""
    import ctypes as C
    from python_zstandard_b200 import _native
    L = _native.lib()
    rng = np.random.default_rng(11)
    text = corpus.text_corpus(1 << 20)
    items = [text[o:o + int(s)].tobytes() for o, s in zip(rng.integers(0, (1 << 20) - 70000, 90), rng.integers(1, 60000, 90))]
    samples = [text[i * 997:i * 997 + 600].tobytes() for i in range(400)]
    dict_raw = ref.train_dictionary(16384, samples)
    for use_dict in (False, True):
        frames = [ref.compress(s, level=3, dict_data=dict_raw if use_dict else b"") for s in items]
        blob = np.frombuffer(b"".join(frames), dtype=np.uint8)
        lens = np.array([len(f) for f in frames], dtype=np.uint64)
        segs = np.stack([np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64), lens], axis=1).astype(np.uint64)
        devs = (C.c_int * 3)(0, 0, 0); results = (C.c_void_p * 3)(); first = (C.c_size_t * 3)()
        rc = L.zb200_decompress_batch_multi(devs, 3, blob.ctypes.data, segs.ctypes.data, len(items), None,
                                            dict_raw if use_dict else None, len(dict_raw) if use_dict else 0, None, 0, results, first)
        assert rc == 0, L.zb200_multi_last_error()
        assert first[0] == 0 and 0 < first[1] < first[2] < len(items)
        got = []
        for k in range(3):
            assert not L.zb200_result_first_error(results[k], None, None, None, None)
            n_k = L.zb200_result_count(results[k]); base = L.zb200_result_data(results[k])
            st = np.ctypeslib.as_array(C.cast(L.zb200_result_segments(results[k]), C.POINTER(C.c_uint64)), shape=(n_k, 2))
            got += [C.string_at(base + int(o), int(l)) for o, l in st]
            L.zb200_result_free(results[k])
        assert got == items
    # compression over the same three contexts: every frame regenerated by the reference, in item order
    blob = np.frombuffer(b"".join(items), dtype=np.uint8)
    lens = np.array([len(s) for s in items], dtype=np.uint64)
    segs = np.stack([np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64), lens], axis=1).astype(np.uint64)
    results = (C.c_void_p * 3)(); first = (C.c_size_t * 3)()
    rc = L.zb200_compress_batch_multi((C.c_int * 3)(0, 0, 0), 3, blob.ctypes.data, segs.ctypes.data, len(items), None, None, 0, 0, results, first)
    assert rc == 0, L.zb200_multi_last_error()
    back = []
    for k in range(3):
        n_k = L.zb200_result_count(results[k]); base = L.zb200_result_data(results[k])
        st = np.ctypeslib.as_array(C.cast(L.zb200_result_segments(results[k]), C.POINTER(C.c_uint64)), shape=(n_k, 2))
        back += [ref.decompress(C.string_at(base + int(o), int(l)), 70000) for o, l in st]
        L.zb200_result_free(results[k])
    assert back == items
    # a damaged item: the call succeeds, the item's error sits in its range's result
    frames = [ref.compress(s, level=3) for s in items]
    bad = bytearray(frames[80]); bad[len(bad) // 2] ^= 0x55; frames[80] = bytes(bad[:len(bad) - 3])
    blob = np.frombuffer(b"".join(frames), dtype=np.uint8)
    lens = np.array([len(f) for f in frames], dtype=np.uint64)
    segs = np.stack([np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64), lens], axis=1).astype(np.uint64)
    rc = L.zb200_decompress_batch_multi((C.c_int * 3)(0, 0, 0), 3, blob.ctypes.data, segs.ctypes.data, len(items), None, None, 0, None, 0, results, first)
    assert rc == 0
    item = C.c_size_t()
    flags = [bool(L.zb200_result_first_error(results[k], C.byref(item), None, None, None)) for k in range(3)]
    k80 = max(k for k in range(3) if first[k] <= 80)
    assert flags == [k == k80 for k in range(3)]
    L.zb200_result_first_error(results[k80], C.byref(item), None, None, None)
    assert first[k80] + item.value == 80
    for k in range(3):
        L.zb200_result_free(results[k])
