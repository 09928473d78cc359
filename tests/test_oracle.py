"""Pins the oracle: the plain-C restatement (oracle/zstd_oracle.c) against
(a) the known-answer vectors of the reference's own tests,
(b) golden frames made by the unmodified reference (tests/golden, see make_golden.py),
(c) frames produced live by oracle/_ref when it is present."""
import random

import pytest

from oracle import Oracle, RefZstd, have_ref
from tests import helpers


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def test_reference_known_answers(oracle):
    assert oracle.decompress(helpers.KAT_EMPTY_NOFCS, 0) == b""
    assert oracle.decompress(helpers.KAT_EMPTY_FCS, 0) == b""
    assert oracle.decompress(helpers.KAT_FOO, 3) == b"foo"
    assert oracle.decompress(helpers.KAT_LARGE, 131073) == b"f" * 131072 + b"o"


def test_golden_vectors(oracle):
    vecs = helpers.golden_vectors()
    assert len(vecs) >= 12
    for name, frame, raw, dct in vecs:
        assert oracle.decompress(frame, len(raw), dct) == raw, name


def test_xxh64_known_values(oracle):
    # XXH64 test vectors (seed 0): empty input and "abc"
    assert oracle.xxh64(b"") == 0xEF46DB3751D8E999
    assert oracle.xxh64(b"abc") == 0x44BC2CF5AD770999


def test_corruption_is_rejected(oracle):
    name, frame, raw, dct = helpers.golden_vectors()[0]
    bad = frame[:15] + b"extra" + frame[15:]
    with pytest.raises(Oracle.Error, match="Data corruption detected|Destination buffer is too small|Src size"):
        oracle.decompress(bad, len(raw) + 100)
    with pytest.raises(Oracle.Error, match="Unknown frame descriptor"):
        oracle.decompress(b"foobarbaz", 100)
    # checksum flip
    for n2, f2, r2, d2 in helpers.golden_vectors():
        if n2 == "text4k_l3_ck":
            f3 = f2[:-1] + bytes([f2[-1] ^ 1])
            with pytest.raises(Oracle.Error, match="checksum"):
                oracle.decompress(f3, len(r2))


def test_trace_matches_output(oracle):
    name, frame, raw, dct = [v for v in helpers.golden_vectors() if v[0] == "text4k_l3"][0]
    out, lits, seqs, per_block = oracle.trace(frame, len(raw))
    assert out == raw
    assert sum(ll + ml for ll, ml, off in seqs) + (len(lits) - sum(ll for ll, _, _ in seqs)) == len(raw)
    assert per_block == [len(seqs)]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_against_live_reference(oracle):
    ref = RefZstd()
    rnd = random.Random(7)
    import corpus
    text = corpus.text_corpus().tobytes()
    cases = [b"", b"a", b"ab" * 500, text[:3000], text[1000:70000], bytes(rnd.getrandbits(8) for _ in range(2000)),
             b"\0" * 70000, text[:200000]]
    for level in (1, 3, 6, 12, 19, -3):
        for data in cases:
            for ck in (False, True):
                frame = ref.compress(data, level=level, checksum=ck)
                assert oracle.decompress(frame, len(data)) == data
                assert oracle.frame_compressed_size(frame) == len(frame)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_mutated_frames_vs_live_reference(oracle):
    """The oracle never accepts what the reference rejects, and agrees byte for byte when both accept."""
    import numpy as np
    import corpus
    ref = RefZstd()
    rng = np.random.default_rng(3)
    base = ref.compress(corpus.text_corpus().tobytes()[3000:7096], level=3)
    both = 0
    for k in range(300):
        f = bytearray(base)
        f[int(rng.integers(4, len(f)))] ^= 1 << int(rng.integers(0, 8))
        f = bytes(f)
        try:
            exp = ref.decompress(f, 4096)
        except RefZstd.Error:
            exp = None
        try:
            got = oracle.decompress(f, 4096)
        except Oracle.Error:
            got = None
        if exp is None:
            assert got is None
        elif got is not None:
            assert got == exp
            both += 1
    assert both > 50


def test_from_level_equals_getcparams():
    """ZstdCompressionParameters.from_level(level, source_size, dict_size) carries what ZSTD_getCParams(...) of the unmodified
    reference returns -- all seven fields, every level (negative ones, 0, beyond 22) and a sweep of source / dictionary sizes
    across the table and downsizing thresholds.  window_log is the field that reaches this backend."""
    import ctypes as C
    import python_zstandard_b200 as zstd

    class CP(C.Structure):
        _fields_ = [(k, C.c_uint) for k in ("windowLog", "chainLog", "hashLog", "searchLog", "minMatch", "targetLength", "strategy")]
    from oracle import RefZstd, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not present")
    Z = RefZstd().Z
    Z.ZSTD_getCParams.restype = CP
    Z.ZSTD_getCParams.argtypes = [C.c_int, C.c_ulonglong, C.c_size_t]
    sizes = [0, 1, 63, 64, 65, 1000, 16 << 10, (16 << 10) + 1, 100000, 128 << 10, (128 << 10) + 1, 256 << 10, (256 << 10) + 1,
             1 << 20, 5 << 20, 1 << 27, 1 << 30, (1 << 30) + 1, 1 << 33]
    dicts = [0, 1, 500, 16000, 112640, 130000, 300000, 1 << 22, 1 << 29, (1 << 30) + 5, 1 << 31]
    names = ("window_log", "chain_log", "hash_log", "search_log", "min_match", "target_length", "strategy")
    checked = 0
    for level in list(range(-7, 23)) + [30, -200000]:
        for s in sizes:
            for d in dicts:
                w = Z.ZSTD_getCParams(level, s, d)
                want = (w.windowLog, w.chainLog, w.hashLog, w.searchLog, w.minMatch, w.targetLength, w.strategy)
                p = zstd.ZstdCompressionParameters.from_level(level, source_size=s, dict_size=d)
                assert tuple(getattr(p, k) for k in names) == want, (level, s, d)
                checked += 1
    assert checked > 6000
    p = zstd.ZstdCompressionParameters.from_level(3, source_size=1000, window_log=15)
    assert p.window_log == 15 and p.hash_log == Z.ZSTD_getCParams(3, 1000, 0).hashLog
    with pytest.raises(zstd.ZstdError, match="not supported by the B200 backend"):
        zstd.ZstdCompressionParameters.from_level(3, hash_log=20)            # asking for a different match finder is still refused
