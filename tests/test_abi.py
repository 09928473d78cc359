"""The C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol that
include/zb200.h declares.  No compute calls here."""
import ctypes
import os
import re

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    entry.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "python_zstandard_b200", "libzb200.so"))
    header = open(os.path.join(ROOT, "include", "zb200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = sorted(set(re.findall(r"\b(zb200_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_error_strings_match_reference_texts():
    entry.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "python_zstandard_b200", "libzb200.so"))
    lib.zb200_error_string.restype = ctypes.c_char_p
    assert lib.zb200_error_string(20) == b"Data corruption detected"
    assert lib.zb200_error_string(70) == b"Destination buffer is too small"
    assert lib.zb200_error_string(10) == b"Unknown frame descriptor"


def test_frame_info_host_parse():
    import python_zstandard_b200 as zstd
    from tests import helpers
    assert zstd.frame_content_size(helpers.KAT_FOO) == 3
    assert zstd.frame_content_size(helpers.KAT_EMPTY_NOFCS) == -1
    assert zstd.frame_header_size(helpers.KAT_FOO) == 6
    for name, frame, raw, dct in helpers.golden_vectors():
        if "nocs" in name:
            assert zstd.frame_content_size(frame) == -1
        else:
            assert zstd.frame_content_size(frame) == len(raw), name


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product path must fail loudly, not fall back."""
    import pytest
    import python_zstandard_b200 as zstd
    from python_zstandard_b200 import _native
    if _native.device_count() > 0:
        pytest.skip("a GPU is present")
    from tests import helpers
    with pytest.raises(_native.NativeError, match="no CPU fallback"):
        zstd.ZstdDecompressor().decompress(helpers.KAT_FOO)


def test_compress_bound_equals_the_reference():
    """zb200_compress_bound == ZSTD_compressBound (zstd/zstd.c:4547) -- a host-only call of the C ABI."""
    entry.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "python_zstandard_b200", "libzb200.so"))
    lib.zb200_compress_bound.restype = ctypes.c_uint64
    lib.zb200_compress_bound.argtypes = [ctypes.c_uint64]
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libzstd_ref.so")
    expect = None
    if os.path.exists(ref_path):
        ref = ctypes.CDLL(ref_path)
        ref.ZSTD_compressBound.restype = ctypes.c_size_t
        ref.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
        expect = ref.ZSTD_compressBound
    for n in (0, 1, 255, 1024, 4096, 131071, 131072, 131073, 1 << 20, (1 << 32) + 5):
        want = expect(n) if expect else n + (n >> 8) + (((128 << 10) - n) >> 11 if n < (128 << 10) else 0)
        assert lib.zb200_compress_bound(n) == want, n
    assert lib.zb200_compress_bound(131072) == 131584 and lib.zb200_compress_bound(4096) == 4174


def test_multi_device_partition_equals_the_python_rule():
    """zb200_*_batch_multi cuts a batch exactly like sharding.split_ranges (the reference's worker partition).  Without a
    device the call fails when it creates its first context -- after the partition has been written to first_item[]."""
    import ctypes as C
    import numpy as np
    from python_zstandard_b200 import _native
    from python_zstandard_b200.sharding import split_ranges
    L = _native.lib()
    rng = np.random.default_rng(3)
    for n, parts in ((1, 4), (2, 2), (5, 8), (100, 3), (1000, 7), (4096, 8), (77, 1)):
        lens = rng.integers(0, 5000, n).astype(np.uint64)
        if n > 3:
            lens[n // 2] = 1 << 20                       # one dominant item
        segs = np.stack([np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64), lens], axis=1).astype(np.uint64)
        devs = (C.c_int * parts)(*([0] * parts))
        results = (C.c_void_p * parts)()
        first = (C.c_size_t * parts)()
        src = np.zeros(max(1, int(lens.sum())), dtype=np.uint8)
        rc = L.zb200_decompress_batch_multi(devs, parts, src.ctypes.data, segs.ctypes.data, n, None, None, 0, None, 0, results, first)
        want = split_ranges(lens, parts)
        got = [first[k] for k in range(parts)]
        assert got[:len(want)] == [lo for lo, _ in want], (n, parts)
        assert all(g == n for g in got[len(want):])
        if rc != 0:                                     # no GPU here: a clean failure with a message, nothing returned
            assert all(not results[k] for k in range(parts)) and L.zb200_multi_last_error()
        else:
            for k in range(parts):
                if results[k]:
                    L.zb200_result_free(results[k])
