"""GPU test of the level >= 4 match finder (two hash tables, zb_compress_blocks<true>).  Sorted last on purpose: the
mode was added at the end of round 1 and validated on the CPU build of the kernels (tests/test_compress_kernel_host.py);
this is its first contact with the device."""
import numpy as np
import pytest

import corpus
import python_zstandard_b200 as zstd

pytestmark = pytest.mark.gpu


def test_level5_round_trips_and_beats_the_single_table():
    from oracle import RefZstd
    ref = RefZstd()
    text = corpus.text_corpus(1 << 20)
    rng = np.random.default_rng(44)
    segs = [bytes(text[:131072]), bytes(text[500000:500000 + 4096]), bytes(text[1234:1234 + 1500]), b"foo" * 12, bytes(9000),
            rng.integers(0, 256, 5000).astype(np.uint8).tobytes(), bytes(text[700000:700000 + 300000])]
    segs += [bytes(text[i * 7000:i * 7000 + 4096]) for i in range(64)]
    one = zstd.ZstdCompressor(level=3, write_checksum=True).multi_compress_to_buffer(segs)
    two = zstd.ZstdCompressor(level=5, write_checksum=True).multi_compress_to_buffer(segs)
    assert len(two) == len(segs)
    for i, s in enumerate(segs):
        assert ref.decompress(two[i].tobytes(), len(s)) == s, i
    assert two.size() < one.size()
    assert len(two[0]) <= len(ref.compress(segs[0], level=3, checksum=True)) * 1.01
    out = zstd.ZstdDecompressor().multi_decompress_to_buffer(two)
    assert [out[i].tobytes() for i in range(len(segs))] == segs
