import glob
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_vectors():
    """[(name, frame bytes, raw bytes, dict bytes)] committed under tests/golden (made by the reference)."""
    dct = open(os.path.join(GOLDEN, "dict.bin"), "rb").read()
    out = []
    for z in sorted(glob.glob(os.path.join(GOLDEN, "*.zst"))):
        name = os.path.basename(z)[:-4]
        raw = open(z[:-4] + ".raw", "rb").read()
        out.append((name, open(z, "rb").read(), raw, dct if "dict" in name else b""))
    return out


# known-answer vectors quoted from the reference's own tests
KAT_EMPTY_NOFCS = bytes.fromhex("28b52ffd0000010000")          # tests/test_compressor_compress.py:19
KAT_EMPTY_FCS = bytes.fromhex("28b52ffd2000010000")            # tests/test_compressor_compress.py:28
KAT_FOO = bytes.fromhex("28b52ffd2003190000666f6f")            # tests/test_compressor_compress.py:34
# level 1, no content size, 131072 x 'f' + 'o'                  tests/test_compressor_compress.py:58-68
KAT_LARGE = bytes.fromhex("28b52ffd0040540000106666" "0100fbff39c002" "090000" "6f")
