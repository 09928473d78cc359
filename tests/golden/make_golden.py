"""Regenerate tests/golden/*.bin from the UNMODIFIED reference codec.

Run in the build container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
Each vector is `<name>.zst` (a frame produced by the reference at the stated level/flags) and
`<name>.raw` (the input).  `dict.bin` is a dictionary trained by the reference's ZDICT_trainFromBuffer.
The inputs are generated from fixed seeds / the image's stdlib text, see below.
"""
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import corpus  # noqa: E402
from oracle import RefZstd  # noqa: E402


def main():
    r = RefZstd()
    text = corpus.text_corpus().tobytes()
    rnd = random.Random(42)
    vectors = {
        "text4k_l3": (text[10000:14096], dict(level=3)),
        "text4k_l3_ck": (text[20000:24096], dict(level=3, checksum=True)),
        "text1k_l1": (text[500:1524], dict(level=1)),
        "text20k_l9": (text[30000:50000], dict(level=9)),
        "text20k_l19": (text[60000:80000], dict(level=19)),
        "text64k_l3": (text[100000:165536], dict(level=3)),
        "text150k_l3_multiblock": (text[200000:350000], dict(level=3, checksum=True)),
        "text4k_nocs": (text[40000:44096], dict(level=3, content_size=False)),
        "rle100k": (b"\x07" * 100000, dict(level=3)),
        "rand3k": (bytes(rnd.getrandbits(8) for _ in range(3000)), dict(level=3)),
        "abc_lowentropy": (bytes(rnd.choice(b"abc") for _ in range(30000)), dict(level=3)),
        "tiny_foo": (b"foo" * 12, dict(level=3, checksum=True)),
        "neg5": (text[70000:78000], dict(level=-5)),
    }
    recs = corpus.json_records(600)
    dct = r.train_dictionary(16384, recs[:500])
    open(os.path.join(HERE, "dict.bin"), "wb").write(dct)
    for i in range(3):
        vectors["json_dict_%d" % i] = (recs[500 + i], dict(level=3, dict_data=dct))
    for name, (raw, kw) in vectors.items():
        frame = r.compress(raw, **kw)
        assert r.decompress(frame, len(raw), kw.get("dict_data", b"")) == raw
        open(os.path.join(HERE, name + ".raw"), "wb").write(raw)
        open(os.path.join(HERE, name + ".zst"), "wb").write(frame)
        print(name, len(raw), "->", len(frame))


if __name__ == "__main__":
    main()
