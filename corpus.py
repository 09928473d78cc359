"""Deterministic synthetic corpora for tests and bench.py.

Nothing here reads /root/reference (absent on the GPU box).  "S-text" is real source text that
ships with the image (the CPython standard library, present here and on the GPU box alike); if it
is missing a seeded generator stands in.  Everything else is generated from fixed seeds.
"""
import glob
import os
import random

import numpy as np

_TEXT_CACHE = None


def text_corpus(limit=8 << 20):
    """A few MB of real program text: sorted *.py files of the image's standard library."""
    global _TEXT_CACHE
    if _TEXT_CACHE is not None:
        return _TEXT_CACHE
    files = sorted(glob.glob("/usr/lib/python3.12/*.py")) + sorted(glob.glob("/usr/lib/python3.12/*/*.py"))
    parts, total = [], 0
    for f in files:
        try:
            b = open(f, "rb").read()
        except OSError:
            continue
        parts.append(b)
        total += len(b)
        if total >= limit:
            break
    if total < (1 << 20):       # image without a stdlib source tree: seeded stand-in
        rnd = random.Random(99)
        words = ["".join(rnd.choice("abcdefghijklmnopqrstuvwxyz_") for _ in range(rnd.randint(2, 10))) for _ in range(3000)]
        out = []
        while total < limit:
            line = "    " * rnd.randint(0, 3) + " ".join(rnd.choice(words) for _ in range(rnd.randint(2, 12))) + "\n"
            out.append(line.encode())
            total += len(line)
        parts = out
    _TEXT_CACHE = np.frombuffer(b"".join(parts)[:limit], dtype=np.uint8)
    return _TEXT_CACHE


def text_segments(n, size, unique=None):
    """n segments of `size` bytes cut from the text corpus at stride offsets (SURVEY 8d S-text).
    Returns (blob uint8, offsets uint64, lengths uint64); `unique` bounds the distinct offsets."""
    t = text_corpus()
    span = len(t) - size
    idx = np.arange(n, dtype=np.int64)
    if unique:
        idx = idx % unique
    starts = (idx * size + idx * 977) % span
    blob = np.empty(n * size, dtype=np.uint8)
    for i, s in enumerate(starts):
        blob[i * size:(i + 1) * size] = t[s:s + size]
    off = (np.arange(n, dtype=np.uint64) * np.uint64(size))
    ln = np.full(n, size, dtype=np.uint64)
    return blob, off, ln


def json_records(n, seed=1234):
    """JSON-like records of ~1.1 KiB (SURVEY 8d S-json).  Returns a list of bytes."""
    rnd = random.Random(seed)
    vocab = ["".join(rnd.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(rnd.randint(3, 9))) for _ in range(400)]
    tags = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta"]
    out = []
    for i in range(n):
        words = " ".join(rnd.choice(vocab) for _ in range(rnd.randint(80, 140)))
        rec = ('{"id": %d, "ts": "2026-%02d-%02dT%02d:%02d:%02dZ", "tags": ["%s"], "user": {"name": "%s %s", '
               '"address": {"street": "%d %s st", "city": "%s", "zip": "%05d"}}, "score": %d, "text": "%s"}'
               % (i * 7 + rnd.randint(0, 6), rnd.randint(1, 12), rnd.randint(1, 28), rnd.randint(0, 23),
                  rnd.randint(0, 59), rnd.randint(0, 59), '", "'.join(rnd.sample(tags, 3)), rnd.choice(vocab),
                  rnd.choice(vocab), rnd.randint(1, 9999), rnd.choice(vocab), rnd.choice(vocab),
                  rnd.randint(0, 99999), rnd.randint(0, 1000), words))
        out.append(rec.encode())
    return out


def binary_blob(nbytes, seed=7):
    """32-bit little-endian counters with geometric deltas + low-entropy noise (SURVEY 8d S-bin)."""
    rng = np.random.default_rng(seed)
    n = nbytes // 4 + 1
    deltas = rng.geometric(0.2, size=n).astype(np.uint32)
    vals = np.cumsum(deltas, dtype=np.uint32)
    noise = (rng.integers(0, 4, size=n, dtype=np.uint32) << 24)
    return (vals ^ noise).view(np.uint8)[:nbytes].copy()


def silesia_mix(n, size, seed=3):
    """Mixed-compressibility segments: 50% text / 25% json / 15% binary / 5% random / 5% zero,
    shuffled (SURVEY 8d "Silesia-mix").  Returns (blob, offsets, lengths)."""
    rnd = random.Random(seed)
    kinds = (["text"] * 10 + ["json"] * 5 + ["bin"] * 3 + ["rand"] + ["zero"])
    order = [kinds[i % len(kinds)] for i in range(n)]
    rnd.shuffle(order)
    t = text_corpus()
    jrec = np.frombuffer(b"\n".join(json_records(4096)), dtype=np.uint8)
    binb = binary_blob(16 << 20)
    rng = np.random.default_rng(11)
    randb = rng.integers(0, 256, size=4 << 20, dtype=np.uint8)
    blob = np.zeros(n * size, dtype=np.uint8)
    for i, k in enumerate(order):
        dst = blob[i * size:(i + 1) * size]
        if k == "text":
            s = (i * size + i * 977) % (len(t) - size)
            dst[:] = t[s:s + size]
        elif k == "json":
            s = (i * 7919) % (len(jrec) - size)
            dst[:] = jrec[s:s + size]
        elif k == "bin":
            s = ((i * 104729) % (len(binb) - size)) & ~3
            dst[:] = binb[s:s + size]
        elif k == "rand":
            s = (i * 15485863) % (len(randb) - size)
            dst[:] = randb[s:s + size]
    off = (np.arange(n, dtype=np.uint64) * np.uint64(size))
    ln = np.full(n, size, dtype=np.uint64)
    return blob, off, ln
