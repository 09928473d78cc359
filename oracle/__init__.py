"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  It holds

* ``libzo.so``  (zstd_oracle.c): a plain-C restatement of the reference decoder, and
* ``_ref/libzstd_ref.so`` + ``_ref/libref_batch.so``: the UNMODIFIED reference codec
  compiled from /root/reference/zstd/zstd.c (oracle/Makefile) plus a pthread driver
  that restates the reference batch orchestration (ref_batch.c).

Parity status: PINNED (tests/test_oracle.py: reference golden vectors + reference-made frames).
"""
from .pyoracle import Oracle, RefZstd, build, have_ref  # noqa: F401
