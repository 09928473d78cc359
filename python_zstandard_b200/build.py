"""Build libzb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libzb200.so")
SOURCES = ["zb_decode.cu", "zb_encode.cu", "zb_api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math"]
if os.environ.get("ZB200_DEBUG_BLOCKS"):                 # debugging build of the block-parallel decode path (device printf)
    NVCC_FLAGS.append("-DZB_DEBUG_BLOCKS")
    LIB = os.path.join(HERE, "libzb200_dbg.so")
TIMERS = bool(os.environ.get("ZB200_PHASE_TIMERS"))     # tuning builds only: per-phase clock64 counters inside the kernels;
if TIMERS:                                              # they go to their own library (load it with ZB200_LIB=...)
    NVCC_FLAGS.append("-DZB_PHASE_TIMERS")
    LIB = os.path.join(HERE, "libzb200_timers.so")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "zb200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs, procs = [], []
    for src in SOURCES:                                  # the three translation units compile side by side
        obj = os.path.join(CSRC, src.replace(".cu", ".timers.o" if TIMERS else (".dbg.o" if os.environ.get("ZB200_DEBUG_BLOCKS") else ".o")))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
