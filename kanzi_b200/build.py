"""Build libkanzi_b200.so (nvcc, sm_100a) in-tree.  No torch dependency: a JVM loads this library.
KZG_A1_TIMING=1 in the environment compiles the clock64 phase probes of the ANS / ROLZ kernels in (developer aid: they printf per chunk)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkanzi_b200.so")
NVCC_FLAGS = (["-DKZG_A1_TIMING", "-DKZG_RZ_TIMING"] if os.environ.get("KZG_A1_TIMING") else []) + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Xcudafe", "--diag_suppress=177", "-cudart", "shared"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "kzg.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in sources():
        obj = os.path.join(HERE, "build", os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if (not force) and os.path.exists(obj) and os.path.getmtime(obj) > max(
                [os.path.getmtime(src)] + [os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
                + [os.path.getmtime(os.path.join(HERE, "..", "include", "kzg.h"))]):
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {os.path.basename(src)} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, "-shared", "-cudart", "shared", "-o", LIB] + objs)
    return LIB


NATIVE_SRC = os.path.join(HERE, "..", "tests", "native", "kzg_blocks_mt.c")
NATIVE_BIN = os.path.join(HERE, "..", "tests", "native", "kzg_blocks_mt")


def build_native():
    """The pthread driver of the per-block C ABI (tests/native/kzg_blocks_mt.c): what Kanzi's Java pool threads do, in C."""
    if os.path.exists(NATIVE_BIN) and os.path.getmtime(NATIVE_BIN) > max(os.path.getmtime(NATIVE_SRC), os.path.getmtime(LIB)):
        return NATIVE_BIN
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-Wall", "-pthread", NATIVE_SRC, "-L" + HERE, "-lkanzi_b200", "-Wl,-rpath,$ORIGIN/../../kanzi_b200", "-o", NATIVE_BIN])
    return NATIVE_BIN


SIBLING_SRC = os.path.join(HERE, "..", "tests", "native", "kzg_sibling_check.c")
SIBLING_BIN = os.path.join(HERE, "..", "tests", "native", "kzg_sibling_check")


def build_sibling_check():
    """tests/native/kzg_sibling_check.c: torch-free parity check of the sibling codecs against the oracle (dlopen), for short GPU slots."""
    if os.path.exists(SIBLING_BIN) and os.path.getmtime(SIBLING_BIN) > max(os.path.getmtime(SIBLING_SRC), os.path.getmtime(LIB)):
        return SIBLING_BIN
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-Wall", SIBLING_SRC, "-L" + HERE, "-lkanzi_b200", "-ldl", "-Wl,-rpath,$ORIGIN/../../kanzi_b200", "-o", SIBLING_BIN])
    return SIBLING_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
