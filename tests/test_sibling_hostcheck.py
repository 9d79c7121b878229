"""The loops lane 0 of `lzp_kernel` and `rlt_kernel` runs (kanzi_b200/csrc/lzp_core.cuh: grouped table lookups, 8-byte probes and copies;
rlt_core.cuh) are plain C++ when no CUDA compiler is looking; tests/native/sibling_hostcheck.cpp compiles those same headers for the host
and this file holds them against the oracle's restatements (K/transform/LZCodec.java:973-1287, K/transform/RLT.java) on seeded inputs.  What it proves: the reordering the
device code does (four lookups at a time with in-group forwarding, literal-only table reads in the inverse) keeps the serial
semantics.  What it cannot prove: anything about the launch, the scratch carve or device memory; that is `-m gpu` territory
(tests/test_gpu_siblings.py).  Test infrastructure only: the product library exports none of this."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "sibling_hostcheck.cpp")
LIB = os.path.join(HERE, "native", "libsibling_hostcheck.so")
CORES = [os.path.join(HERE, "..", "kanzi_b200", "csrc", f) for f in ("lzp_core.cuh", "rlt_core.cuh", "rolzx_core.cuh")]
u8p = C.POINTER(C.c_uint8)


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max([os.path.getmtime(SRC)] + [os.path.getmtime(c) for c in CORES]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", SRC, "-o", LIB])
    L = C.CDLL(LIB)
    L.lzp_host_forward.argtypes = [u8p, C.c_int, u8p, C.POINTER(C.c_int)]
    L.lzp_host_inverse.argtypes = [u8p, C.c_int, u8p, C.c_int, C.POINTER(C.c_int)]
    L.rlt_host_forward.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.rlt_host_inverse.argtypes = [u8p, C.c_int, u8p, C.c_int, C.POINTER(C.c_int)]
    L.rolzx_host_forward.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.rolzx_host_inverse.argtypes = [u8p, C.c_int, u8p, C.c_int, C.POINTER(C.c_int)]
    return L


_libc = C.CDLL(None, use_errno=True)
_libc.mmap.restype = C.c_void_p
_libc.mmap.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_long]
_libc.mprotect.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
_libc.munmap.argtypes = [C.c_void_p, C.c_size_t]
PAGE = 4096


class _Guarded:
    """`data` placed so that its last byte is the last byte of a mapping whose next page is unreadable: one byte read past the end
    of the input kills the test process instead of going unnoticed."""
    def __init__(self, data):
        n = len(data)
        self.size = (n + PAGE - 1) // PAGE * PAGE + PAGE
        self.base = _libc.mmap(None, self.size, 3, 0x22, -1, 0)      # PROT_READ|PROT_WRITE, MAP_PRIVATE|MAP_ANONYMOUS
        assert self.base not in (None, C.c_void_p(-1).value)
        assert _libc.mprotect(self.base + self.size - PAGE, PAGE, 0) == 0
        self.addr = self.base + self.size - PAGE - n
        C.memmove(self.addr, bytes(data), n)
        self.ptr = C.cast(self.addr, u8p)

    def __del__(self):
        _libc.munmap(self.base, self.size)


def _fwd(L, d):
    a = np.frombuffer(d, dtype=np.uint8)
    g = _Guarded(d)
    dst = np.full(len(a) + 64, 0xA5, dtype=np.uint8)
    n = C.c_int(0)
    ok = L.lzp_host_forward(g.ptr, len(a), dst.ctypes.data_as(u8p), C.byref(n))
    assert (dst[len(a) - (len(a) >> 6):] == 0xA5).all()               # nothing written at or beyond dstEnd
    return ok, dst[:n.value].tobytes()


def _inv(L, s, dst_end):
    a = np.frombuffer(s, dtype=np.uint8)
    g = _Guarded(s)
    dst = np.full(dst_end + 64, 0xA5, dtype=np.uint8)
    n = C.c_int(0)
    ok = L.lzp_host_inverse(g.ptr, len(a), dst.ctypes.data_as(u8p), dst_end, C.byref(n))
    assert (dst[dst_end:] == 0xA5).all()
    return ok, dst[:n.value].tobytes()


def _inputs():
    from kanzi_b200 import synth
    r = np.random.default_rng(77)
    t = synth.text(40000, 5).tobytes()
    noise = bytes(r.integers(0, 256, 4000, dtype=np.uint8))
    flags = bytes(r.choice(np.array([0xFC, 0xFE, 0xFF, 0x41], dtype=np.uint8), 3000))
    out = [t + t[500:9000] + t[:6000], noise + noise + flags + noise[:1500] + flags, bytes(5000), (b"0123456789abcdef" * 7 + b"\xfc") * 300,
           synth.records(30000, 7).tobytes(), noise, b"\xfc" * 700, t[:127], t[:128], t[:129], t[:191], t[:192], t[:200], (noise[:300] + b"\xfc\xfc") * 40,
           synth.exe_like(30000, 6).tobytes() * 2, bytes(70000), (t[:1000] + noise[:7]) * 60]
    for k in range(60):                                               # seeded fuzz: phrases re-pasted at random, flag bytes sprinkled in, ragged lengths
        n = int(r.integers(128, 6000))
        base = bytearray(r.integers(0, int(r.choice([2, 4, 16, 256])), n, dtype=np.uint8).tobytes())
        for _ in range(int(r.integers(0, 12))):
            a, ln = int(r.integers(0, n)), int(r.integers(1, 700))
            b = int(r.integers(0, n))
            seg = bytes(base[a:a + ln])
            base[b:b + len(seg)] = seg
        for _ in range(int(r.integers(0, 30))):
            base[int(r.integers(0, len(base)))] = int(r.choice([0xFC, 0xFE, 0xFF]))
        out.append(bytes(base[:max(n, 128)]))
    return out


def test_device_loops_on_the_host_match_the_oracle(host):
    applied = 0
    for d in _inputs():
        ok_ref, ref, _, _ = O.transform("LZP", d)
        if len(d) < 128:
            assert ok_ref == 0
            continue                                                  # the kernel refuses these before the loops run
        ok, got = _fwd(host, d)
        assert ok == ok_ref, (len(d), ok, ok_ref)
        if not ok:
            continue
        applied += 1
        assert got == ref, (len(d), len(got), len(ref))
        assert _inv(host, ref, len(d)) == (1, d)
        assert _inv(host, ref, len(d) + 1000) == (1, d)
        assert _inv(host, ref, len(d) - 1)[0] == 0 == O.transform("LZP", ref, inverse=True, dst_cap=len(d) - 1, dst_len=len(d) - 1)[0]
        for cut in (len(ref) // 2, len(ref) - 1, 5):
            o = O.transform("LZP", ref[:cut], inverse=True, dst_cap=len(d), dst_len=len(d))
            h = _inv(host, ref[:cut], len(d))
            assert o[0] == h[0] and (not h[0] or o[1] == h[1]), (len(d), cut)
    assert applied >= 40


def test_corrupt_streams_fail_alike(host):
    from kanzi_b200 import synth
    r = np.random.default_rng(3)
    t = synth.text(30000, 2).tobytes()
    d = t + t[2000:20000]
    ok, ref, _, _ = O.transform("LZP", d)
    assert ok == 1
    for k in range(40):
        bad = bytearray(ref)
        for _ in range(1 + k % 3):
            bad[int(r.integers(4, len(bad)))] = int(r.choice([0xFC, 0xFE, 0xFF, 0x00, int(r.integers(0, 256))]))
        o = O.transform("LZP", bytes(bad), inverse=True, dst_cap=len(d), dst_len=len(d))
        h = _inv(host, bytes(bad), len(d))
        assert o[0] == h[0], k
        if h[0]:
            assert o[1] == h[1]


# ---- RLT ------------------------------------------------------------------------------------------------------------------------
def _rlt_fwd(L, d, dst_end, best, dt=0):
    g = _Guarded(d)
    dst = np.full(dst_end + 64, 0xA5, dtype=np.uint8)
    n, t = C.c_int(0), C.c_int(dt)
    ok = L.rlt_host_forward(g.ptr, len(d), dst.ctypes.data_as(u8p), dst_end, best, C.byref(t), C.byref(n))
    assert (dst[dst_end:] == 0xA5).all()
    return ok, dst[:n.value].tobytes(), t.value


def _rlt_inv(L, s, dst_end):
    g = _Guarded(s)
    dst = np.full(dst_end + 64, 0xA5, dtype=np.uint8)
    n = C.c_int(0)
    ok = L.rlt_host_inverse(g.ptr, len(s), dst.ctypes.data_as(u8p), dst_end, C.byref(n))
    assert (dst[dst_end:] == 0xA5).all()
    return ok, dst[:n.value].tobytes()


def rlt_inputs():
    from test_oracle_crosscheck import _rlt_cases
    r = np.random.default_rng(91)
    out = list(_rlt_cases())
    for k in range(60):                                               # seeded fuzz: run lengths around every coding threshold, escapes inside and at the ends
        parts = []
        for _ in range(int(r.integers(2, 40))):
            v = int(r.choice([0xFB, 0, 1, 0xFF, int(r.integers(0, 256))]))
            ln = int(r.choice([1, 2, 3, 4, 5, 226, 227, 228, 7938, 7939, 7940, int(r.integers(1, 300))]))
            parts.append(bytes([v]) * ln)
        d = b"".join(parts)
        out.append(d if len(d) >= 16 else d + bytes(16))
    return out


@pytest.mark.parametrize("entropy,eid", [("NONE", 0), ("ANS1", 8)])
def test_rlt_device_loops_on_the_host_match_the_oracle(host, entropy, eid):
    applied = 0
    best = 0 if entropy == "NONE" else 1
    for d in rlt_inputs():
        if len(d) < 16:
            continue
        cap = len(d) + 32 if len(d) <= 512 else len(d)
        ok_ref, ref, _, cv = O.transform("RLT", d, ctx=[7, max(len(d), 1024), len(d), 1, 0, eid << 8])
        ok, got, dt = _rlt_fwd(host, d, cap, best)
        assert (ok, dt) == (ok_ref, cv[4]), (len(d), ok, ok_ref, dt, cv[4])
        if not ok:
            continue
        applied += 1
        assert got == ref, (len(d), len(got), len(ref))
        assert _rlt_inv(host, ref, len(d)) == (1, d)
        for cap2 in (len(d) - 1, len(d) + 77):
            o = O.transform("RLT", ref, inverse=True, dst_cap=cap2, dst_len=cap2)
            h = _rlt_inv(host, ref, cap2)
            assert max(o[0], 0) == h[0] and (not h[0] or o[1] == h[1]), (len(d), cap2)
        for cut in (len(ref) // 2, len(ref) - 1, 3, 2):
            o = O.transform("RLT", ref[:cut], inverse=True, dst_cap=len(d), dst_len=len(d))
            h = _rlt_inv(host, ref[:cut], len(d))
            assert max(o[0], 0) == h[0] and (not h[0] or o[1] == h[1]), (len(d), "cut", cut)
    assert applied >= 40


def test_rlt_corrupt_streams_fail_alike(host):
    r = np.random.default_rng(8)
    d = rlt_inputs()[0]
    ok, ref, _, _ = O.transform("RLT", d)
    assert ok == 1
    for k in range(60):
        bad = bytearray(ref)
        for _ in range(1 + k % 3):
            bad[int(r.integers(0, len(bad)))] = int(r.choice([ref[0], 0xFF, 0xE0, 0x00, int(r.integers(0, 256))]))
        o = O.transform("RLT", bytes(bad), inverse=True, dst_cap=len(d), dst_len=len(d))
        h = _rlt_inv(host, bytes(bad), len(d))
        assert max(o[0], 0) == h[0], k
        if h[0]:
            assert o[1] == h[1]


# ---- ROLZX ----------------------------------------------------------------------------------------------------------------------
def rolzx_inputs():
    from kanzi_b200 import synth
    from test_oracle_crosscheck import _rolzx_cases
    r = np.random.default_rng(71)
    t = synth.text(300_000, 8).tobytes()
    out = list(_rolzx_cases()) + [t, t[:70000] + t[1000:40000], synth.exe_like(200_000, 3).tobytes(), synth.records(150_000, 5).tobytes(), bytes(100_000),
                                  bytes(r.integers(0, 256, 50_000, dtype=np.uint8)), (b"ACGTTGCA" * 9 + b"N") * 800, t[:64], t[:65], t[:71], t[:72], t[:73]]
    for k in range(25):
        n = int(r.integers(64, 20000))
        base = bytearray(r.integers(0, int(r.choice([2, 4, 16, 256])), n, dtype=np.uint8).tobytes())
        for _ in range(int(r.integers(0, 20))):
            a, ln, b = int(r.integers(0, n)), int(r.integers(3, 400)), int(r.integers(0, n))
            seg = bytes(base[a:a + ln])
            base[b:b + len(seg)] = seg
        out.append(bytes(base[:n]))
    return out


def _rolzx_fwd(L, d, dt):
    g = _Guarded(d)
    cap = len(d) + 1024 if len(d) <= 16384 else len(d) + len(d) // 32
    dst = np.full(cap + 64, 0xA5, dtype=np.uint8)
    n = C.c_int(0)
    ok = L.rolzx_host_forward(g.ptr, len(d), dst.ctypes.data_as(u8p), cap, dt, C.byref(n))
    assert (dst[cap:] == 0xA5).all()
    return ok, dst[:n.value].tobytes()


def _rolzx_inv(L, s, dst_len):
    g = _Guarded(s)
    dst = np.full(dst_len + 64, 0xA5, dtype=np.uint8)
    n = C.c_int(0)
    ok = L.rolzx_host_inverse(g.ptr, len(s), dst.ctypes.data_as(u8p), dst_len, C.byref(n))
    assert (dst[dst_len:] == 0xA5).all()
    return ok, dst[:n.value].tobytes()


def test_rolzx_device_loops_on_the_host_match_the_oracle(host):
    applied = 0
    for d in rolzx_inputs():
        ok_ref, ref, _, cv = O.transform("ROLZX", d)
        if len(d) < 64:
            assert ok_ref == 0
            continue
        ok, got = _rolzx_fwd(host, d, cv[4])                      # (the kernel's histogram step decides the type; here the oracle's answer)
        assert ok == ok_ref, (len(d), ok, ok_ref)
        if ok != 1:
            continue
        applied += 1
        assert got == ref, (len(d), len(got), len(ref))
        assert _rolzx_inv(host, ref, len(d)) == (1, d)
        assert _rolzx_inv(host, ref, len(d) + 500) == (1, d)
        assert _rolzx_inv(host, ref, len(d) - 1)[0] == 0
        for cut in (len(ref) // 2, len(ref) - 4, 13):
            o = O.transform("ROLZX", ref[:cut], inverse=True, dst_cap=len(d), dst_len=len(d))
            h = _rolzx_inv(host, ref[:cut], len(d))
            assert max(o[0], 0) == h[0], (len(d), cut)
    assert applied >= 35


def test_rolzx_corrupt_streams_fail_alike(host):
    from kanzi_b200 import synth
    r = np.random.default_rng(12)
    d = synth.text(40000, 2).tobytes()
    ok, ref, _, _ = O.transform("ROLZX", d)
    assert ok == 1
    for k in range(40):
        bad = bytearray(ref)
        for _ in range(1 + k % 3):
            bad[int(r.integers(5 if k % 4 else 0, len(bad)))] ^= 1 << int(r.integers(0, 8))
        o = O.transform("ROLZX", bytes(bad), inverse=True, dst_cap=len(d), dst_len=len(d))
        h = _rolzx_inv(host, bytes(bad), len(d))
        assert max(o[0], 0) == h[0], k
        if h[0]:
            assert o[1] == h[1]
