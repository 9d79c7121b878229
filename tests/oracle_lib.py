"""ctypes binding of oracle/libkzoracle.so (TEST INFRASTRUCTURE ONLY — never imported by kanzi_b200/)."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None

T = dict(NONE=0, BWT=1, BWTS=2, LZ=3, RLT=5, ZRLT=6, MTFT=7, RANK=8, ROLZ=11, ROLZX=12, SRT=13, LZP=14, LZX=16)
E = dict(NONE=0, HUFFMAN=1, FPAQ=2, RANGE=4, ANS0=5, ANS1=8)

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)


def build():
    so = os.path.join(ORACLE_DIR, "libkzoracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".hpp", ".cpp"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libkzoracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.kzo_entropy_encode.restype = C.c_int64
        L.kzo_entropy_encode.argtypes = [C.c_int, u8p, C.c_int32, u8p, C.c_int64, i64p]
        L.kzo_entropy_decode.restype = C.c_int32
        L.kzo_entropy_decode.argtypes = [C.c_int, u8p, C.c_int64, u8p, C.c_int32, i64p]
        L.kzo_transform.restype = C.c_int
        L.kzo_transform.argtypes = [C.c_int, C.c_int, i32p, u8p, C.c_int32, C.c_int32, u8p, C.c_int32, C.c_int32, i32p, i32p]
        L.kzo_transform_max_encoded_len.restype = C.c_int32
        L.kzo_transform_max_encoded_len.argtypes = [C.c_int, C.c_int32]
        L.kzo_sequence_forward.restype = C.c_int32
        L.kzo_sequence_forward.argtypes = [i32p, C.c_int, C.c_int32, C.c_int, u8p, C.c_int32, u8p, C.c_int32, i32p]
        L.kzo_bwt_forward.restype = C.c_int
        L.kzo_bwt_forward.argtypes = [u8p, C.c_int32, u8p, i32p]
        L.kzo_bwt_inverse.restype = C.c_int
        L.kzo_bwt_inverse.argtypes = [u8p, C.c_int32, u8p, i32p, C.c_int]
        L.kzo_compress_stream.restype = C.c_int64
        L.kzo_compress_stream.argtypes = [u8p, C.c_int64, i32p, C.c_int, C.c_int, C.c_int32, C.c_int64, C.c_int, u8p, C.c_int64]
        L.kzo_decompress_stream.restype = C.c_int64
        L.kzo_decompress_stream.argtypes = [u8p, C.c_int64, C.c_int, u8p, C.c_int64]
        L.kzo_encode_blocks_mt.restype = C.c_int32
        L.kzo_encode_blocks_mt.argtypes = [u8p, C.c_int64, i32p, C.c_int, C.c_int, C.c_int32, C.c_int, C.c_int, u8p, C.c_int64, i64p, i64p, C.c_int32]
        L.kzo_decode_blocks_mt.restype = C.c_int64
        L.kzo_decode_blocks_mt.argtypes = [u8p, i64p, i64p, C.c_int32, i32p, C.c_int, C.c_int, C.c_int32, C.c_int, C.c_int, u8p, C.c_int64]
        L.kzo_stream_header.restype = C.c_int32
        L.kzo_stream_header.argtypes = [i32p, C.c_int, C.c_int, C.c_int32, C.c_int64, u8p, C.c_int32]
        L.kzo_normalize_frequencies.restype = C.c_int32
        L.kzo_normalize_frequencies.argtypes = [i32p, i32p, C.c_int32, C.c_int32]
        L.kzo_xxhash32.restype = C.c_uint32
        L.kzo_xxhash32.argtypes = [u8p, C.c_int32, C.c_uint32]
        L.kzo_xxhash64.restype = C.c_uint64
        L.kzo_xxhash64.argtypes = [u8p, C.c_int32, C.c_uint64]
        L.kzo_expgolomb_signed.restype = C.c_int32
        L.kzo_expgolomb_signed.argtypes = [C.c_int8, C.POINTER(C.c_uint32)]
    return _LIB


def _u8(a):
    a = np.ascontiguousarray(np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray)) else a, dtype=np.uint8)
    return a, a.ctypes.data_as(u8p)


def _ids(names):
    arr = (C.c_int32 * 8)(*([T[n] for n in names] + [0] * (8 - len(names))))
    return arr, len(names)


def entropy_encode(kind, data):
    """-> (bytes, nbits)"""
    a, p = _u8(data)
    cap = 2 * len(a) + 4096
    out = np.zeros(cap, dtype=np.uint8)
    bits = C.c_int64(0)
    r = lib().kzo_entropy_encode(E[kind], p, len(a), out.ctypes.data_as(u8p), cap, C.byref(bits))
    if r != len(a):
        raise RuntimeError(f"oracle entropy_encode failed: {r}")
    return out[: (bits.value + 7) // 8].tobytes(), bits.value


def entropy_decode(kind, payload, nbits, n):
    a, p = _u8(payload)
    out = np.zeros(max(n, 1), dtype=np.uint8)
    used = C.c_int64(0)
    r = lib().kzo_entropy_decode(E[kind], p, nbits, out.ctypes.data_as(u8p), n, C.byref(used))
    return out[:n].tobytes(), r, used.value


def transform(kind, data, inverse=False, dst_cap=None, dst_len=None, src_cap=None, ctx=None):
    """One ByteTransform call. -> (ok, out_bytes, src_used, ctxv)"""
    a, _ = _u8(data)
    n = len(a)
    src_cap = n if src_cap is None else src_cap
    sa = np.zeros(src_cap, dtype=np.uint8)
    sa[:n] = a
    if dst_cap is None:
        dst_cap = max(lib().kzo_transform_max_encoded_len(T[kind], n), n) if not inverse else n
    dst_len = dst_cap if dst_len is None else dst_len
    out = np.zeros(max(dst_cap, 1), dtype=np.uint8)
    cv = (C.c_int32 * 6)(*(ctx or [7, max(n, 1024), n, 1, 0, 0]))
    su, du = C.c_int32(0), C.c_int32(0)
    r = lib().kzo_transform(T[kind], 1 if inverse else 0, cv, sa.ctypes.data_as(u8p), n, src_cap, out.ctypes.data_as(u8p), dst_len, dst_cap, C.byref(su), C.byref(du))
    return r, out[: max(du.value, 0)].tobytes(), su.value, list(cv)


def sequence_forward(names, data, block_size, bwt_bounds=1):
    a, p = _u8(data)
    ids, n = _ids(names)
    cap = len(a) + len(a) // 8 + 4096
    out = np.zeros(cap, dtype=np.uint8)
    sf = C.c_int32(0)
    r = lib().kzo_sequence_forward(ids, n, block_size, bwt_bounds, p, len(a), out.ctypes.data_as(u8p), cap, C.byref(sf))
    if r < 0:
        raise RuntimeError(f"oracle sequence_forward failed: {r}")
    return out[:r].tobytes(), sf.value & 0xFF


def bwt_forward(data):
    a, p = _u8(data)
    out = np.zeros(len(a), dtype=np.uint8)
    pi = (C.c_int32 * 8)()
    r = lib().kzo_bwt_forward(p, len(a), out.ctypes.data_as(u8p), pi)
    return r, out.tobytes(), list(pi)


def bwt_inverse(data, pis, algo=0):
    a, p = _u8(data)
    out = np.zeros(len(a), dtype=np.uint8)
    pi = (C.c_int32 * 8)(*pis)
    r = lib().kzo_bwt_inverse(p, len(a), out.ctypes.data_as(u8p), pi, algo)
    return r, out.tobytes()


def xxhash32(data, seed=0):
    a, p = _u8(data)
    return lib().kzo_xxhash32(p, len(a), seed)


def xxhash64(data, seed=0):
    """Kanzi's XXHash64 (not the published XXH64 for inputs of 32+ bytes: see oracle/kz_stream.hpp)"""
    a, p = _u8(data)
    return lib().kzo_xxhash64(p, len(a), seed)


def compress(data, transforms, entropy, block_size, input_size=None, bwt_bounds=1, checksum=0):
    a, p = _u8(data)
    ids, n = _ids(transforms)
    cap = len(a) + len(a) // 4 + (1 << 16)
    out = np.zeros(cap, dtype=np.uint8)
    isz = len(a) if input_size is None else input_size
    r = lib().kzo_compress_stream(p, len(a), ids, n, E[entropy], block_size, isz, bwt_bounds | (checksum << 8), out.ctypes.data_as(u8p), cap)
    if r < 0:
        raise RuntimeError(f"oracle compress failed: {r}")
    return out[:r].tobytes()


def decompress(stream, max_out, bwt_bounds=1):
    a, p = _u8(stream)
    out = np.zeros(max(max_out, 1), dtype=np.uint8)
    r = lib().kzo_decompress_stream(p, len(a), bwt_bounds, out.ctypes.data_as(u8p), max_out)
    if r < 0:
        raise RuntimeError(f"oracle decompress failed: {r}")
    return out[:r].tobytes()


def stream_header(transforms, entropy, block_size, input_size):
    ids, n = _ids(transforms)
    out = np.zeros(64, dtype=np.uint8)
    r = lib().kzo_stream_header(ids, n, E[entropy], block_size, input_size, out.ctypes.data_as(u8p), 64)
    return out[:r].tobytes()


def encode_blocks_mt(data, transforms, entropy, block_size, nthreads, bwt_bounds=1, out=None):
    """`out`: optional preallocated uint8 buffer (timing runs keep the allocation out of the timed region)."""
    a, p = _u8(data)
    ids, n = _ids(transforms)
    nb = (len(a) + block_size - 1) // block_size
    cap = len(a) + len(a) // 4 + (1 << 16) + 64 * nb
    if out is None or len(out) < cap:
        out = np.zeros(cap, dtype=np.uint8)
    cap = len(out)
    off = np.zeros(nb, dtype=np.int64)
    bits = np.zeros(nb, dtype=np.int64)
    r = lib().kzo_encode_blocks_mt(p, len(a), ids, n, E[entropy], block_size, bwt_bounds, nthreads, out.ctypes.data_as(u8p), cap,
                                   off.ctypes.data_as(i64p), bits.ctypes.data_as(i64p), nb)
    if r < 0:
        raise RuntimeError(f"oracle encode_blocks_mt failed: {r}")
    return out, off, bits


def decode_blocks_mt(recs, off, bits, transforms, entropy, block_size, nthreads, out_size, bwt_bounds=1, out=None):
    ids, n = _ids(transforms)
    if out is None or len(out) < out_size:
        out = np.zeros(out_size, dtype=np.uint8)
    r = lib().kzo_decode_blocks_mt(recs.ctypes.data_as(u8p), off.ctypes.data_as(i64p), bits.ctypes.data_as(i64p), len(off), ids, n,
                                   E[entropy], block_size, bwt_bounds, nthreads, out.ctypes.data_as(u8p), out_size)
    if r < 0:
        raise RuntimeError(f"oracle decode_blocks_mt failed: {r}")
    return out[:r]
