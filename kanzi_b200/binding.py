"""ctypes binding of libkanzi_b200.so (include/kzg.h).  Fails loudly when the library is missing."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkanzi_b200.so")
_LIB = None

# ids: K/transform/TransformFactory.java:36-58, K/entropy/EntropyCodecFactory.java:38-47
T = dict(NONE=0, BWT=1, LZ=3, ZRLT=6, MTFT=7, RANK=8, ROLZ=11, SRT=13, LZP=14, LZX=16, RLT=5, ROLZX=12)
E = dict(NONE=0, HUFFMAN=1, FPAQ=2, ANS0=5, ANS1=8)
DT = dict(UNDEFINED=0, TEXT=1, MULTIMEDIA=2, EXE=3, NUMERIC=4, BASE64=5, DNA=6, BIN=7, UTF8=8, SMALL_ALPHABET=9)
FLAG_BWT_ASREF = 1
FLAG_XXH32 = 2          # per-block XXHash32 of the original bytes (-x 32)
FLAG_XXH64 = 4          # per-block XXHash64, Kanzi's variant (-x 64)
ERR_NO_DEVICE = 126

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)


class KzgCtx(C.Structure):
    _fields_ = [("bsVersion", C.c_int32), ("blockSize", C.c_int32), ("size", C.c_int32), ("jobs", C.c_int32),
                ("dataType", C.c_int32), ("flags", C.c_int32)]


class KzgError(RuntimeError):
    def __init__(self, code, what):
        super().__init__(f"{what}: kzg error {code}: {last_error()}")
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -m kanzi_b200.build` (nvcc, sm_100a). "
                              "kanzi_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.kzg_abi_version.restype = C.c_int
        L.kzg_device_count.restype = C.c_int
        L.kzg_set_device.restype = C.c_int
        L.kzg_set_device.argtypes = [C.c_int]
        L.kzg_last_error.restype = C.c_char_p
        L.kzg_launch_count.restype = C.c_int64
        L.kzg_launch_count.argtypes = [C.c_int]
        L.kzg_stream.restype = C.c_void_p
        for name in ("kzg_transform_forward", "kzg_transform_inverse"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.c_int, C.POINTER(KzgCtx), u8p, C.c_int32, u8p, C.c_int32, C.c_int32, i32p, i32p]
        L.kzg_transform_max_encoded_len.restype = C.c_int32
        L.kzg_transform_max_encoded_len.argtypes = [C.c_int, C.c_int32]
        L.kzg_bwt_forward.restype = C.c_int
        L.kzg_bwt_forward.argtypes = [u8p, C.c_int32, u8p, i32p]
        L.kzg_bwt_inverse.restype = C.c_int
        L.kzg_bwt_inverse.argtypes = [u8p, C.c_int32, u8p, i32p]
        L.kzg_entropy_encode.restype = C.c_int64
        L.kzg_entropy_encode.argtypes = [C.c_int, C.POINTER(KzgCtx), u8p, C.c_int32, u8p, C.c_int64, i64p]
        L.kzg_entropy_decode.restype = C.c_int32
        L.kzg_entropy_decode.argtypes = [C.c_int, C.POINTER(KzgCtx), u8p, C.c_int64, i64p, u8p, C.c_int32]
        L.kzg_compress.restype = C.c_int64
        L.kzg_compress.argtypes = [u8p, C.c_int64, i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, u8p, C.c_int64]
        L.kzg_decompress.restype = C.c_int64
        L.kzg_decompress.argtypes = [u8p, C.c_int64, C.c_int32, u8p, C.c_int64]
        L.kzg_compress_bound.restype = C.c_int64
        L.kzg_compress_bound.argtypes = [C.c_int64, C.c_int32]
        L.kzg_compress_dev.restype = C.c_int64
        L.kzg_compress_dev.argtypes = [C.c_void_p, C.c_int64, i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                       C.POINTER(C.c_float)]
        L.kzg_decompress_dev.restype = C.c_int64
        L.kzg_decompress_dev.argtypes = [C.c_void_p, C.c_int64, u8p, C.c_int32, C.c_void_p, C.c_int64, C.POINTER(C.c_float)]
        L.kzg_last_block_bits.restype = C.c_int32
        L.kzg_last_block_bits.argtypes = [i64p, C.c_int32]
        L.kzg_stream_index.restype = C.c_int32
        L.kzg_stream_index.argtypes = [u8p, C.c_int64, i64p, i64p, C.c_int32, i64p]
        _LIB = L
    return _LIB


def device_count():
    return lib().kzg_device_count()


def set_device(d):
    r = lib().kzg_set_device(d)
    if r < 0:
        raise KzgError(r, "kzg_set_device")


def last_error():
    return (lib().kzg_last_error() or b"").decode("utf-8", "replace")


def launch_count(reset=False):
    return lib().kzg_launch_count(1 if reset else 0)


def _u8(a):
    if isinstance(a, (bytes, bytearray, memoryview)):
        a = np.frombuffer(a, dtype=np.uint8)
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(u8p)


def _ctx(ctx):
    c = KzgCtx(7, 0, 0, 1, 0, 0)
    for k, v in (ctx or {}).items():
        setattr(c, k, v)
    return c


def transform_max_encoded_len(kind, n):
    return lib().kzg_transform_max_encoded_len(T[kind] if isinstance(kind, str) else kind, n)


def _transform(fn, kind, data, ctx, dst_len, dst_cap):
    a, p = _u8(data)
    t = T[kind] if isinstance(kind, str) else kind
    n = len(a)
    if dst_cap is None:
        dst_cap = max(transform_max_encoded_len(t, n), n, 1)
    if dst_len is None:
        dst_len = dst_cap
    out = np.zeros(max(dst_cap, 1), dtype=np.uint8)
    c = _ctx(ctx)
    su, du = C.c_int32(0), C.c_int32(0)
    r = fn(t, C.byref(c), p, n, out.ctypes.data_as(u8p), dst_len, dst_cap, C.byref(su), C.byref(du))
    if r < 0:
        raise KzgError(r, "kzg_transform")
    if ctx is not None:
        ctx["dataType"] = c.dataType
    return bool(r), out[: du.value].tobytes(), su.value


def transform_forward(kind, data, ctx=None, dst_len=None, dst_cap=None):
    """ByteTransform.forward on (src: length len(data), index 0) -> (ok, produced bytes, src consumed)."""
    return _transform(lib().kzg_transform_forward, kind, data, ctx, dst_len, dst_cap)


def transform_inverse(kind, data, ctx=None, dst_len=None, dst_cap=None):
    return _transform(lib().kzg_transform_inverse, kind, data, ctx, dst_len, dst_cap)


def bwt_forward(data):
    a, p = _u8(data)
    out = np.zeros(max(len(a), 1), dtype=np.uint8)
    pi = (C.c_int32 * 8)()
    r = lib().kzg_bwt_forward(p, len(a), out.ctypes.data_as(u8p), pi)
    if r < 0:
        raise KzgError(r, "kzg_bwt_forward")
    return bool(r), out[: len(a)].tobytes(), list(pi)


def bwt_inverse(data, primary_indexes):
    a, p = _u8(data)
    out = np.zeros(max(len(a), 1), dtype=np.uint8)
    pi = (C.c_int32 * 8)(*primary_indexes)
    r = lib().kzg_bwt_inverse(p, len(a), out.ctypes.data_as(u8p), pi)
    if r < 0:
        raise KzgError(r, "kzg_bwt_inverse")
    return bool(r), out[: len(a)].tobytes()


def entropy_encode(kind, data, ctx=None):
    """EntropyEncoder.encode + dispose -> (payload bytes, bit length)."""
    a, p = _u8(data)
    cap = 2 * len(a) + (300 << 10)
    out = np.zeros(cap, dtype=np.uint8)
    bits = C.c_int64(0)
    c = _ctx(ctx)
    r = lib().kzg_entropy_encode(E[kind] if isinstance(kind, str) else kind, C.byref(c), p, len(a), out.ctypes.data_as(u8p), cap, C.byref(bits))
    if r != len(a):
        raise KzgError(r, "kzg_entropy_encode")
    return out[: (bits.value + 7) // 8].tobytes(), bits.value


def entropy_decode(kind, payload, nbits, n, ctx=None):
    """EntropyDecoder.decode -> (bytes, return value, bits consumed)."""
    a, p = _u8(payload)
    out = np.zeros(max(n, 1), dtype=np.uint8)
    used = C.c_int64(0)
    c = _ctx(ctx)
    r = lib().kzg_entropy_decode(E[kind] if isinstance(kind, str) else kind, C.byref(c), p, nbits, C.byref(used), out.ctypes.data_as(u8p), n)
    if r < 0:
        raise KzgError(r, "kzg_entropy_decode")
    return out[:n].tobytes(), r, used.value


def _ids(transforms):
    ids = [T[t] if isinstance(t, str) else t for t in transforms]
    return (C.c_int32 * 8)(*(ids + [0] * (8 - len(ids)))), len(ids)


def compress_bound(n, block_size):
    return lib().kzg_compress_bound(n, block_size)


def compress(data, transforms, entropy, block_size, flags=FLAG_BWT_ASREF):
    """Whole .knz stream (CompressedOutputStream semantics) from host bytes."""
    a, p = _u8(data)
    ids, n = _ids(transforms)
    cap = compress_bound(len(a), block_size)
    out = np.zeros(cap, dtype=np.uint8)
    r = lib().kzg_compress(p, len(a), ids, n, E[entropy] if isinstance(entropy, str) else entropy, block_size, flags, out.ctypes.data_as(u8p), cap)
    if r < 0:
        raise KzgError(r, "kzg_compress")
    return out[:r].tobytes()


def decompress(stream, max_out, flags=FLAG_BWT_ASREF):
    a, p = _u8(stream)
    out = np.zeros(max(max_out, 1), dtype=np.uint8)
    r = lib().kzg_decompress(p, len(a), flags, out.ctypes.data_as(u8p), max_out)
    if r < 0:
        raise KzgError(r, "kzg_decompress")
    return out[:r].tobytes()


def last_block_bits():
    """Record bit lengths (5 + lw + written) of the blocks of the calling thread's last compress call (block sharding)."""
    n = lib().kzg_last_block_bits(None, 0)
    out = np.zeros(max(n, 1), dtype=np.int64)
    lib().kzg_last_block_bits(out.ctypes.data_as(i64p), n)
    return out[:n]


def stream_index(stream):
    """Host-side walk of a .knz: (header bits, record bit offsets, record bit lengths)."""
    a, p = _u8(stream)
    hb = C.c_int64(0)
    n = lib().kzg_stream_index(p, len(a), None, None, 0, C.byref(hb))
    if n < 0:
        raise KzgError(n, "kzg_stream_index")
    off = np.zeros(max(n, 1), dtype=np.int64)
    bits = np.zeros(max(n, 1), dtype=np.int64)
    lib().kzg_stream_index(p, len(a), off.ctypes.data_as(i64p), bits.ctypes.data_as(i64p), n, C.byref(hb))
    return hb.value, off[:n], bits[:n]
