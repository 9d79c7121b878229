"""GPU parity of the sibling codecs added after the five BASELINE chains: LZP (K/transform/LZCodec.java:973-1287),
RLT (K/transform/RLT.java) and ROLZX (K/transform/ROLZCodec.java:1016-1770), against the oracle through the C ABI, bit for bit.  The same checks, torch-free, are what
tests/native/kzg_sibling_check.c runs (profiles/r02_sibling_check.log is its output on a B200)."""
import numpy as np
import pytest
import kanzi_b200 as K
import oracle_lib as O
import corpus
from kanzi_b200 import synth
from test_sibling_hostcheck import _inputs as lzp_inputs

pytestmark = pytest.mark.gpu


def first_diff(a, b):
    n = min(len(a), len(b))
    x = np.frombuffer(a[:n], dtype=np.uint8) != np.frombuffer(b[:n], dtype=np.uint8)
    return int(np.argmax(x)) if x.any() else n


def pasted_text(n, seed):
    """prose-like bytes with long passages pasted again further on (what LZP codes as matches) and flag bytes sprinkled in"""
    r = np.random.default_rng(seed)
    a = bytearray(synth.text(n, seed).tobytes())
    for _ in range(n // 3000):
        src, ln, dst = int(r.integers(0, n - 2000)), int(r.integers(70, 1900)), int(r.integers(0, n - 2000))
        a[dst:dst + ln] = a[src:src + ln]
    for _ in range(n // 5000):
        a[int(r.integers(0, n))] = int(r.choice([0xFC, 0xFE, 0xFF]))
    return bytes(a)


def test_lzp_transform_bit_exact():
    applied = 0
    cases = list(corpus.small_cases().values()) + lzp_inputs() + [pasted_text(1_000_003, 9)]
    for d in cases:
        cap = len(d) + len(d) // 64 + 1100
        ok_ref, ref, _, _ = O.transform("LZP", d, dst_cap=cap, ctx=[7, max(len(d), 1024), len(d), 1, 0, 0])
        kctx = {"blockSize": max(len(d), 1024), "size": len(d), "flags": 0}
        ok, got, used = K.transform_forward("LZP", d, kctx, dst_cap=cap)
        assert int(ok) == ok_ref, (len(d), ok, ok_ref)
        if not ok:
            continue
        applied += 1
        assert used == len(d) and got == ref, (len(d), len(got), len(ref), "first differing byte", first_diff(got, ref))
        ok2, back, used2 = K.transform_inverse("LZP", ref, {"blockSize": max(len(d), 1024), "flags": 0}, dst_cap=len(d))
        assert ok2 and back == d and used2 == len(ref), (len(d), first_diff(back, d))
        # LZPCodec.inverse bounds its output by the destination slice: one byte short is a refusal, as in the oracle
        if len(d) > 200:
            assert O.transform("LZP", ref, inverse=True, dst_cap=len(d) - 1, dst_len=len(d) - 1)[0] == 0
            assert not K.transform_inverse("LZP", ref, {"blockSize": max(len(d), 1024), "flags": 0}, dst_cap=len(d) - 1)[0]
    assert applied >= 40


def test_lzp_inverse_of_corrupt_streams_fails_like_the_oracle():
    r = np.random.default_rng(3)
    d = pasted_text(120_000, 4)
    ok, ref, _, _ = O.transform("LZP", d)
    assert ok == 1
    for k in range(24):
        bad = bytearray(ref)
        for _ in range(1 + k % 3):
            bad[int(r.integers(4, len(bad)))] = int(r.choice([0xFC, 0xFE, 0xFF, 0x00, int(r.integers(0, 256))]))
        o = O.transform("LZP", bytes(bad), inverse=True, dst_cap=len(d), dst_len=len(d))
        g = K.transform_inverse("LZP", bytes(bad), {"blockSize": len(d), "flags": 0}, dst_cap=len(d))
        assert bool(g[0]) == (o[0] == 1), k
        if g[0]:
            assert g[1] == o[1]
    assert K.transform_inverse("LZP", ref, {"blockSize": len(d), "flags": 0}, dst_cap=len(d))[1] == d      # no sticky error


@pytest.mark.parametrize("tr,ent,bs", [(["LZP"], "ANS0", 1 << 20), (["LZP", "ZRLT"], "HUFFMAN", 1 << 18), (["LZP"], "NONE", 1 << 16), (["ROLZ", "LZP"], "ANS0", 1 << 19)])
def test_lzp_streams_bit_exact(tr, ent, bs):
    d = pasted_text(2_500_000, 11) + synth.noise(150_000, 4).tobytes() + bytes(70000) + b"tail!"
    ref = O.compress(d, tr, ent, bs)
    got = K.compress(d, tr, ent, bs, flags=K.FLAG_BWT_ASREF)
    assert len(got) == len(ref) and got == ref, (tr, ent, len(got), len(ref), "first differing byte", first_diff(got, ref))
    assert K.decompress(ref, len(d) + 1024, flags=K.FLAG_BWT_ASREF) == d


# ---- RLT (K/transform/RLT.java) ------------------------------------------------------------------------------------------------
from test_sibling_hostcheck import rlt_inputs


@pytest.mark.parametrize("entropy", ["NONE", "FPAQ", "ANS1"])
def test_rlt_transform_bit_exact(entropy):
    """ctx["entropy"] decides the escape byte (RLT.java:101-107): the default 0xFB for NONE / ANS0 / HUFFMAN / RANGE, the block's rarest
    byte otherwise — and in that case RLT also classifies the block (DNA and BASE64 blocks are then left alone)."""
    eid = K.E[entropy]
    applied = 0
    for d in rlt_inputs() + list(corpus.small_cases().values()):
        cap = len(d) + 32 if len(d) <= 512 else len(d)
        ok_ref, ref, _, cv = O.transform("RLT", d, dst_cap=cap, dst_len=cap, ctx=[7, max(len(d), 1024), len(d), 1, 0, eid << 8])
        kctx = {"blockSize": max(len(d), 1024), "size": len(d), "flags": (eid + 1) << 8}
        ok, got, used = K.transform_forward("RLT", d, kctx, dst_cap=cap)
        assert int(ok) == ok_ref and kctx["dataType"] == cv[4], (len(d), ok, ok_ref, kctx["dataType"], cv[4])
        if not ok:
            continue
        applied += 1
        assert used == len(d) and got == ref, (len(d), len(got), len(ref), "first differing byte", first_diff(got, ref))
        ok2, back, used2 = K.transform_inverse("RLT", ref, {"blockSize": max(len(d), 1024), "flags": 0}, dst_cap=len(d))
        assert ok2 and back == d and used2 == len(ref), (len(d), first_diff(back, d))
        # one byte short: refused, except where the reference drops a trailing literal escape and still says yes (RLT.java:308-313)
        o = O.transform("RLT", ref, inverse=True, dst_cap=len(d) - 1, dst_len=len(d) - 1)
        g = K.transform_inverse("RLT", ref, {"blockSize": max(len(d), 1024), "flags": 0}, dst_cap=len(d) - 1)
        assert bool(g[0]) == (o[0] == 1) and (not g[0] or g[1] == o[1]), len(d)
    assert applied >= 40


def test_rlt_inverse_of_corrupt_streams_fails_like_the_oracle():
    r = np.random.default_rng(8)
    d = rlt_inputs()[0]
    ok, ref, _, _ = O.transform("RLT", d)
    assert ok == 1
    for k in range(24):
        bad = bytearray(ref)
        for _ in range(1 + k % 3):
            bad[int(r.integers(0, len(bad)))] = int(r.choice([ref[0], 0xFF, 0xE0, 0x00, int(r.integers(0, 256))]))
        o = O.transform("RLT", bytes(bad), inverse=True, dst_cap=len(d), dst_len=len(d))
        g = K.transform_inverse("RLT", bytes(bad), {"blockSize": len(d), "flags": 0}, dst_cap=len(d))
        assert bool(g[0]) == (o[0] == 1), k
        if g[0]:
            assert g[1] == o[1]


def runs_input(n, seed):
    r = np.random.default_rng(seed)
    vals = r.choice(np.array([0xFB, 0, 7, 0xFF, 65, 66], dtype=np.uint8), n // 20)
    lens = np.where(r.random(n // 20) < 0.02, r.integers(200, 90000, n // 20), r.integers(1, 12, n // 20))
    return bytes(np.repeat(vals, lens)[:n])


@pytest.mark.parametrize("tr,ent,bs", [(["RLT"], "ANS0", 1 << 18), (["RLT"], "FPAQ", 1 << 20), (["RLT", "LZP"], "HUFFMAN", 1 << 19), (["RLT"], "NONE", 1 << 16)])
def test_rlt_streams_bit_exact(tr, ent, bs):
    d = runs_input(2_000_000, 5) + synth.text(300_000, 3).tobytes() + bytes(70000) + b"tail!"
    ref = O.compress(d, tr, ent, bs)
    got = K.compress(d, tr, ent, bs, flags=K.FLAG_BWT_ASREF)
    assert len(got) == len(ref) and got == ref, (tr, ent, len(got), len(ref), "first differing byte", first_diff(got, ref))
    assert K.decompress(ref, len(d) + 1024, flags=K.FLAG_BWT_ASREF) == d


# ---- ROLZX = ROLZCodec2 (K/transform/ROLZCodec.java:1016-1428) --------------------------------------------------------------------
from test_sibling_hostcheck import rolzx_inputs


def test_rolzx_transform_bit_exact():
    applied = 0
    for d in rolzx_inputs():
        ok_ref, ref, _, cv = O.transform("ROLZX", d)
        kctx = {"blockSize": max(len(d), 1024), "size": len(d), "flags": 0}
        ok, got, used = K.transform_forward("ROLZX", d, kctx)
        assert int(ok) == ok_ref and kctx["dataType"] == cv[4], (len(d), ok, ok_ref, kctx["dataType"], cv[4])
        if not ok:
            continue
        applied += 1
        assert used == len(d) and got == ref, (len(d), len(got), len(ref), "first differing byte", first_diff(got, ref))
        ok2, back, used2 = K.transform_inverse("ROLZX", ref, {"blockSize": max(len(d), 1024), "flags": 0}, dst_cap=len(d))
        assert ok2 and back == d and used2 == len(ref), (len(d), first_diff(back, d))
        assert not K.transform_inverse("ROLZX", ref, {"blockSize": max(len(d), 1024), "flags": 0}, dst_cap=len(d) - 1)[0]
        assert not K.transform_inverse("ROLZX", ref[:len(ref) // 2], {"blockSize": max(len(d), 1024), "flags": 0}, dst_cap=len(d))[0]
    assert applied >= 35


def test_rolzx_inverse_of_corrupt_streams_fails_like_the_oracle():
    r = np.random.default_rng(12)
    d = synth.text(40000, 2).tobytes()
    ok, ref, _, _ = O.transform("ROLZX", d)
    assert ok == 1
    for k in range(24):
        bad = bytearray(ref)
        for _ in range(1 + k % 3):
            bad[int(r.integers(5 if k % 4 else 0, len(bad)))] ^= 1 << int(r.integers(0, 8))
        o = O.transform("ROLZX", bytes(bad), inverse=True, dst_cap=len(d), dst_len=len(d))
        g = K.transform_inverse("ROLZX", bytes(bad), {"blockSize": len(d), "flags": 0}, dst_cap=len(d))
        assert bool(g[0]) == (o[0] == 1), k
        if g[0]:
            assert g[1] == o[1]
    assert K.transform_inverse("ROLZX", ref, {"blockSize": len(d), "flags": 0}, dst_cap=len(d))[1] == d


def test_rolzx_block_of_two_chunks():
    """a block over 16 MiB: the ring table is cleared between the chunks, the coder and the counters carry over (ROLZCodec.java:1235-1283)"""
    n = (17 << 20) + 4321
    a = np.zeros(n, dtype=np.uint8)
    t = np.frombuffer(synth.text(20000, 6).tobytes(), dtype=np.uint8)
    for o in range(0, n - 64, 65536):
        a[o:o + 48] = t[(o >> 10) % 19000:(o >> 10) % 19000 + 48]
    d = a.tobytes()
    ok_ref, ref, _, _ = O.transform("ROLZX", d)
    ok, got, used = K.transform_forward("ROLZX", d, {"blockSize": n, "size": n, "flags": 0})
    assert ok_ref == 1 and ok and got == ref and used == n
    ok2, back, _ = K.transform_inverse("ROLZX", ref, {"blockSize": n, "flags": 0}, dst_cap=n)
    assert ok2 and back == d


@pytest.mark.parametrize("tr,ent,bs", [(["ROLZX"], "NONE", 1 << 18), (["ROLZX"], "ANS0", 1 << 20), (["RLT", "ROLZX"], "HUFFMAN", 1 << 19)])
def test_rolzx_streams_bit_exact(tr, ent, bs):
    d = (runs_input(1_500_000, 5) if "RLT" in tr else b"") + pasted_text(1_500_000, 11) + bytes(70000) + b"tail!"
    ref = O.compress(d, tr, ent, bs)
    got = K.compress(d, tr, ent, bs, flags=K.FLAG_BWT_ASREF)
    assert len(got) == len(ref) and got == ref, (tr, ent, len(got), len(ref), "first differing byte", first_diff(got, ref))
    assert K.decompress(ref, len(d) + 1024, flags=K.FLAG_BWT_ASREF) == d


def test_none_tokens_are_dropped_from_a_transform_list():
    """TransformFactory.getType (K/transform/TransformFactory.java:140-153) skips NONE tokens: the stream of "-t NONE+LZ" is the stream of "-t LZ"."""
    d = pasted_text(600_000, 21)
    ref = O.compress(d, ["LZ"], "ANS0", 1 << 18)
    assert O.compress(d, ["NONE", "LZ"], "ANS0", 1 << 18) == ref
    assert K.compress(d, ["NONE", "LZ"], "ANS0", 1 << 18, flags=K.FLAG_BWT_ASREF) == ref
    assert K.compress(d, ["LZ", "NONE"], "ANS0", 1 << 18, flags=K.FLAG_BWT_ASREF) == ref
    assert K.decompress(ref, len(d) + 1024, flags=K.FLAG_BWT_ASREF) == d
