"""Differential fuzz of the match-search codecs (LZ, LZX, ROLZ) and the exactness fallback of the LZ forward:
seeded adversarial inputs, CUDA path vs the oracle, bit for bit (VERDICT r01 "next round" 1b/1c).

Input families (every one seeded, >= 200 inputs per codec):
  * periodic data, period 1 .. 70 000 (long matches, repeat offsets, overlapping copies);
  * runs longer than MAX_MATCH = 65 793 (K/transform/LZCodec.java:452-456 clamps a match by moving its start);
  * blocks of 262 143 / 262 144 / 262 145 + 18 bytes (the maxDist switch, LZCodec.java:340: srcEnd < 4 * MAX_DISTANCE1);
  * 2-symbol and DNA alphabets (ROLZ / LZ dataType sniffing, minMatch 6), hash-collision-heavy data (few distinct 5-grams);
  * sparse data with planted repeats (the miss acceleration jumps over most positions).
tkBuf overflow (SURVEY E-3) needs more than count/5 tokens, i.e. matches of exactly minMatch = 4 bytes with no literal
between them, which the 5-byte hash (LZCodec.java:909-910) only yields on hash collisions: not constructible; documented, not tested."""
import os
import numpy as np
import pytest
import kanzi_b200 as K
import oracle_lib as O
import corpus
from kanzi_b200 import synth

pytestmark = pytest.mark.gpu


def first_diff(a, b):
    n = min(len(a), len(b))
    x = np.frombuffer(a[:n], dtype=np.uint8) != np.frombuffer(b[:n], dtype=np.uint8)
    return int(np.argmax(x)) if x.any() else n


def periodic(n, period, seed, noise_every=0):
    r = np.random.default_rng(seed)
    base = r.integers(0, 256, period, dtype=np.uint8)
    d = np.tile(base, n // period + 1)[:n].copy()
    if noise_every:
        at = np.arange(noise_every, n, noise_every)
        d[at] = r.integers(0, 256, len(at), dtype=np.uint8)
    return d.tobytes()


def long_runs(n, seed):
    r = np.random.default_rng(seed)
    parts, total = [], 0
    while total < n:
        kind = int(r.integers(0, 3))
        if kind == 0:
            ln = int(r.integers(65000, 140000))
            parts.append(np.full(ln, int(r.integers(0, 256)), dtype=np.uint8))
        elif kind == 1:
            ln = int(r.integers(100, 3000))
            parts.append(r.integers(0, 256, ln, dtype=np.uint8))
        else:
            ln = int(r.integers(66000, 90000))
            parts.append(np.tile(r.integers(0, 256, int(r.integers(2, 9)), dtype=np.uint8), ln)[:ln])
        total += ln
    return np.concatenate(parts)[:n].tobytes()


def small_alphabet(n, seed, syms):
    r = np.random.default_rng(seed)
    return r.choice(np.frombuffer(syms, dtype=np.uint8), n).tobytes()


def few_grams(n, seed, nwords=24, wlen=5):
    """few distinct 5-grams: every hash bucket is hit over and over, match candidates everywhere"""
    r = np.random.default_rng(seed)
    words = r.integers(0, 256, (nwords, wlen), dtype=np.uint8)
    return words[r.integers(0, nwords, n // wlen + 1)].reshape(-1)[:n].tobytes()


def fuzz_inputs(kind):
    """-> list of (name, bytes); deterministic"""
    out = []
    periods = [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 255, 256, 257, 1000, 4095, 4096, 4097, 16383, 16384, 16385,
               32767, 32768, 65533, 65534, 65535, 65536, 65537, 65792, 65793, 65794, 69999, 70000]
    for i, p in enumerate(periods):
        n = max(3 * p + 1000 + 37 * i, 40000 + 1111 * i)
        out.append((f"per{p}", periodic(n, p, 1000 + i)))
        out.append((f"per{p}n", periodic(n, p, 2000 + i, noise_every=997 + 13 * i)))
    for i in range(24):
        out.append((f"runs{i}", long_runs(150000 + 9973 * i, 3000 + i)))
    for i, n in enumerate([262143 + 18, 262144 + 18, 262145 + 18, 262143, 262144, 262145, 262161, 262163]):
        out.append((f"edge{n}", synth.text(n, 4000 + i).tobytes()))
        out.append((f"edgeP{n}", periodic(n, 70001 + i, 4100 + i)))
        out.append((f"edgeX{n}", synth.exe_like(n, 4200 + i).tobytes()))
    for i in range(16):
        out.append((f"two{i}", small_alphabet(30000 + 3331 * i, 5000 + i, bytes([3 + i, 200 - i]))))
        out.append((f"dna{i}", small_alphabet(50000 + 7777 * i, 5100 + i, b"ACGT")))
        out.append((f"dnaN{i}", small_alphabet(40000 + 5555 * i, 5200 + i, b"ACGTN\n")))
    for i in range(20):
        out.append((f"grams{i}", few_grams(60000 + 4001 * i, 6000 + i, nwords=4 + 3 * i, wlen=4 + (i % 4))))
    for i in range(14):
        out.append((f"sparse{i}", corpus.sparse_with_repeats(120000 + 10007 * i, 7000 + i, pcm=bool(i & 1), every=3000 + 500 * i, length=200 + 60 * i)))
    for i in range(14):
        n = 50000 + 12347 * i
        out.append((f"mix{i}", (synth.text(n // 2, 8000 + i).tobytes() + synth.noise(n // 4, 8100 + i).tobytes() + synth.records(n // 4, 8200 + i).tobytes())))
    assert len(out) >= 200, len(out)
    return out


@pytest.mark.timeout(1800)
@pytest.mark.parametrize("tr", ["LZ", "LZX", "ROLZ"])
def test_fuzz_match_codecs_bit_exact(tr):
    bad = []
    applied = 0
    for name, d in fuzz_inputs(tr):
        cap = len(d) + len(d) // 64 + 1100
        octx = [7, max(len(d), 1024), len(d), 1, 0, 0]
        ok_ref, ref, _, octx_out = O.transform(tr, d, dst_cap=cap, ctx=octx)
        kctx = {"blockSize": max(len(d), 1024), "size": len(d), "flags": 0}
        ok, got, used = K.transform_forward(tr, d, kctx, dst_cap=cap)
        if int(ok) != ok_ref:
            bad.append((name, "result", ok, ok_ref))
            continue
        if not ok:
            continue
        applied += 1
        if got != ref:
            bad.append((name, "forward", len(got), len(ref), first_diff(got, ref)))
            continue
        ok2, back, _ = K.transform_inverse(tr, ref, {"blockSize": max(len(d), 1024), "flags": 0}, dst_cap=len(d) + 512)
        if not ok2 or back != d:
            bad.append((name, "inverse", ok2, first_diff(back, d)))
    assert not bad, (tr, len(bad), bad[:8])
    assert applied >= 150, applied


def _lz_stream_input():
    return (synth.text(1_200_000, 51).tobytes() + corpus.sparse_with_repeats(700_000, 52) + synth.exe_like(600_000, 53).tobytes()
            + long_runs(400_000, 54) + periodic(300_000, 65793, 55) + corpus.sparse_with_repeats(500_000, 56, pcm=True) + synth.records(400_001, 57).tobytes())


@pytest.mark.timeout(900)
@pytest.mark.parametrize("knob", [("KZG_DEBUG", "8"), ("KZG_LZ_MAXROUNDS", "1"), ("KZG_LZ_MAXROUNDS", "2"), ("KZG_LZ_GROUPS", "1"), ("KZG_LZ_GROUPS", "3")])
def test_lz_forward_fallback_paths_bit_exact(knob):
    """The exactness fallback of the LZ forward (lzf_walk_kernel: a block whose fixed point has not closed after maxRounds rounds
    is parsed by the one-warp serial walker) is never reached on ordinary data with 8 rounds; KZG_DEBUG=8 sends every block
    through it, KZG_LZ_MAXROUNDS=1/2 sends the blocks that still have open marks after 1 / 2 rounds.  Same bytes as the oracle."""
    d = _lz_stream_input()
    name, val = knob
    old = os.environ.get(name)
    os.environ[name] = val
    try:
        for tr, ent, bs in ((["LZ"], "ANS0", 1 << 18), (["LZX"], "NONE", 1 << 19)):
            ref = O.compress(d, tr, ent, bs)
            got = K.compress(d, tr, ent, bs)
            assert len(got) == len(ref) and got == ref, (knob, tr, ent, len(got), len(ref), "first differing byte", first_diff(got, ref))
    finally:
        if old is None:
            del os.environ[name]
        else:
            os.environ[name] = old
    assert K.decompress(ref, len(d) + 1024) == d


@pytest.mark.timeout(900)
def test_many_seeds_per_chain():
    """every chain sees many seeds, not one or two."""
    chains = [(["LZ"], "ANS0", 1 << 17), (["ROLZ"], "ANS0", 1 << 17), (["BWT", "RANK", "ZRLT"], "ANS1", 1 << 17), (["BWT", "SRT", "ZRLT"], "FPAQ", 1 << 17),
              (["NONE"], "HUFFMAN", 1 << 16), (["LZX"], "HUFFMAN", 1 << 17)]
    gens = [synth.text, synth.markup, synth.records, synth.exe_like, synth.pcm_like, lambda n, s: synth.skewed(n, s, 2.5)]
    for ci, (tr, ent, bs) in enumerate(chains):
        for seed in range(10):
            g = gens[(seed + ci) % len(gens)]
            d = g(150_000 + 20_011 * seed, 900 + 17 * seed + ci).tobytes()
            for flags in ((K.FLAG_BWT_ASREF, 0) if "BWT" in tr and seed < 3 else (K.FLAG_BWT_ASREF,)):
                ref = O.compress(d, tr, ent, bs, bwt_bounds=1 if flags else 0)
                got = K.compress(d, tr, ent, bs, flags=flags)
                assert got == ref, (tr, ent, seed, flags, len(got), len(ref), first_diff(got, ref))
                assert K.decompress(ref, len(d) + 1024, flags=flags) == d, (tr, ent, seed, flags)
