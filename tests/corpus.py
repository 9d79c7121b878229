"""Test inputs: the literal arrays / shapes of the reference's own tests (T/test/TestEntropyCodec.java:219-237,
TestTransforms.java:188-253, TestBWT.java:85-103) plus seeded synthetic data."""
import numpy as np
from kanzi_b200 import synth

# TestEntropyCodec.java:219-237 literal inputs
ENTROPY_LITERALS = [
    bytes([0x3d, 0x4d, 0x54, 0x47, 0x5a, 0x36, 0x39, 0x26, 0x72, 0x6f, 0x6c, 0x65, 0x3d, 0x70, 0x72, 0x65]),
    bytes([65, 71, 74, 66, 76, 65, 69, 77, 74, 79, 68, 75, 73, 72, 77, 68, 78, 65, 79, 79, 78, 66, 77, 71, 64, 70, 74, 77, 64, 67, 71, 64]),
    bytes(32),
    bytes([0, 2] * 16),
    bytes([0, 1, 2, 2, 2, 2, 7, 9, 9, 16, 16, 16, 1, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3]),
    bytes(range(32)),
]

BWT_LITERALS = [b"mississippi", b"3.14159265358979323846264338327950288419716939937510",
                b"SIX.MIXED.PIXIES.SIFT.SIXTY.PIXIE.DUST.BOXES", b"a", b"ab", b"aaaaaaaaaaaaaaaa"]


def rng(seed):
    return np.random.default_rng(seed)


def small_cases():
    """name -> bytes; sizes the oracle finishes instantly; covers empty-ish, ragged and degenerate inputs."""
    r = rng(1234)
    c = {}
    c["one"] = b"x"
    c["tiny3"] = b"xyz"
    c["len15"] = bytes(range(15))
    c["len16"] = bytes(range(16))
    c["len31"] = bytes(r.integers(0, 4, 31, dtype=np.uint8))
    c["len32"] = bytes(r.integers(0, 4, 32, dtype=np.uint8))
    c["len33"] = bytes(r.integers(0, 4, 33, dtype=np.uint8))
    c["len63"] = bytes(r.integers(0, 9, 63, dtype=np.uint8))
    c["len64"] = bytes(r.integers(0, 9, 64, dtype=np.uint8))
    c["zeros80k"] = bytes(80000)                      # TestTransforms "80000 identical bytes"
    c["ff80k"] = b"\xff" * 80000
    c["lots_of_zeros"] = bytes(np.where(r.random(70000) < 0.9, 0, r.integers(0, 256, 70000)).astype(np.uint8))
    c["fe_ff"] = bytes(r.choice([0xFE, 0xFF, 0, 1, 2], 50000).astype(np.uint8))
    c["runs"] = bytes(np.repeat(r.integers(0, 256, 3000, dtype=np.uint8), r.integers(1, 60, 3000)))
    c["text64k"] = synth.text(65536, 5).tobytes()
    c["text16385"] = synth.text(16385, 6).tobytes()
    c["text16383"] = synth.text(16383, 7).tobytes()
    c["text49155"] = synth.text(49155, 8).tobytes()
    c["rand20k"] = bytes(r.integers(0, 256, 20000, dtype=np.uint8))
    c["two_syms"] = bytes(r.integers(0, 2, 40000, dtype=np.uint8) * 7 + 3)
    c["one_sym_chunk"] = bytes(16384) + synth.text(20000, 9).tobytes() + b"\x05" * 16384
    c["skew"] = synth.skewed(100000, 10, 2.0).tobytes()
    c["exe"] = synth.exe_like(150000, 11).tobytes()
    c["records"] = synth.records(120000, 12).tobytes()
    c["dna"] = bytes(r.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 90000))
    c["numeric"] = bytes(r.choice(np.frombuffer(b"0123456789", dtype=np.uint8), 70000))
    c["rep17"] = (b"0123456789abcdef" * 7 + b"Z") * 900
    # sparse parses (LZ jumps over most positions): noise / PCM with and without long repeats
    c["pcm300k"] = synth.pcm_like(300000, 13).tobytes()
    c["noise_rep400k"] = sparse_with_repeats(400000, 14)
    c["pcm_rep300k"] = sparse_with_repeats(300000, 15, pcm=True)
    return c


def sparse_with_repeats(n, seed, pcm=False, every=6600, length=600):
    """Incompressible data with a long repeat every `every` bytes: LZ's miss acceleration jumps over most positions and
    still finds the repeats (the device parses such blocks in order against a real table)."""
    r = rng(seed)
    d = (synth.pcm_like(n, seed) if pcm else synth.noise(n, seed)).copy()
    for at in range(every, n - length - 1, every):
        at2 = at + int(r.integers(0, every // 2))
        src = int(r.integers(0, at2 - length))
        ln = int(r.integers(length // 4, length))
        if at2 + ln < n:
            d[at2:at2 + ln] = d[src:src + ln]
    return d.tobytes()


def fibonacci_chunk():
    """A 16 KiB chunk whose symbol counts follow a Fibonacci law: forces Huffman code lengths > 12 (limitCodeLengths)."""
    f = [1, 1]
    while sum(f) + f[-1] + f[-2] <= 16384:
        f.append(f[-1] + f[-2])
    sym = np.concatenate([np.full(c, i, dtype=np.uint8) for i, c in enumerate(f)])
    pad = 16384 - len(sym)
    sym = np.concatenate([sym, np.full(pad, len(f) - 1, dtype=np.uint8)])
    rng(99).shuffle(sym)
    return sym.tobytes()
