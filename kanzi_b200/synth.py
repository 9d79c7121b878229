"""Deterministic synthetic corpora shaped like the reference's benchmark inputs (no corpora exist on the
box: SURVEY.md §8d).  numpy only; every generator is a pure function of (n, seed)."""
import numpy as np

_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_LETTER_P = np.array([12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2, 2.0, 2.0, 1.9, 1.5, 1.0, 0.8,
                      0.15, 0.15, 0.1, 0.07])
_LETTER_P = _LETTER_P / _LETTER_P.sum()


def _vocab(rng, nwords):
    lens = np.clip(rng.poisson(4.2, nwords) + 1, 1, 14)
    total = int(lens.sum())
    chars = _LETTERS[rng.choice(26, size=total, p=_LETTER_P)]
    offs = np.zeros(nwords + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    return chars, offs, lens


def _emit_words(rng, n, chars, offs, lens, zipf_s=1.1, sep=b" ", extra=None):
    """Zipf-distributed words joined by separators until n bytes."""
    nw = len(lens)
    p = 1.0 / np.arange(1, nw + 1) ** zipf_s
    cdf = np.cumsum(p / p.sum())
    out = np.empty(n + 64, dtype=np.uint8)
    pos = 0
    while pos < n:
        k = max(1024, (n - pos) // 5 + 16)
        idx = np.minimum(np.searchsorted(cdf, rng.random(k)), nw - 1)
        wl = lens[idx] + 1
        ends = np.cumsum(wl)
        keep = int(np.searchsorted(ends, n - pos, side="left")) + 1
        idx, wl, ends = idx[:keep], wl[:keep], ends[:keep]
        tot = int(ends[-1])
        starts = ends - wl
        buf = np.full(tot, sep[0], dtype=np.uint8)
        # gather characters of each word
        rep = np.repeat(np.arange(keep), lens[idx])
        within = np.arange(int(lens[idx].sum())) - np.repeat(np.cumsum(lens[idx]) - lens[idx], lens[idx])
        buf[starts[rep] + within] = chars[offs[idx][rep] + within]
        if extra is not None:
            extra(rng, buf, starts)
        take = min(tot, n - pos)
        out[pos:pos + take] = buf[:take]
        pos += take
    return out[:n]


def text(n, seed=1, nwords=50000, zipf_s=1.1):
    """English-like text: Zipf vocabulary, sentence punctuation and line breaks."""
    rng = np.random.default_rng(seed)
    chars, offs, lens = _vocab(rng, nwords)

    def punct(rng, buf, starts):
        m = rng.random(len(starts)) < 0.07
        s = starts[m]
        s = s[s > 0]
        buf[s - 1] = np.frombuffer(b".,;\n", dtype=np.uint8)[rng.integers(0, 4, len(s))]

    return _emit_words(rng, n, chars, offs, lens, zipf_s, extra=punct)


def pasted(n, seed=1):
    """Text in which long passages (70-1900 bytes) come back further on and the bytes 0xFC / 0xFE / 0xFF turn up here and there: what LZP
    (matches of 64+ bytes behind a 4-byte context, escape byte 0xFC) has something to do on."""
    rng = np.random.default_rng(seed + 977)
    a = text(n, seed).copy()
    for _ in range(n // 3000):
        src, ln, dst = int(rng.integers(0, max(n - 2000, 1))), int(rng.integers(70, 1900)), int(rng.integers(0, max(n - 2000, 1)))
        seg = a[src:src + ln].copy()
        a[dst:dst + len(seg)] = seg[:len(a[dst:dst + len(seg)])]
    for _ in range(n // 5000):
        a[int(rng.integers(0, n))] = int(rng.choice([0xFC, 0xFE, 0xFF]))
    return a


def ascii_markov(n, seed=1):
    """cfg1 input: printable ASCII with an English-like skew."""
    return text(n, seed, nwords=8000, zipf_s=1.0)


def markup(n, seed=1):
    """XML/HTML-like tagged text with repeated attribute names."""
    rng = np.random.default_rng(seed)
    body = text(n, seed + 101, nwords=20000)
    tags = [b"<row id=\"", b"\" name=\"", b"\"><value type=\"int\">", b"</value></row>\n", b"<entry key=\"", b"</entry>\n", b"<td class=\"c\">", b"</td>"]
    out = bytearray()
    pos = 0
    while len(out) < n:
        t = tags[int(rng.integers(0, len(tags)))]
        k = int(rng.integers(3, 40))
        out += t + body[pos:pos + k].tobytes()
        pos = (pos + k) % (n - 64)
    return np.frombuffer(bytes(out[:n]), dtype=np.uint8).copy()


def records(n, seed=1, width=48):
    """Database-like fixed-width records: counters, slowly varying fields, small enums."""
    rng = np.random.default_rng(seed)
    rows = n // width + 1
    rec = np.zeros((rows, width), dtype=np.uint8)
    ids = np.arange(rows, dtype=np.uint32)
    rec[:, 0:4] = ids.view(np.uint8).reshape(rows, 4)
    walk = np.cumsum(rng.integers(-3, 4, rows)).astype(np.int32)
    rec[:, 4:8] = walk.view(np.uint8).reshape(rows, 4)
    rec[:, 8] = rng.choice(5, rows, p=[0.6, 0.2, 0.1, 0.07, 0.03])
    rec[:, 9:17] = rng.integers(48, 58, (rows, 8))
    rec[:, 17:32] = _LETTERS[rng.choice(26, (rows, 15), p=_LETTER_P)]
    rec[:, 32:40] = 0
    rec[:, 40:48] = (rng.integers(0, 1 << 12, (rows, 1)) >> np.arange(8)) & 0xFF
    return rec.reshape(-1)[:n].copy()


def exe_like(n, seed=1, elf_header=True):
    """x86-like opcode stream: skewed opcode bytes, 32-bit displacements with locality, zero padding runs."""
    rng = np.random.default_rng(seed)
    ops = np.array([0x8B, 0x89, 0xE8, 0x48, 0x83, 0xFF, 0x0F, 0x85, 0x74, 0x75, 0xC3, 0x90, 0x55, 0x5D, 0x8D, 0x24, 0x00, 0x01, 0x44, 0x4C],
                   dtype=np.uint8)
    p = 1.0 / np.arange(1, len(ops) + 1)
    p /= p.sum()
    out = ops[rng.choice(len(ops), n, p=p)]
    # sprinkle displacement words with shared high bytes
    k = n // 12
    at = rng.integers(0, max(n - 8, 1), k)
    disp = (rng.integers(0, 1 << 14, k) + 0x00401000).astype(np.uint32)
    for i in range(4):
        out[np.minimum(at + i, n - 1)] = (disp >> (8 * i)) & 0xFF
    z = rng.integers(0, max(n - 64, 1), n // 4096 + 1)
    for s in z:
        out[s:s + int(rng.integers(8, 64))] = 0
    if elf_header and n >= 4:
        out[0:4] = np.frombuffer(b"\x7fELF", dtype=np.uint8)
    return out


def pcm_like(n, seed=1):
    """16-bit little-endian random walk (audio/image-like)."""
    rng = np.random.default_rng(seed)
    m = n // 2 + 1
    w = np.cumsum(rng.normal(0, 40, m)).astype(np.int64)
    w = ((w + 32768) % 65536).astype(np.uint16)
    return w.view(np.uint8)[:n].copy()


def noise(n, seed=1):
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8)


def skewed(n, seed=1, bits=3.0):
    """order-0 source with roughly `bits` bits/byte of entropy (geometric law over byte values)."""
    rng = np.random.default_rng(seed)
    if bits >= 7.99:
        return noise(n, seed)
    q = 1.0 - 2.0 ** (-bits / 1.45)
    v = rng.geometric(1.0 - q, n) - 1
    return (v % 256).astype(np.uint8)


def silesia_like(n=211957760, seed=2):
    """cfg2 input: segments mimicking the classes (and proportions) of silesia.tar's members."""
    parts = [(text, 0.05), (exe_like, 0.24), (pcm_like, 0.05), (markup, 0.16), (exe_like, 0.03), (records, 0.05), (text, 0.03),
             (markup, 0.10), (records, 0.03), (text, 0.20), (markup, 0.02), (pcm_like, 0.04)]
    out = np.empty(n, dtype=np.uint8)
    pos = 0
    for i, (fn, frac) in enumerate(parts):
        k = n - pos if i == len(parts) - 1 else min(n - pos, int(n * frac))
        if k <= 0:
            break
        out[pos:pos + k] = fn(k, seed * 1000 + i)
        pos += k
    return out


def enwik_like(n=100000000, seed=3):
    """cfg3/cfg4 input: Zipf(1.1) vocabulary text with wiki markup tokens and 1 % two-byte UTF-8."""
    rng = np.random.default_rng(seed)
    out = text(n, seed, nwords=50000, zipf_s=1.1)
    k = n // 100
    at = rng.integers(0, max(n - 2, 1), k)
    out[at] = 0xC3
    out[np.minimum(at + 1, n - 1)] = rng.integers(0x80, 0xC0, k)
    for tok in (b"[[", b"]]", b"'''", b"==", b"{{", b"}}", b"&lt;", b"&gt;"):
        t = np.frombuffer(tok, dtype=np.uint8)
        at = rng.integers(0, max(n - 8, 1), n // 400)
        for i in range(len(t)):
            out[at + i] = t[i]
    return out


def mixed_entropy(n, seed=5, block=16 << 20):
    """cfg5 input: blocks cycling through ~1, 3, 5, 7, 8 bits/byte with LZ-style repeats at distances < 64 Ki."""
    rng = np.random.default_rng(seed)
    out = np.empty(n, dtype=np.uint8)
    levels = [1.0, 3.0, 5.0, 7.0, 8.0]
    for b, pos in enumerate(range(0, n, block)):
        k = min(block, n - pos)
        seg = skewed(k, seed * 100 + b, levels[b % 5])
        nrep = k // 200
        dst = rng.integers(1024, max(k - 64, 1025), nrep)
        dist = rng.integers(1, 65535, nrep)
        ln = rng.integers(4, 48, nrep)
        for d, ds, l in zip(dst[:20000], dist[:20000], ln[:20000]):
            s = d - ds
            if s >= 0:
                seg[d:d + l] = seg[s:s + l]
        out[pos:pos + k] = seg
    return out


CONFIGS = {
    # name: (generator, full size, transforms, entropy, block size)   — BASELINE.json configs[0..4]
    "cfg1": (ascii_markov, 1 << 20, ["NONE"], "HUFFMAN", 64 << 10),
    "cfg2": (silesia_like, 211957760, ["LZ"], "ANS0", 4 << 20),
    "cfg3": (enwik_like, 100000000, ["BWT", "RANK", "ZRLT"], "ANS1", 8 << 20),
    "cfg4": (enwik_like, 1000000000, ["BWT", "SRT", "ZRLT"], "FPAQ", 32 << 20),
    "cfg5": (mixed_entropy, 8 << 30, ["ROLZ"], "ANS0", 16 << 20),
}


def make(cfg, scale=1.0):
    gen, size, tr, ent, bs = CONFIGS[cfg]
    n = max(1024, int(size * scale))
    seed = {"cfg1": 1, "cfg2": 2, "cfg3": 3, "cfg4": 4, "cfg5": 5}[cfg]
    return gen(n, seed), tr, ent, bs


if __name__ == "__main__":      # python -m kanzi_b200.synth <generator> <bytes> <seed> <out file>   (tools/make_goldens.sh)
    import sys
    _gen, _n, _seed, _out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    globals()[_gen](_n, _seed).tofile(_out)
