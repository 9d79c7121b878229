"""A second, independent restatement (plain Python, written from the Java sources alone) of the small transforms, held against the
C++ oracle on seeded inputs.  The oracle cannot be pinned to real Kanzi output in this image (no JVM: DESIGN.md "parity unpinned");
two restatements that were written separately and agree bit for bit are the next best evidence that the oracle reads the Java right.
Java semantics reproduced here: `int` wraps at 32 bits, bytes are unsigned after `& 0xFF`.
  ZRLT  K/transform/ZRLT.java:54-136 (forward), 146-233 (inverse)
  SBRT  K/transform/SBRT.java:87-151 (forward), 154-214 (inverse); modes MTF = 1, RANK = 2, TIMESTAMP = 3
  SRT   K/transform/SRT.java:73-168 (forward), 178-257 (inverse), preprocess :266-302, header :312-353
Further down, each with its own citation: the FPAQ, ANS0/ANS1, Huffman and RANGE encoders, the LZ/LZX and ROLZ forward transforms, the BWT
held against a naive suffix sort, the block framing of CompressedOutputStream, the sibling codecs LZP, RLT and ROLZX both ways, and the
decode side: LZ/LZX, ROLZ and SRT inverses, the rANS and FPAQ decoders, Huffman by a bit-by-bit prefix decoder, whole streams parsed the
way CompressedInputStream parses them.  The BWT inverses are
covered through the oracle's round trips (tests/test_oracle.py): an inverse that returns the input of a doubly-checked forward is
checked too."""
import numpy as np
import pytest
import oracle_lib as O


def _i32(x):
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def zrlt_forward(src):
    n = len(src)
    dst = bytearray()
    i = 0
    while i < n:
        if src[i] == 0:
            run = 1
            while i + run < n and src[i + run] == 0:
                run += 1
            i += run
            run += 1
            lg = run.bit_length() - 1
            if len(dst) >= n - lg:
                return False, bytes(dst)
            while lg > 0:
                lg -= 1
                dst.append((run >> lg) & 1)
            continue
        v = src[i]
        if v >= 0xFE:
            if len(dst) >= n - 1:
                return False, bytes(dst)
            dst += bytes([0xFF, v - 0xFE])
        else:
            if len(dst) >= n:
                return False, bytes(dst)
            dst.append(v + 1)
        i += 1
    return i == n, bytes(dst)


def zrlt_inverse(src, dst_end):
    n = len(src)
    dst = bytearray()
    i = 0
    run = 0
    while True:
        v = src[i]
        in_digits_at_end = False
        if v <= 1:
            run = 1
            while True:
                run = _i32(run + _i32(run + v))
                i += 1
                if i >= n:
                    in_digits_at_end = True
                    break
                v = src[i]
                if v > 1:
                    break
            if in_digits_at_end:
                break
            run = _i32(run - 1)
            if run > 0:
                if len(dst) + run >= dst_end:
                    break
                dst += bytes(run)
                run = 0
        if v == 0xFF:
            i += 1
            if i >= n:
                break
            dst.append((0xFE + src[i]) & 0xFF)
        else:
            dst.append((v - 1) & 0xFF)
        i += 1
        if i >= n or len(dst) >= dst_end:
            break
    if run > 0:
        run -= 1
        if len(dst) + run > dst_end:
            return False, bytes(dst)
        dst += bytes(run)
    return i == n, bytes(dst)


def sbrt(src, mode, inverse):
    m1 = 0 if mode == 3 else -1
    m2 = 0 if mode == 1 else -1
    s = 1 if mode == 2 else 0
    p, q = [0] * 256, [0] * 256
    r2s, s2r = list(range(256)), list(range(256))
    out = bytearray()
    for i, b in enumerate(src):
        if inverse:
            r = b
            c = r2s[r]
            out.append(c)
        else:
            c = b
            r = s2r[c]
            out.append(r)
        qc = ((i & m1) + (p[c] & m2)) >> s
        p[c] = i
        q[c] = qc
        while r > 0 and q[r2s[r - 1]] <= qc:
            r2s[r] = r2s[r - 1]
            s2r[r2s[r]] = r
            r -= 1
        r2s[r] = c
        s2r[c] = r
    return bytes(out)


def _srt_order(freqs):
    return sorted((c for c in range(256) if freqs[c] > 0), key=lambda c: (-freqs[c], c))


def srt_forward(src):
    n = len(src)
    freqs = [0] * 256
    r2s, s2r = [0] * 256, [0] * 256
    b = 0
    for v in src:
        if freqs[v] == 0:
            r2s[b] = v
            s2r[v] = b
            b += 1
        freqs[v] += 1
    buckets = [0] * 256
    pos = 0
    for c in _srt_order(freqs):
        buckets[c] = pos
        pos += freqs[c]
    hdr = bytearray()
    for f in freqs:
        while f >= 128:
            hdr.append(0x80 | (f & 0x7F))
            f >>= 7
        hdr.append(f)
    body = bytearray(n)
    i = 0
    while i < n:
        c = src[i]
        r = s2r[c]
        pp = buckets[c]
        body[pp] = r
        pp += 1
        if r != 0:
            while r != 0:
                r2s[r] = r2s[r - 1]
                s2r[r2s[r]] = r
                r -= 1
            r2s[0] = c
            s2r[c] = 0
        i += 1
        while i < n and src[i] == c:
            body[pp] = 0
            pp += 1
            i += 1
        buckets[c] = pp
    return bytes(hdr) + bytes(body)


def _cases():
    r = np.random.default_rng(77)
    out = []
    for k in range(24):
        n = int(r.integers(17, 6000))
        kind = k % 6
        if kind == 0:
            d = np.where(r.random(n) < 0.8, 0, r.integers(0, 256, n))
        elif kind == 1:
            d = r.choice([0, 0, 0, 0xFE, 0xFF, 3, 200], n)
        elif kind == 2:
            d = np.repeat(r.integers(0, 256, n // 7 + 1), r.integers(1, 15, n // 7 + 1))[:n]
        elif kind == 3:
            d = r.integers(0, 4, n) * 60
        elif kind == 4:
            d = r.integers(0, 256, n)
        else:
            d = r.choice(np.frombuffer(b"the quick brown fox jumps over the lazy dog\n", dtype=np.uint8), n)
        out.append(np.asarray(d, dtype=np.uint8).tobytes())
    return out


def test_zrlt_forward_and_inverse_agree_with_the_oracle():
    r = np.random.default_rng(5)
    for d in _cases():
        ok_ref, ref, _, _ = O.transform("ZRLT", d, dst_cap=len(d))
        ok, got = zrlt_forward(d)
        assert int(ok) == ok_ref
        if ok:
            assert got == ref
            ok2, back = zrlt_inverse(ref, len(d) + 64)
            assert ok2 and back == d
    # arbitrary bytes through the inverse (streams no encoder writes: long digit runs wrap a Java int, unpaired escapes)
    for k in range(40):
        n = int(r.integers(1, 400))
        d = bytes(r.choice([0, 1, 0, 1, 2, 7, 0xFF, 0xFE], n).astype(np.uint8))
        for cap in (n + 5000, 40):
            ok_ref, ref, _, _ = O.transform("ZRLT", d, inverse=True, dst_cap=cap)
            ok, got = zrlt_inverse(d, cap)
            assert int(ok) == ok_ref, (k, cap, d[:40])
            if ok:
                assert got == ref, (k, cap)


@pytest.mark.parametrize("name,mode", [("MTFT", 1), ("RANK", 2)])
def test_sbrt_agrees_with_the_oracle(name, mode):
    for d in _cases():
        ok_ref, ref, _, _ = O.transform(name, d, dst_cap=len(d))
        assert ok_ref == 1 and sbrt(d, mode, False) == ref
        assert sbrt(ref, mode, True) == d


def test_srt_forward_agrees_with_the_oracle():
    for d in _cases():
        ok_ref, ref, _, _ = O.transform("SRT", d, dst_cap=len(d) + 1024)
        assert ok_ref == 1 and srt_forward(d) == ref


# ---- FPAQ: K/entropy/FPAQEncoder.java:128-238 (64-bit Java longs: wrapping, `>>>` logical) -------------------------------------
M64 = (1 << 64) - 1


class _Bits:
    def __init__(self):
        self.v, self.n = 0, 0

    def write(self, value, nbits):
        self.v = (self.v << nbits) | (value & ((1 << nbits) - 1))
        self.n += nbits

    def varint(self, value):          # EntropyUtils.writeVarInt (K/entropy/EntropyUtils.java:259-276)
        while value >= 128:
            self.write(0x80 | (value & 0x7F), 8)
            value >>= 7
        self.write(value, 8)

    def bytes(self):
        pad = (-self.n) % 8
        return (self.v << pad).to_bytes((self.n + pad) // 8, "big"), self.n


def fpaq_encode(data, chunk=4 << 20):
    TOP, M24_56, M0_24, M0_32, PSCALE = 0x00FFFFFFFFFFFFFF, 0x00FFFFFFFF000000, 0xFFFFFF, 0xFFFFFFFF, 65536
    probs = [[PSCALE >> 1] * 256 for _ in range(4)]
    low, high = 0, TOP
    out = _Bits()
    start, end = 0, len(data)
    while start < end:
        size = min(chunk, end - start)
        sba = bytearray()
        p = probs[0]
        for i in range(start, start + size):
            val = data[i]
            bits = val + 256
            for k in range(8):
                bit = val & (0x80 >> k)
                idx = 1 if k == 0 else bits >> (8 - k)
                split = ((((high - low) & M64) >> 8) * p[idx] & M64) >> 8
                if bit == 0:
                    low = (low + split + 1) & M64
                    p[idx] -= p[idx] >> 6
                else:
                    high = (low + split) & M64
                    p[idx] -= (p[idx] - PSCALE + 64) >> 6
                while ((low ^ high) & M24_56) == 0:
                    sba += ((high >> 24) & 0xFFFFFFFF).to_bytes(4, "big")
                    low = (low << 32) & M64
                    high = ((high << 32) | M0_32) & M64
            p = probs[val >> 6]
        out.varint(len(sba))
        for b in sba:
            out.write(b, 8)
        start += size
        if start < end:
            out.write(low | M0_24, 56)
    out.write(low | M0_24, 56)          # dispose()
    return out.bytes()


def test_fpaq_encoder_agrees_with_the_oracle():
    for d in _cases()[:12] + [bytes(3000), bytes([255]) * 2000]:
        ref, nbits = O.entropy_encode("FPAQ", d)
        got, gbits = fpaq_encode(d)
        assert gbits == nbits and got == ref, (len(d), gbits, nbits)


# ---- rANS order 0 / order 1: K/entropy/ANSRangeEncoder.java:171-449, Symbol.reset :473-496; EntropyUtils.encodeAlphabet :38-75,
#      normalizeFrequencies :141-250 ---------------------------------------------------------------------------------------------------
def _normalize(freqs, total, scale):
    """in-place on freqs[0..255]; -> alphabet (list of symbols)"""
    if total == 0:
        return []
    if total == scale:
        return [i for i in range(256) if freqs[i] != 0]
    alphabet, sum_scaled, sum_freq, idx_max = [], 0, 0, 0
    for i in range(256):
        f = freqs[i]
        if f == 0:
            continue
        sf = f * scale
        scaled = 1 if sf <= total else (sf + (total >> 1)) // total
        alphabet.append(i)
        sum_scaled += scaled
        freqs[i] = scaled
        sum_freq += f
        if scaled > freqs[idx_max]:
            idx_max = i
        if sum_freq >= total:
            break
    if not alphabet:
        return []
    if len(alphabet) == 1:
        freqs[alphabet[0]] = scale
        return alphabet
    if sum_scaled == scale:
        return alphabet
    delta = sum_scaled - scale
    thr = freqs[idx_max] >> 4
    if abs(delta) <= thr:
        freqs[idx_max] -= delta
        return alphabet
    if delta < 0:
        delta += thr
        freqs[idx_max] += thr
    else:
        delta -= thr
        freqs[idx_max] -= thr
    inc = -1 if delta > 0 else 1
    delta = abs(delta)
    rnd = 0
    while True:
        rnd += 1
        if not (rnd < 6 and delta > 0):
            break
        adjustments = 0
        for idx in alphabet:
            if freqs[idx] <= 2:
                continue
            freqs[idx] += inc
            adjustments += 1
            delta -= 1
            if delta == 0:
                break
        if adjustments == 0:
            break
    freqs[idx_max] = max(freqs[idx_max] - delta, 1)
    return alphabet


def _encode_alphabet(out, alphabet):
    n = len(alphabet)
    if n == 0:
        out.write(0, 1); out.write(1, 1)
    elif n == 256:
        out.write(0, 1); out.write(0, 1)
    else:
        out.write(1, 1)
        masks = [0] * 32
        for a in alphabet:
            masks[a >> 3] |= 1 << (a & 7)
        last = alphabet[-1] >> 3
        out.write(last, 5)
        for i in range(last + 1):
            out.write(masks[i], 8)


def _symbol(cum, freq, lr):
    """-> (xMax, bias, cmplFreq, invShift, invFreq)"""
    if freq >= 1 << lr:
        freq = (1 << lr) - 1
    x_max = ((32768 >> lr) << 16) * freq
    cmpl = (1 << lr) - freq
    if freq < 2:
        return (x_max, cum + (1 << lr) - 1, cmpl, 32, 0xFFFFFFFF)
    shift = 0
    while freq > (1 << shift):
        shift += 1
    return (x_max, cum, cmpl, 32 + shift - 1, (((1 << (shift + 31)) + freq - 1) // freq) & 0xFFFFFFFF)


def ans_encode(data, order):
    out = _Bits()
    _ans_encode_into(out, data, order, 16384 if order == 0 else (4 << 20))
    return out.bytes()


def _ans_encode_into(out, data, order, chunk):
    """One ANSRangeEncoder.encode() call on an open bitstream (the ROLZ codec makes several, with its own chunk size)."""
    lr = 12 if order == 0 else 11
    n = len(data)
    if n <= 32:
        for b in data:
            out.write(b, 8)
        return
    nctx = 255 * order + 1
    zero = (0, 0, 0, 0, 0)
    symbols = [[zero] * 256 for _ in range(nctx)]            # `new Symbol()` per encode() call (:277-282)
    start = 0
    while start < n:
        end = min(start + chunk, n)
        freqs = [[0] * 257 for _ in range(nctx)]
        if order == 0:
            for b in data[start:end]:
                freqs[0][b] += 1
            freqs[0][256] = end - start
        else:
            quarter = (end - start) >> 2
            spans = [(start, end)] if quarter == 0 else [(start + q * quarter, start + (q + 1) * quarter) for q in range(4)]
            for a, z in spans:
                prv = 0
                for b in data[a:z]:
                    freqs[prv][b] += 1
                    freqs[prv][256] += 1
                    prv = b
        out.write(lr - 8, 3)
        total_alpha = 0
        for k in range(nctx):
            f = freqs[k]
            alphabet = _normalize(f, f[256], 1 << lr)
            cum = 0
            for s in alphabet:
                symbols[k][s] = _symbol(cum, f[s], lr)
                cum += f[s]
            _encode_alphabet(out, alphabet)
            if len(alphabet) > 1:
                chk = 8 if len(alphabet) >= 64 else 6
                llr = 3
                while (1 << llr) <= lr:
                    llr += 1
                for i in range(1, len(alphabet), chk):
                    endj = min(i + chk, len(alphabet))
                    mx = max(f[alphabet[j]] - 1 for j in range(i, endj))
                    log_max = 0
                    while (1 << log_max) <= mx:
                        log_max += 1
                    out.write(log_max, llr)
                    if log_max == 0:
                        continue
                    for j in range(i, endj):
                        out.write(f[alphabet[j]] - 1, log_max)
            total_alpha += len(alphabet)
        if total_alpha <= 1 and order == 0:
            start = end
            continue
        # encodeChunk (:337-407)
        buf = bytearray()                   # bytes in the order they are written, i.e. from the END of Java's buffer backwards
        end4 = start + ((end - start) & -4)
        for i in range(end - 1, end4 - 1, -1):
            buf.append(data[i])
        st = [32768] * 4

        def step(k, sym):
            x_max, bias, cmpl, inv_shift, inv = sym
            s = st[k]
            if s >= x_max:
                buf.append(s & 0xFF)
                buf.append((s >> 8) & 0xFF)
                s >>= 16
            st[k] = s + bias + ((s * inv) >> inv_shift) * cmpl

        if order == 0:
            i = end4 - 1
            while i > start:
                for k in range(4):
                    step(k, symbols[0][data[i - k]])
                i -= 4
        else:
            quarter = (end4 - start) >> 2
            idx = [start + (q + 1) * quarter - 2 for q in range(4)]
            prv = [data[j + 1] if j + 1 >= 0 else 0 for j in idx]
            while idx[0] >= start:
                for k in range(4):
                    cur = data[idx[k]]
                    step(k, symbols[cur][prv[k]])
                    prv[k] = cur
                    idx[k] -= 1
            for k in range(4):
                step(k, symbols[0][prv[k]])
        out.varint(len(buf))
        for k in range(4):
            out.write(st[k], 32)
        for b in reversed(buf):
            out.write(b, 8)
        start = end


@pytest.mark.parametrize("kind,order", [("ANS0", 0), ("ANS1", 1)])
def test_ans_encoder_agrees_with_the_oracle(kind, order):
    r = np.random.default_rng(9)
    # (order 1 writes 256 context headers: inputs are sized so that the oracle wrapper's 2 n + 4 KiB output buffer holds them)
    noise = bytes(r.integers(0, 256, 40000 if order == 0 else 200000, dtype=np.uint8))
    cases = [c for c in _cases()[:10] if order == 0 or len(set(c)) <= 64] + [noise, bytes(20000), bytes([7]) * 33,
             bytes(r.choice([65, 66, 67, 200], 50001).astype(np.uint8)), bytes(range(32)), bytes(r.choice([1, 2, 3], 35).astype(np.uint8))]
    for d in cases:
        ref, nbits = O.entropy_encode(kind, d)
        got, gbits = ans_encode(d, order)
        assert gbits == nbits and got == ref, (kind, len(d), gbits, nbits)


# ---- Huffman: K/entropy/HuffmanEncoder.java:103-493, HuffmanCommon.generateCanonicalCodes :71-111, ExpGolombEncoder (signed table) ----
def _expgolomb_signed(val):
    """-> (bits, nbits) as the reference's signed cache emits them (ExpGolombEncoder.java:52-70: emit & 0x1FF in emit >>> 9 bits)"""
    if val == 0:
        return 1, 1
    a = abs(val)
    lg = (a + 1).bit_length() - 1
    return (1 << (lg + 1)) | (((a + 1 - (1 << lg)) << 1) | (1 if val < 0 else 0)), 2 * lg + 2


def test_expgolomb_formula_matches_the_reference_table():
    # entries of CACHE_VALUES[1] (signed), index = val & 0xFF
    table = {1: 2052, 2: 2054, 3: 3080, 4: 3082, 7: 4112, 15: 5152, 255: 2053, 254: 2055, 253: 3081, 249: 4113}
    for idx, emit in table.items():
        val = idx if idx < 128 else idx - 256
        bits, n = _expgolomb_signed(val)
        assert (n << 9) | bits == emit, (idx, bits, n, emit)


def _hf_phase1(data, n):
    s = r = 0
    for t in range(n - 1):
        total = 0
        for _ in range(2):
            if s >= n or (r < t and data[r] < data[s]):
                total += data[r]
                data[r] = t
                r += 1
                continue
            total += data[s]
            if s > t:
                data[s] = 0
            s += 1
        data[t] = total


def _hf_phase2(data, n):
    if n < 2:
        return 0
    level_top, depth, i, total_nodes = n - 2, 1, n, 2
    while i > 0:
        k = level_top
        while k > 0 and data[k - 1] >= level_top:
            k -= 1
        internal = level_top - k
        for _ in range(total_nodes - internal):
            i -= 1
            data[i] = depth
        total_nodes = internal << 1
        level_top = k
        depth += 1
    return depth - 1


def _hf_code_lengths(sizes, ranks, count):
    ranks[:count] = sorted(ranks[:count])
    freqs = [0] * 256
    for i in range(count):
        freqs[i] = ranks[i] >> 8
        ranks[i] &= 0xFF
        if freqs[i] == 0:
            return 0
    _hf_phase1(freqs, count)
    mx = _hf_phase2(freqs, count)
    for i in range(count):
        sizes[ranks[i]] = freqs[i]
    return mx


def _hf_limit(alphabet, freqs, sizes, ranks, count):
    MAXL = 12
    n = debt = 0
    while sizes[ranks[n]] >= MAXL:
        debt += sizes[ranks[n]] - MAXL
        sizes[ranks[n]] = MAXL
        n += 1
    ll = [[] for _ in range(6)]
    while n < count:
        idx = MAXL - 1 - sizes[ranks[n]]
        if idx >= 6 or debt < (1 << idx):
            break
        ll[idx].append(ranks[n])
        n += 1
    idx = 5
    while debt > 0 and idx >= 0:
        if not ll[idx] or debt < (1 << idx):
            idx -= 1
            continue
        sizes[ll[idx].pop(0)] += 1
        debt -= 1 << idx
    idx = 0
    while debt > 0 and idx < 6:
        if not ll[idx]:
            idx += 1
            continue
        sizes[ll[idx].pop(0)] += 1
        debt -= 1 << idx
    if debt > 0:
        f = [freqs[alphabet[i]] for i in range(count)] + [0] * (256 - count)
        total = sum(f)
        _normalize(f, total, 16384 >> 3)        # (over the first `count` entries: the reference passes a count-long array)
        for i in range(count):
            freqs[alphabet[i]] = f[i]
            ranks[i] = (f[i] << 8) | alphabet[i]
        return _hf_code_lengths(sizes, ranks, count)
    return MAXL


def huffman_encode(data, chunk=16384):
    out = _Bits()
    n = len(data)
    start = 0
    while start < n:
        size = min(chunk, n - start)
        blk = data[start:start + size]
        if size < 32:
            for b in blk:
                out.write(b, 8)
            start += size
            continue
        freqs = [0] * 256
        for b in blk:
            freqs[b] += 1
        alphabet = [i for i in range(256) if freqs[i] > 0]
        count = len(alphabet)
        codes, sizes = [0] * 256, [0] * 256
        _encode_alphabet(out, alphabet)
        if count == 1:
            sizes[alphabet[0]] = 1
        else:
            ranks = [(freqs[a] << 8) | a for a in alphabet] + [0] * (256 - count)
            mx = _hf_code_lengths(sizes, ranks, count)
            assert mx != 0
            if mx > 12:
                mx = _hf_limit(alphabet, freqs, sizes, ranks, count)
                assert mx != 0
            if mx > 12:
                for i, a in enumerate(alphabet):
                    codes[a] = i
                    sizes[a] = 8
            else:
                order = sorted(ranks[:count], key=lambda s: (sizes[s], s))
                code, cur = 0, sizes[order[0]]
                for s in order:
                    code <<= sizes[s] - cur
                    cur = sizes[s]
                    codes[s] = code
                    code += 1
        prev = 2
        for a in alphabet:
            bits, nb = _expgolomb_signed(sizes[a] - prev)
            out.write(bits, nb)
            prev = sizes[a]
        if count > 1:
            frag = size // 4
            parts = []
            for j in range(4):
                v, nb = 0, 0
                for b in blk[j * frag:(j + 1) * frag]:
                    v = (v << sizes[b]) | codes[b]
                    nb += sizes[b]
                parts.append((v, nb))
            for _, nb in parts:
                out.varint(nb)
            for v, nb in parts:
                out.write(v, nb)
            for b in blk[4 * frag:]:
                out.write(b, 8)
        start += size
    return out.bytes()


def test_huffman_encoder_agrees_with_the_oracle():
    import corpus
    r = np.random.default_rng(11)
    cases = _cases()[:12] + [bytes(r.integers(0, 256, 40000, dtype=np.uint8)), bytes(20000), bytes([7]) * 33, bytes(range(32)), bytes(range(40)),
                             corpus.fibonacci_chunk(), corpus.fibonacci_chunk() + bytes(r.integers(0, 7, 5000, dtype=np.uint8))]
    for d in cases:
        ref, nbits = O.entropy_encode("HUFFMAN", d)
        got, gbits = huffman_encode(d)
        assert gbits == nbits and got == ref, (len(d), gbits, nbits)


# ---- LZ / LZX forward: K/transform/LZCodec.java:299-597 (LZXCodec.forward), hash :904-911, emitLength :209-231, findMatch :271-287 ----
def lzx_forward(src, extra=False, data_type_dna=False):
    """-> (ok, out bytes) for a slice with index 0, dst capacity getMaxEncodedLength; None if the Java code would throw"""
    count = len(src)
    if count == 0:
        return True, b""
    if count < 24:
        return False, b""
    hlog = 19 if extra else 16
    hashes = [0] * (1 << hlog)
    min_buf = max(count // 5, 256)
    mbuf, mlen, tk = bytearray(min_buf), bytearray(min_buf), bytearray(min_buf)
    MAXD1, MAXD2, MAX_MATCH = (1 << 16) - 2, (1 << 24) - 2, 65535 + 254 + 4
    src_end = count - 16 - 2
    max_dist = MAXD1 if src_end < 4 * MAXD1 else MAXD2
    mm = 6 if data_type_dna else 4
    dst = bytearray(count + (16 if count <= 1024 else count // 64) + 2 + 64)
    dst[12] = (0 if max_dist == MAXD1 else 1) | (((mm - 2) & 7) << 1)

    def h(i):
        v = int.from_bytes(src[i:i + 8], "little")
        return (((v << 24) * 0x1E35A7BD) & M64) >> (64 - hlog)

    def diff4(a, b):
        return src[a:a + 4] != src[b:b + 4]

    def find(a, ref, max_match):
        best = 0
        while best + 8 <= max_match:
            x = int.from_bytes(src[a + best:a + best + 8], "little") ^ int.from_bytes(src[ref + best:ref + best + 8], "little")
            if x != 0:
                best += ((x & -x).bit_length() - 1) >> 3
                break
            best += 8
        return best

    def emit_len(buf, idx, length):
        if length < 254:
            buf[idx] = length
            return idx + 1
        if length < 65536 + 254:
            length -= 254
            buf[idx] = 254; buf[idx + 1] = (length >> 8) & 0xFF; buf[idx + 2] = length & 0xFF
            return idx + 3
        length -= 255
        buf[idx] = 255; buf[idx + 1] = (length >> 16) & 0xFF; buf[idx + 2] = (length >> 8) & 0xFF; buf[idx + 3] = length & 0xFF
        return idx + 4

    si = anchor = 0
    di = 13
    m_idx = ml_idx = tk_idx = 0
    repd = [count, count]
    rep_idx = 0
    src_inc = 0
    try:
        while si < src_end:
            best = 0
            h0 = h(si)
            ref0 = hashes[h0]
            hashes[h0] = si
            si1 = si + 1
            ref = si1 - repd[rep_idx]
            min_ref = max(si - max_dist, 0)
            if ref > min_ref and not diff4(ref, si1):
                best = find(si1, ref, min(src_end - si1, MAX_MATCH))
            else:
                ref = si1 - repd[rep_idx ^ 1]
                if ref > min_ref and not diff4(ref, si1):
                    best = find(si1, ref, min(src_end - si1, MAX_MATCH))
            if best < mm:
                ref = ref0
                if ref > min_ref and not diff4(ref, si):
                    best = find(si, ref, min(src_end - si, MAX_MATCH))
                if best < mm:
                    si = si1 + (src_inc >> 6)
                    src_inc += 1
                    rep_idx = 0
                    continue
                if ref != si - repd[0] and ref != si - repd[1]:
                    h1 = h(si1)
                    ref1 = hashes[h1]
                    hashes[h1] = si1
                    if ref1 > min_ref + 1 and not diff4(ref1 + best - 3, si1 + best - 3):
                        b1 = find(si1, ref1, min(src_end - si1, MAX_MATCH))
                        if b1 >= best:
                            ref, best, si = ref1, b1, si1
                    if extra:
                        si2 = si1 + 1
                        h2 = h(si2)
                        ref2 = hashes[h2]
                        hashes[h2] = si2
                        if ref2 > min_ref + 2 and not diff4(ref2 + best - 3, si2 + best - 3):
                            b2 = find(si2, ref2, min(src_end - si2, MAX_MATCH))
                            if b2 >= best:
                                ref, best, si = ref2, b2, si2
                while si > anchor and ref > min_ref and src[si - 1] == src[ref - 1]:
                    best += 1
                    ref -= 1
                    si -= 1
                if best > MAX_MATCH:
                    ref += best - MAX_MATCH
                    si += best - MAX_MATCH
                    best = MAX_MATCH
            else:
                if best >= MAX_MATCH or src[si] != src[ref - 1]:
                    si += 1
                    hashes[h(si)] = si
                else:
                    best += 1
                    ref -= 1
            src_inc = 0
            dist = si - ref
            if dist == repd[0]:
                token, th = 0x00, 3
            elif dist == repd[1]:
                token, th = 0x04, 3
            else:
                mbuf[m_idx] = (dist >> 16) & 0xFF
                inc1 = 1 if dist >= 65536 else 0
                m_idx += inc1
                mbuf[m_idx] = (dist >> 8) & 0xFF
                inc2 = 1 if dist >= 256 else 0
                m_idx += inc2
                mbuf[m_idx] = dist & 0xFF
                m_idx += 1
                token, th = (inc1 + inc2 + 1) << 3, 7
            m_len = best - mm
            if m_len >= th:
                token += th
                ml_idx = emit_len(mlen, ml_idx, m_len - th)
            else:
                token += m_len
            repd[1] = repd[0]
            repd[0] = dist
            rep_idx = 1
            lit = si - anchor
            if lit == 0:
                tk[tk_idx] = token
                tk_idx += 1
            else:
                if lit >= 7:
                    if lit >= (1 << 24):
                        return False, b""
                    tk[tk_idx] = (7 << 5) | token
                    tk_idx += 1
                    di = emit_len(dst, di, lit - 7)
                else:
                    tk[tk_idx] = (lit << 5) | token
                    tk_idx += 1
                dst[di:di + lit] = src[anchor:anchor + lit]
                di += lit
            if m_idx >= len(mbuf) - 8:
                mbuf += bytearray((len(mbuf) * 3) // 2 - len(mbuf))
                if ml_idx >= len(mlen) - 4:
                    mlen += bytearray((len(mlen) * 3) // 2 - len(mlen))
            anchor = si + best
            while si + 4 < anchor:
                si += 4
                for k in (3, 2, 1, 0):
                    hashes[h(si - k)] = si - k
            si += 1
            while si < anchor:
                hashes[h(si)] = si
                si += 1
        lit = count - anchor
        if di + lit + tk_idx + m_idx + ml_idx >= count:
            return False, b""
        if lit >= 7:
            tk[tk_idx] = 7 << 5
            tk_idx += 1
            di = emit_len(dst, di, lit - 7)
        else:
            tk[tk_idx] = lit << 5
            tk_idx += 1
        dst[di:di + lit] = src[anchor:anchor + lit]
        di += lit
    except IndexError:
        return None, b""            # Java: ArrayIndexOutOfBoundsException (tkBuf is never grown)
    dst[0:4] = di.to_bytes(4, "little"); dst[4:8] = tk_idx.to_bytes(4, "little"); dst[8:12] = m_idx.to_bytes(4, "little")
    out = bytes(dst[:di]) + bytes(tk[:tk_idx]) + bytes(mbuf[:m_idx]) + bytes(mlen[:ml_idx])
    return len(out) <= count - count // 100, out


@pytest.mark.parametrize("name,extra", [("LZ", False), ("LZX", True)])
def test_lz_forward_agrees_with_the_oracle(name, extra):
    import corpus
    from kanzi_b200 import synth
    r = np.random.default_rng(21)
    cases = [synth.text(30000, 5).tobytes(), synth.exe_like(40000, 6).tobytes(), synth.records(25000, 7).tobytes(), (b"0123456789abcdef" * 7 + b"Z") * 300,
             bytes(np.repeat(r.integers(0, 256, 900, dtype=np.uint8), r.integers(1, 60, 900))), corpus.sparse_with_repeats(60000, 14),
             bytes(70000), synth.text(70000, 9).tobytes() + synth.text(70000, 9).tobytes(), bytes(r.integers(0, 256, 5000, dtype=np.uint8)), b"ab" * 20]
    applied = 0
    for d in cases:
        cap = len(d) + len(d) // 64 + 1100
        ok_ref, ref, _, _ = O.transform(name, d, dst_cap=cap, ctx=[7, max(len(d), 1024), len(d), 1, 0, 0])
        ok, got = lzx_forward(d, extra)
        assert ok is not None
        assert int(ok) == ok_ref, (name, len(d), ok, ok_ref)
        if ok:
            applied += 1
            assert got == ref, (name, len(d), len(got), len(ref))
    assert applied >= 6


# ---- BWT forward by definition: K/transform/BWT.java:186-187, DivSufSort.java:204-227 (output layout), :230-322 (primary indexes) ----
def bwt_by_definition(d):
    """Suffixes sorted naively (a suffix that is a prefix of another sorts first).  out[0] = last byte; rank i before suffix 0's rank
    p -> out[i+1], after -> out[i].  indexes[k] = rank(suffix k*step)+1 for the chunk starts, indexes[0] = p+1."""
    n = len(d)
    sa = sorted(range(n), key=lambda i: d[i:])
    rank = {s: i for i, s in enumerate(sa)}
    p = rank[0]
    out = bytearray(n)
    out[0] = d[n - 1]
    for i, s in enumerate(sa):
        if i < p:
            out[i + 1] = d[s - 1]
        elif i > p:
            out[i] = d[s - 1]
    chunks = 1 if n < 256 else 8
    st = n // chunks
    step = st + 1 if st * chunks != n else st
    idx = [0] * 8
    for s in range(0, n, step):
        idx[s // step] = rank[s] + 1
    return bytes(out), idx


def test_bwt_forward_matches_a_naive_suffix_sort():
    import corpus
    from kanzi_b200 import synth
    r = np.random.default_rng(33)
    cases = list(corpus.BWT_LITERALS) + [synth.text(3001, 4).tobytes(), bytes(r.integers(0, 4, 2048, dtype=np.uint8)), bytes(r.integers(0, 256, 255, dtype=np.uint8)),
                                         bytes(r.integers(0, 256, 256, dtype=np.uint8)), b"abcab" * 400, bytes(1500), b"ba" * 700 + b"c", synth.records(4000, 8).tobytes()]
    for d in cases:
        if len(d) < 2:
            continue
        ok, out, pi = O.bwt_forward(d)
        want, idx = bwt_by_definition(d)
        assert ok == 1 and out == want, len(d)
        assert pi[:len(idx)] == idx, (len(d), pi, idx)


# ---- ROLZ forward (ROLZCodec1): K/transform/ROLZCodec.java:419-677, findMatch :365-416, emitLength :679-693, keys/hash :123-149 ----
def _is_dna(d):   # Global.detectSimpleType's first rule (Global.java:556-566); the other types it can return leave ROLZ's parameters alone
    return len(d) > 0 and sum(d.count(bytes([c])) for c in b"acgntuACGNTU") > len(d) - len(d) // 12


def rolz_forward(src, log_checks=4):
    """-> (ok, out) for ROLZCodec1 with a non-null context of undefined data type; ok None where the Java code would throw"""
    M32, HASH, HMASK = 0xFFFFFFFF, 200002979, 0xFF000000
    count = len(src)
    if count == 0:
        return True, b""
    if count < 64:
        return False, b""
    src_end = count - 4
    size_chunk = min(count, 16 << 20)
    lit = bytearray(size_chunk + 64 if size_chunk <= 512 else size_chunk)
    lenb, midx, tk = bytearray(size_chunk // 5), bytearray(size_chunk // 4), bytearray(size_chunk // 4)
    counters = [0] * 65536
    lit_order = 0 if count < (1 << 17) else 1
    flags, mm, dt = lit_order, 3, 2
    if _is_dna(src):
        dt, mm, flags = 8, 7, flags | 4
    flags |= log_checks << 4
    out = bytearray(count.to_bytes(4, "big")) + bytes([flags])
    checks = 1 << log_checks
    mask = checks - 1

    def key_at(i):
        if mm == 3:
            return src[i] | (src[i + 1] << 8)
        return (((int.from_bytes(src[i:i + 8], "little") * HASH) & M64) >> 40) & 0xFFFF

    def hash_at(i):
        return ((((int.from_bytes(src[i:i + 4], "little") << 8) & M32) * HASH) & M32) & HMASK

    def emit_len(n):
        nonlocal nl
        if n >= 1 << 7:
            if n >= 1 << 14:
                if n >= 1 << 21:
                    lenb[nl] = 0x80 | (n >> 21) & 0xFF; nl += 1
                lenb[nl] = (0x80 | (n >> 14)) & 0xFF; nl += 1
            lenb[nl] = (0x80 | (n >> 7)) & 0xFF; nl += 1
        lenb[nl] = n & 0x7F; nl += 1

    start = 0
    try:
        while start < src_end:
            nlit = nl = nm = nt = 0
            matches = [0] * (65536 << log_checks)
            end_chunk = min(start + size_chunk, src_end)
            size_chunk = end_chunk - start
            si = start

            def find(pos, h32, counter, base):
                best_len, best_idx = 0, -1
                max_match = min(3 + 65535, end_chunk - pos) - 8
                for i in range(counter, counter - checks, -1):
                    ref = matches[base + (i & mask)]
                    if (ref & HMASK) != h32:
                        continue
                    ref = (ref & 0xFFFFFF) + start
                    if src[ref + best_len] != src[pos + best_len]:
                        continue
                    n = 0
                    while n < max_match:
                        x = int.from_bytes(src[ref + n:ref + n + 8], "little") ^ int.from_bytes(src[pos + n:pos + n + 8], "little")
                        if x:
                            n += ((x & -x).bit_length() - 1) >> 3
                            break
                        n += 8
                    if n > best_len:
                        best_idx, best_len = counter - i, n
                return -1 if best_len < mm else (best_idx << 16) | (best_len - mm)

            for _ in range(min(src_end - start, 8)):
                lit[nlit] = src[si]; nlit += 1; si += 1
            first_lit = si
            src_inc = 0
            while si < end_chunk:
                key = key_at(si - dt)
                base = key << log_checks
                h32 = hash_at(si)
                match = find(si, h32, counters[key], base)
                counters[key] = (counters[key] + 1) & mask
                matches[base + counters[key]] = h32 | (si - start)
                if match == -1:
                    si += 1 + (src_inc >> 6)
                    src_inc += 1
                    continue
                key = key_at(si + 1 - dt)
                base = key << log_checks
                h32 = hash_at(si + 1)
                match2 = find(si + 1, h32, counters[key], base)
                if match2 >= 0 and (match2 & 0xFFFF) > (match & 0xFFFF):
                    match = match2
                    si += 1
                    counters[key] = (counters[key] + 1) & mask
                    matches[base + counters[key]] = h32 | (si - start)
                lit_len = si - first_lit
                token = (lit_len << 3) if lit_len < 31 else 0xF8
                m_len = match & 0xFFFF
                if m_len >= 7:
                    tk[nt] = token | 7; nt += 1
                    emit_len(m_len - 7)
                else:
                    tk[nt] = token | m_len; nt += 1
                if lit_len >= 31:
                    emit_len(lit_len - 31)
                if nlit + lit_len > len(lit):
                    raise IndexError
                lit[nlit:nlit + lit_len] = src[first_lit:first_lit + lit_len]
                nlit += lit_len
                midx[nm] = (match >> 16) & 0xFF; nm += 1
                si += m_len + mm
                first_lit = si
                src_inc = 0
            lit_len = size_chunk - (first_lit - start)
            if nt != 0:
                tk[nt] = 0xF8 if lit_len >= 31 else (lit_len << 3); nt += 1
            if lit_len >= 31:
                emit_len(lit_len - 31)
            if nlit + lit_len > len(lit):
                raise IndexError
            lit[nlit:nlit + lit_len] = src[first_lit:first_lit + lit_len]
            nlit += lit_len
            bits = _Bits()
            for v in (nlit, nt, nl, nm):
                bits.write(v, 32)
            _ans_encode_into(bits, bytes(lit[:nlit]), lit_order, 16384 if lit_order == 0 else (4 << 20))
            for part in (tk[:nt], lenb[:nl], midx[:nm]):
                _ans_encode_into(bits, bytes(part), 0, 32768)
            out += bits.bytes()[0]
            start = end_chunk
    except IndexError:
        return None, b""
    out += src[src_end:src_end + 4]
    return True, bytes(out)


def test_rolz_forward_agrees_with_the_oracle():
    import corpus
    from kanzi_b200 import synth
    r = np.random.default_rng(27)
    dna = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[r.integers(0, 4, 6000)])
    dna = dna + dna[1000:3000] + dna[:2500]
    cases = [synth.text(30000, 5).tobytes(), synth.exe_like(40000, 6).tobytes(), synth.records(25000, 7).tobytes(), (b"0123456789abcdef" * 7 + b"Z") * 300,
             corpus.sparse_with_repeats(60000, 14), synth.text(70000, 9).tobytes() + synth.text(70000, 9).tobytes(), dna,
             bytes(r.integers(0, 256, 5000, dtype=np.uint8)), b"ab" * 20, b"xyz" * 40]
    applied = 0
    for d in cases:
        cap = max(len(d) + 64, 1024) + 1024
        ok_ref, ref, _, _ = O.transform("ROLZ", d, dst_cap=cap, ctx=[7, max(len(d), 1024), len(d), 1, 0, 0])
        ok, got = rolz_forward(d)
        assert ok is not None
        if ok and len(got) > cap:
            ok = False
        assert int(ok) == ok_ref, (len(d), ok, ok_ref)
        if ok:
            applied += 1
            assert got == ref, (len(d), len(got), len(ref))
    assert applied >= 7


# ---- block framing: K/io/CompressedOutputStream.java:677-1035 (encodeBlock), close :483-493 -------------------------------------
def _mix32(c, h, v):
    c ^= (h * (~v & 0xFFFFFFFF)) & 0xFFFFFFFF
    c = ((c << 13) | (c >> 19)) & 0xFFFFFFFF
    return (c * 5 + 0x52DCE729) & 0xFFFFFFFF


def frame_blocks(data, transforms, entropy, bs, checksum, header_len):
    """The stream after its header, built from the oracle's per-block transform sequence and entropy coder plus a restatement of the
    block framing: mode byte, skip flags, post-transform length, header check byte, block checksum, raw fallback, length prefix."""
    nfun = len(transforms)
    out = _Bits()
    for b0 in range(0, len(data), bs):
        blk = data[b0:b0 + bs]
        n = len(blk)
        ck = None
        if checksum == 32:
            ck = (O.xxhash32(blk, 0x4B414E5A) & 0xFFFFFFFF, 32)
        elif checksum == 64:
            ck = (O.xxhash64(blk, 0x4B414E5A) & M64, 64)
        mode = 0
        if n <= 15:
            post, skip, nf, ent = blk, 0x7F, 1, "NONE"      # a NONE sequence: its one function "applies" (Sequence.java:107), the unused slots stay set
            mode |= 0x80
        else:
            post, skip = O.sequence_forward(transforms, blk, bs)
            nf, ent = nfun, entropy
        pl = len(post)
        size = 1 if pl < 256 else ((pl.bit_length() - 1) >> 3) + 1
        mode |= ((size - 1) & 3) << 5

        def header(w, m, raw):
            hs = skip
            if raw:
                w.write(m, 8)
                if nf > 4:
                    w.write(skip, 8)
                else:
                    hs = ((m << 4) | 0x0F) & 0xFF
            elif (m & 0x80) or nf <= 4:
                m |= skip >> 4
                hs = 0 if (m & 0x80) else ((m << 4) | 0x0F) & 0xFF
                w.write(m, 8)
            else:
                m |= 0x10
                w.write(m, 8)
                w.write(skip, 8)
            w.write(pl, 8 * size)
            at = w.n
            w.write(0, 8)
            if ck:
                w.write(*ck)
            return m, hs, at

        w = _Bits()
        mode2, hs, at = header(w, mode, False)
        payload, nbits = O.entropy_encode(ent, post)
        w.write(int.from_bytes(payload[:nbits // 8], "big"), 8 * (nbits // 8))
        if nbits % 8:
            w.write(payload[nbits // 8] >> (8 - nbits % 8), nbits % 8)
        if not (mode2 & 0x80) and pl < ((w.n + 7) >> 3):
            w = _Bits()
            mode2, hs, at = header(w, mode2 | 0x80 | 0x10, True)
            w.write(int.from_bytes(post, "big"), 8 * len(post))
        written = w.n
        c = (0x1E35A7BD * 0x01030507) & 0xFFFFFFFF
        for v in (mode2 & 0xFF, hs & 0xFF, pl, written >> 32, written & 0xFFFFFFFF):
            c = _mix32(c, 0x1E35A7BD, v)
        body = w.v | ((((c >> 23) ^ (c >> 3)) & 0xFF) << (w.n - at - 8))
        lw = 3 if written < 8 else ((written >> 3).bit_length() - 1) + 4
        out.write(lw - 3, 5)
        out.write(written, lw)
        out.v = (out.v << written) | body
        out.n += written
    out.write(0, 5)
    out.write(0, 3)
    return out.bytes()[0]


@pytest.mark.parametrize("transforms,entropy,bs,checksum", [(["LZ"], "HUFFMAN", 65536, 0), (["BWT", "RANK", "ZRLT"], "ANS0", 32768, 32), (["ROLZ"], "NONE", 65536, 64),
                                                            (["NONE"], "FPAQ", 4096, 0), (["LZX"], "ANS1", 1 << 20, 32), (["LZ", "RANK", "ZRLT", "SRT", "MTFT"], "HUFFMAN", 16384, 0)])
def test_block_framing_agrees_with_the_oracle(transforms, entropy, bs, checksum):
    from kanzi_b200 import synth
    r = np.random.default_rng(5)
    data = synth.text(100000, 3).tobytes() + bytes(r.integers(0, 256, 70000, dtype=np.uint8)) + synth.records(40000 + 9, 4).tobytes()
    for d in (data, data[:bs + 7], data[:12], data[100000:100000 + 2 * bs]):
        ref = O.compress(d, transforms, entropy, bs, checksum=checksum)
        hl = len(O.stream_header(transforms, entropy, bs, len(d)))
        assert ref[hl:] == frame_blocks(d, transforms, entropy, bs, checksum, hl), (len(d), transforms)


# ---- LZP: K/transform/LZCodec.java:1003-1118 (forward), :1121-1263 (inverse), findMatch :1267-1280 ---------------------------
def lzp_forward(src):
    """-> (ok, out)"""
    M32 = 0xFFFFFFFF
    count = len(src)
    if count == 0:
        return True, b""
    if count < 128:
        return False, b""
    hashes = [0] * 65536
    dst = bytearray(count + (16 if count <= 1024 else count // 64))
    src_end, dst_end = count, count - (count >> 6)
    dst[0:4] = src[0:4]
    ctx = int.from_bytes(src[0:4], "little")
    si = di = 4
    while si < src_end - 64 and di < dst_end:
        h = ((0x7FEB352D * ctx) & M32) >> 16
        ref = hashes[h]
        hashes[h] = si
        best = 0
        if ref != 0 and src[ref + 60:ref + 64] == src[si + 60:si + 64]:
            mx = src_end - si
            while best + 8 <= mx:
                a, b = src[si + best:si + best + 8], src[ref + best:ref + best + 8]
                if a != b:
                    best += next(k for k in range(8) if a[k] != b[k])
                    break
                best += 8
        if best < 64:
            val = src[si]
            ctx = ((ctx << 8) | val) & M32
            dst[di] = val
            di += 1
            si += 1
            if ref != 0 and val == 0xFC:
                if di >= dst_end:
                    return False, b""
                dst[di] = 0xFF
                di += 1
            continue
        si += best
        ctx = int.from_bytes(src[si - 4:si], "little")
        dst[di] = 0xFC
        di += 1
        best -= 64
        while best >= 254:
            best -= 254
            dst[di] = 0xFE
            di += 1
            if di >= dst_end:
                break
        if di >= dst_end:
            return False, b""
        dst[di] = best
        di += 1
    while si < src_end and di < dst_end:
        h = ((0x7FEB352D * ctx) & M32) >> 16
        ref = hashes[h]
        hashes[h] = si
        val = src[si]
        ctx = ((ctx << 8) | val) & M32
        dst[di] = val
        di += 1
        si += 1
        if ref != 0 and val == 0xFC:
            if di >= dst_end:
                return False, b""
            dst[di] = 0xFF
            di += 1
    return si == count and di < dst_end, bytes(dst[:di])


def lzp_inverse(src, dst_end):
    """-> (ok, out); dst_end = the destination slice's length"""
    M32 = 0xFFFFFFFF
    count = len(src)
    if count == 0:
        return True, b""
    if dst_end < count:
        return False, b""
    hashes = [0] * 65536
    dst = bytearray(src[0:4])
    ctx = int.from_bytes(dst[0:4], "little")
    si = 4
    while si < count:
        h = ((0x7FEB352D * ctx) & M32) >> 16
        ref = hashes[h]
        hashes[h] = len(dst)
        if ref == 0 or src[si] != 0xFC:
            if len(dst) >= dst_end:
                return False, bytes(dst)
            dst.append(src[si])
            ctx = ((ctx << 8) | src[si]) & M32
            si += 1
            continue
        si += 1
        if si >= count:
            return False, bytes(dst)
        if src[si] == 0xFF:
            if len(dst) >= dst_end:
                return False, bytes(dst)
            dst.append(0xFC)
            ctx = ((ctx << 8) | 0xFC) & M32
            si += 1
            continue
        m_len = 64
        if src[si] == 0xFE:
            while si < count and src[si] == 0xFE:
                si += 1
                m_len += 254
            if si >= count:
                return False, bytes(dst)
        m_len += src[si]
        si += 1
        if len(dst) + m_len > dst_end:
            return False, bytes(dst)
        for i in range(m_len):
            dst.append(dst[ref + i])
        ctx = int.from_bytes(dst[-4:], "little")
    return si == count, bytes(dst)


def _lzp_cases():
    from kanzi_b200 import synth
    r = np.random.default_rng(41)
    t = synth.text(20000, 5).tobytes()
    noise = bytes(r.integers(0, 256, 3000, dtype=np.uint8))
    flags = bytes(r.choice(np.array([0xFC, 0xFE, 0xFF, 0x41], dtype=np.uint8), 2000))
    return [t + t[500:9000] + t[:6000], noise + noise + flags + noise[:1500] + flags, bytes(5000), (b"0123456789abcdef" * 7 + b"\xfc") * 300,
            synth.records(30000, 7).tobytes(), noise, b"\xfc" * 700, t[:127], t[:128], (noise[:300] + b"\xfc\xfc") * 40, synth.exe_like(30000, 6).tobytes() * 2]


def test_lzp_agrees_with_the_oracle_both_ways():
    applied = 0
    for d in _lzp_cases():
        ok_ref, ref, used, _ = O.transform("LZP", d)
        ok, got = lzp_forward(d)
        assert int(ok) == ok_ref, (len(d), ok, ok_ref)
        if not ok:
            continue
        applied += 1
        assert got == ref and used == len(d)
        ok_i, back_ref, used_i, _ = O.transform("LZP", ref, inverse=True, dst_cap=len(d), dst_len=len(d))
        ok_p, back = lzp_inverse(ref, len(d))
        assert ok_i == 1 and ok_p and back_ref == d and back == d and used_i == len(ref)
        # a destination one byte short fails the same way in both; a truncated stream is accepted or refused alike
        assert O.transform("LZP", ref, inverse=True, dst_cap=len(d) - 1, dst_len=len(d) - 1)[0] == int(lzp_inverse(ref, len(d) - 1)[0]) == 0
        cut = ref[:len(ref) * 2 // 3]
        o2 = O.transform("LZP", cut, inverse=True, dst_cap=len(d), dst_len=len(d))
        p2 = lzp_inverse(cut, len(d))
        assert o2[0] == int(p2[0]) and (not p2[0] or o2[1] == p2[1])
    assert applied >= 6


# ---- RLT: K/transform/RLT.java:62-231 (forward), :233-249 (emitRunLength), :252-352 (inverse) --------------------------------
def _detect_simple_type(count, f):   # Global.java:556-608 -> DataType ordinal (DNA 6, NUMERIC 4, BASE64 5, BIN 7, SMALL_ALPHABET 9, UNDEFINED 0)
    if count == 0:
        return 0
    if sum(f[c] for c in b"acgntuACGNTU") > count - count // 12:
        return 6
    if sum(f[c] for c in b"0123456789+-*/=,.:; ") == count:
        return 4
    if (1 if f[0x3D] == 1 else 0) + sum(f[c] for c in b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/") == count:
        return 5
    n = sum(1 for x in f if x > 0)
    return 7 if n == 256 else (9 if n <= 4 else 0)


def rlt_forward(src, entropy="NONE", data_type=0, dst_end=None):
    """-> (ok, out, data type after the call); dst_end = dst.array.length"""
    count = len(src)
    if count == 0:
        return True, b"", data_type
    if count < 16:
        return False, b"", data_type
    if dst_end is None:
        dst_end = count + 32 if count <= 512 else count
    if data_type in (6, 5, 8):
        return False, b"", data_type
    escape = 0xFB
    if entropy not in ("NONE", "ANS0", "HUFFMAN", "RANGE"):
        f = [0] * 256
        for b in src:
            f[b] += 1
        if data_type == 0:
            data_type = _detect_simple_type(count, f)
            if data_type in (6, 5, 8):
                return False, b"", data_type
        m = 0
        if f[0] > 0:
            for i in range(1, 256):
                if f[i] < f[m]:
                    m = i
                    if f[i] == 0:
                        break
        escape = m
    dst = bytearray()
    si, end4 = 0, count - 4
    res, run = True, 0
    prev = src[si]
    si += 1
    dst += bytes([escape, prev])
    if prev == escape:
        dst.append(0)
    while True:
        again = False
        for _ in range(4):
            if prev != src[si]:
                break
            si += 1
            run += 1
        else:
            again = run < 73469 and si < end4
        if again:
            continue
        if run > 3:
            if len(dst) + 6 >= dst_end:
                res = False
                break
            dst.append(prev)
            if prev == escape:
                dst.append(0)
            dst.append(escape)
            r = run - 3
            if r >= 224:
                if r < 7936:
                    r -= 224
                    dst.append(224 + (r >> 8))
                else:
                    r -= 7936
                    dst += bytes([0xFF, (r >> 8) & 0xFF])
            dst.append(r & 0xFF)
        elif prev != escape:
            if len(dst) + run >= dst_end:
                res = False
                break
            dst += bytes([prev]) * run
        else:
            if len(dst) + 2 * run >= dst_end:
                res = False
                break
            dst += bytes([escape, 0]) * run
        prev = src[si]
        si += 1
        run = 1
        if si >= end4:
            break
    if res:
        if prev != escape:
            if len(dst) + run < dst_end:
                dst += bytes([prev]) * run
        elif len(dst) + 2 * run < dst_end:
            dst += bytes([escape, 0]) * run
        while si < count and len(dst) < dst_end:
            if src[si] == escape:
                if len(dst) + 2 >= dst_end:
                    res = False
                    break
                dst += bytes([escape, 0])
                si += 1
                continue
            dst.append(src[si])
            si += 1
        res = res and si == count
    res = res and len(dst) < si
    return res, bytes(dst), data_type


def rlt_inverse(src, dst_end):
    """-> (ok, out); None for ok where the Java code would throw"""
    count = len(src)
    if count == 0:
        return True, b""
    try:
        dst = bytearray()
        si = 0
        escape = src[si]
        si += 1
        if src[si] == escape:
            si += 1
            if si < count and src[si] != 0:
                return False, b""
            dst.append(escape)
            si += 1
        res = True
        while si < count:
            if src[si] != escape:
                if len(dst) >= dst_end:
                    break
                dst.append(src[si])
                si += 1
                continue
            si += 1
            if si >= count:
                res = False
                break
            val = dst[len(dst) - 1] if len(dst) > 0 else [][0]
            run = src[si]
            si += 1
            if run == 0:
                if len(dst) >= dst_end:
                    break
                dst.append(escape)
                continue
            if run == 0xFF:
                if si >= count - 1:
                    res = False
                    break
                run = ((src[si] << 8) | src[si + 1]) + 7936
                si += 2
            elif run >= 224:
                if si >= count:
                    res = False
                    break
                run = (((run - 224) << 8) | src[si]) + 224
                si += 1
            run += 2
            if len(dst) + run > dst_end or run > 73473:
                res = False
                break
            dst += bytes([val]) * run
        return res and si == count, bytes(dst)
    except IndexError:
        return None, b""


def _rlt_cases():
    from kanzi_b200 import synth
    r = np.random.default_rng(51)
    runs = bytes(np.repeat(r.integers(0, 256, 4000, dtype=np.uint8), r.integers(1, 40, 4000)))
    long_runs = bytes(np.repeat(r.integers(0, 6, 40, dtype=np.uint8), r.integers(1, 90000, 40)))
    esc = bytes(np.repeat(r.choice(np.array([0xFB, 0xFB, 0, 7, 0xFF], dtype=np.uint8), 3000), r.integers(1, 9, 3000)))
    return [runs, long_runs, esc, bytes(80000), b"\xfb" * 80000, bytes([0xFB]) + bytes(50), synth.text(30000, 3).tobytes(), synth.records(40000, 4).tobytes(),
            b"ab" * 8, b"a" * 16, b"a" * 15, b"abcdefgh" * 40 + b"z" * 73480 + b"q" * 5 + b"y" * 300, b"y" * 7 + b"z" * (73473 + 3) + b"x",
            bytes(r.integers(0, 2, 5000, dtype=np.uint8)), b"ACGT" * 3000 + b"A" * 900, b"0123456789" * 500 + b"7" * 99, runs[:517], runs[:511], runs[:512] + b"\x00" * 4]


@pytest.mark.parametrize("entropy", ["NONE", "FPAQ"])
def test_rlt_agrees_with_the_oracle_both_ways(entropy):
    eid = {"NONE": 0, "FPAQ": 2}[entropy]
    applied = 0
    for d in _rlt_cases():
        ok_ref, ref, used, cv = O.transform("RLT", d, ctx=[7, max(len(d), 1024), len(d), 1, 0, eid << 8])
        ok, got, dt = rlt_forward(d, entropy)
        assert int(ok) == ok_ref and dt == cv[4], (len(d), ok, ok_ref, dt, cv[4])
        if not ok:
            continue
        applied += 1
        assert got == ref and used == len(d), (len(d), len(got), len(ref))
        for cap in (len(d), len(d) + 100, len(d) - 1):
            o = O.transform("RLT", ref, inverse=True, dst_cap=cap, dst_len=cap)
            p = rlt_inverse(ref, cap)
            assert o[0] == (-1 if p[0] is None else int(p[0])) and (o[0] != 1 or o[1] == p[1]), (len(d), cap)
        assert O.transform("RLT", ref, inverse=True, dst_cap=len(d), dst_len=len(d))[1] == d
        for cut in (len(ref) // 2, len(ref) - 1, 3, 1):
            o = O.transform("RLT", ref[:cut], inverse=True, dst_cap=len(d), dst_len=len(d))
            p = rlt_inverse(ref[:cut], len(d))
            assert o[0] == (-1 if p[0] is None else int(p[0])) and (o[0] != 1 or o[1] == p[1]), (len(d), "cut", cut)
    assert applied >= 8


# ---- ROLZX: K/transform/ROLZCodec.java:1176-1294 (ROLZCodec2.forward), :1114-1173 (findMatch), :1296-1414 (inverse),
#      ROLZEncoder :1431-1597, ROLZDecoder :1599-1770 ---------------------------------------------------------------------------------
class _RolzCoder:
    """the adaptive binary arithmetic coder; Java `long` arithmetic = 64-bit wrap, kept with explicit masks"""
    def __init__(self, buf, index, decode):
        self.low, self.high = 0, 0x00FFFFFFFFFFFFFF
        self.probs = [[0xFFFF >> 1] * (256 << 9), [0xFFFF >> 1] * (256 << 5)]      # LITERAL_CTX = 0 (9 bits), MATCH_CTX = 1 (logPosChecks = 5)
        self.shift = (9, 5)
        self.buf, self.index = buf, index
        self.c1, self.ctx, self.p = 1, 0, self.probs[0]
        if decode:
            self.current = int.from_bytes(bytes(buf[index:index + 8]).ljust(8, b"\0"), "big")
            if index + 8 > len(buf):
                raise IndexError
            self.index += 8

    def set_context(self, n, byte):
        self.p = self.probs[n]
        self.ctx = byte << self.shift[n]

    def encode_bit(self, bit):
        i = self.ctx + self.c1
        pr = self.p[i]
        split = ((((self.high - self.low) & M64) >> 4) * (pr >> 4) & M64) >> 8
        if bit == 0:
            self.low = (self.low + split + 1) & M64
            self.p[i] = pr - (pr >> 5)
            self.c1 += self.c1
        else:
            self.high = (self.low + split) & M64
            self.p[i] = pr - (((pr - 0xFFFF) >> 5) + 1)
            self.c1 += self.c1 + 1
        while ((self.low ^ self.high) >> 24) == 0:
            if self.index + 4 > len(self.buf):
                raise IndexError                                      # Java: ArrayIndexOutOfBoundsException in writeInt32
            self.buf[self.index:self.index + 4] = ((self.high >> 32) & 0xFFFFFFFF).to_bytes(4, "big")
            self.index += 4
            self.low = (self.low << 32) & M64
            self.high = ((self.high << 32) | 0xFFFFFFFF) & M64

    def encode(self, val, nbits):
        self.c1 = 1
        for k in range(nbits - 1, -1, -1):
            self.encode_bit(val & (1 << k))

    def finish(self):
        if self.index + 8 > len(self.buf):
            raise IndexError
        self.buf[self.index:self.index + 8] = self.low.to_bytes(8, "big")
        self.index += 8

    def decode_bit(self):
        i = self.ctx + self.c1
        pr = self.p[i]
        mid = (self.low + (((((self.high - self.low) & M64) >> 4) * (pr >> 4) & M64) >> 8)) & M64
        signed = lambda v: v - (1 << 64) if v >> 63 else v
        if signed(mid) >= signed(self.current):
            self.high = mid
            self.p[i] = pr - (((pr - 0xFFFF) >> 5) + 1)
            self.c1 += self.c1 + 1
        else:
            self.low = (mid + 1) & M64
            self.p[i] = pr - (pr >> 5)
            self.c1 += self.c1
        while ((self.low ^ self.high) >> 24) == 0:
            self.low = (self.low << 32) & 0x00FFFFFFFFFFFFFF
            self.high = ((self.high << 32) | 0xFFFFFFFF) & 0x00FFFFFFFFFFFFFF
            if self.index + 4 > len(self.buf):
                raise IndexError
            self.current = ((self.current << 32) | int.from_bytes(self.buf[self.index:self.index + 4], "big")) & 0x00FFFFFFFFFFFFFF
            self.index += 4

    def decode(self, nbits):
        self.c1 = 1
        for _ in range(nbits):
            self.decode_bit()
        return self.c1 & ((1 << nbits) - 1)


def _rolz_key(buf, i, mm):
    if mm == 3:
        return buf[i] | (buf[i + 1] << 8)
    return (((int.from_bytes(buf[i:i + 8], "little") * 200002979) & M64) >> 40) & 0xFFFF


def rolzx_forward(src, data_type=0):
    """-> (ok, out, data type after the call) for a non-null ctx; ok None where the Java code would throw"""
    M32, HMASK = 0xFFFFFFFF, 0xFF000000
    count = len(src)
    if count == 0:
        return True, b"", data_type
    if count < 64:
        return False, b"", data_type
    src_end = count - 4
    dst = bytearray(count + 1024 if count <= 16384 else count + count // 32)
    dst[0:4] = count.to_bytes(4, "big")
    if data_type == 0:
        f = [0] * 256
        for b in src:
            f[b] += 1
        data_type = _detect_simple_type(count, f)
    mm, dt, flags = 3, 2, 0
    if data_type == 3:
        dt, flags = 3, 8
    elif data_type == 6:
        dt, mm, flags = 8, 7, 4
    dst[4] = flags
    re = _RolzCoder(dst, 5, False)
    counters = [0] * 65536
    size_chunk = min(count, 16 << 20)
    start = 0
    si = 0
    try:
        while start < src_end:
            matches = [0] * (65536 << 5)
            end_chunk = min(start + size_chunk, src_end)
            si = start
            re.set_context(0, 0)
            for _ in range(min(src_end - start, 8)):
                re.encode((1 << 8) | src[si], 9)
                si += 1
            while si < end_chunk:
                re.set_context(0, src[si - 1])
                key = _rolz_key(src, si - dt, mm)
                base = key << 5
                h32 = ((((int.from_bytes(src[si:si + 4], "little") << 8) & M32) * 200002979) & M32) & HMASK
                counter = counters[key]
                best_len, best_idx = 0, -1
                max_match = min(3 + 255, end_chunk - si) - 8
                for i in range(counter, counter - 32, -1):
                    ref = matches[base + (i & 31)]
                    if (ref & HMASK) != h32:
                        continue
                    ref = (ref & 0xFFFFFF) + start
                    if src[ref + best_len] != src[si + best_len]:
                        continue
                    n = 0
                    while n < max_match:
                        x = int.from_bytes(src[ref + n:ref + n + 8], "little") ^ int.from_bytes(src[si + n:si + n + 8], "little")
                        if x:
                            n += ((x & -x).bit_length() - 1) >> 3
                            break
                        n += 8
                    if n > best_len:
                        best_idx, best_len = counter - i, n
                        if best_len == max_match:
                            break
                counters[key] = (counters[key] + 1) & 31
                matches[base + counters[key]] = h32 | (si - start)
                if best_len < mm:
                    re.encode((1 << 8) | src[si], 9)
                    si += 1
                    continue
                re.encode(best_len - mm, 9)                     # MATCH_FLAG = 0 in bit 8
                re.set_context(1, src[si - 1])
                re.encode(best_idx, 5)
                si += best_len
            start = end_chunk
        for _ in range(4):
            re.set_context(0, src[si - 1])
            re.encode((1 << 8) | src[si], 9)
            si += 1
        re.finish()
        if re.index > len(dst):
            return None, b"", data_type
    except IndexError:
        return None, b"", data_type
    return si == src_end + 4, bytes(dst[:re.index]), data_type


def rolzx_inverse(src, dst_len):
    """-> (ok, out); ok None where the Java code would throw.  dst_len = dst slice length = array length here"""
    count = len(src)
    if count == 0:
        return True, b""
    try:
        if count < 5:
            raise IndexError
        sz = int.from_bytes(src[0:4], "big")
        if sz >= 1 << 31:
            sz -= 1 << 32
        if sz <= 0 or sz > dst_len:
            return False, b""
        flags = src[4]
        mm, dt = 3, 2
        if (flags & 0x0E) == 8:
            dt = 3
        elif (flags & 0x0E) == 4:
            dt, mm = 8, 7
        rd = _RolzCoder(src, 5, True)
        counters = [0] * 65536
        dst = bytearray()
        size_chunk = min(sz, 16 << 20)
        start = 0
        while start < sz:
            matches = [0] * (65536 << 5)
            end_chunk = min(start + size_chunk, sz)
            base0 = len(dst)                                      # output.index
            rd.set_context(0, 0)
            for _ in range(min(sz - start, 8)):
                v = rd.decode(9)
                if (v >> 8) == 0:
                    return False, bytes(dst)
                dst.append(v & 0xFF)
            while len(dst) < end_chunk:
                saved = len(dst)
                if saved - dt < 0:
                    raise IndexError
                key = _rolz_key(bytes(dst[saved - dt:saved - dt + 8]).ljust(8, b"\0"), 0, mm)
                base = key << 5
                rd.set_context(0, dst[-1])
                v = rd.decode(9)
                if (v >> 8) == 1:
                    dst.append(v & 0xFF)
                else:
                    m_len = v & 0xFF
                    if len(dst) + m_len + 3 > sz:
                        return False, bytes(dst)
                    rd.set_context(1, dst[-1])
                    m_idx = rd.decode(5)
                    ref = base0 + matches[base + ((counters[key] - m_idx) & 31)]
                    if len(dst) + m_len + mm > dst_len:
                        raise IndexError
                    for k in range(m_len + mm):
                        dst.append(dst[ref + k])
                counters[key] = (counters[key] + 1) & 31
                matches[base + counters[key]] = saved - base0
            start = end_chunk
        return rd.index == count, bytes(dst)
    except IndexError:
        return None, b""


def _rolzx_cases():
    from kanzi_b200 import synth
    r = np.random.default_rng(61)
    dna = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[r.integers(0, 4, 3000)])
    return [synth.text(12000, 5).tobytes(), synth.exe_like(15000, 6).tobytes(), synth.records(10000, 7).tobytes(), (b"0123456789abcdef" * 7 + b"Z") * 120,
            bytes(20000), dna + dna[500:2500] + dna[:900], bytes(r.integers(0, 256, 3000, dtype=np.uint8)), b"ab" * 32, b"xyz" * 21, b"q" * 63,
            bytes(r.integers(0, 3, 5000, dtype=np.uint8))]


def test_rolzx_agrees_with_the_oracle_both_ways():
    applied = 0
    for d in _rolzx_cases():
        ok_ref, ref, used, cv = O.transform("ROLZX", d)
        ok, got, dt = rolzx_forward(d)
        assert ok is not None and int(ok) == ok_ref and dt == cv[4], (len(d), ok, ok_ref, dt, cv[4])
        if not ok:
            continue
        applied += 1
        assert got == ref and used == len(d), (len(d), len(got), len(ref))
        o = O.transform("ROLZX", ref, inverse=True, dst_cap=len(d), dst_len=len(d))
        p = rolzx_inverse(ref, len(d))
        assert o[0] == 1 and p[0] is True and o[1] == d and p[1] == d and o[2] == len(ref)
        assert O.transform("ROLZX", ref, inverse=True, dst_cap=len(d) - 1, dst_len=len(d) - 1)[0] == 0 and rolzx_inverse(ref, len(d) - 1)[0] is False
        for cut in (len(ref) // 2, len(ref) - 4):
            o = O.transform("ROLZX", ref[:cut], inverse=True, dst_cap=len(d), dst_len=len(d))
            p = rolzx_inverse(ref[:cut], len(d))
            assert o[0] == (-1 if p[0] is None else int(p[0])), (len(d), cut, o[0], p[0])
    assert applied >= 8


def test_rolzx_on_incompressible_data_overruns_its_buffer_in_both():
    """Random bytes cost ROLZX's literal coder about nine bits each (256 x 512 adaptive cells never settle on 20 000 samples), more than
    the n + n/32 bytes getMaxEncodedLength grants (ROLZCodec.java:1417-1421): the reference's writeInt32 then throws
    ArrayIndexOutOfBoundsException and EncodingTask reports "Error in block" (COS:1042-1045).  Oracle and restatement both say so."""
    d = bytes(np.random.default_rng(1).integers(0, 256, 20000, dtype=np.uint8))
    assert O.transform("ROLZX", d)[0] == -1
    assert rolzx_forward(d)[0] is None
    with pytest.raises(RuntimeError):
        O.compress(d, ["ROLZX"], "NONE", 65536)


# ---- RANGE: K/entropy/RangeEncoder.java:160-301 (not on the CUDA path; the oracle carries it for RLT's ctx["entropy"] and for what comes next)
def range_encode(data, chunk=1 << 15, log_range=12):
    TOP, BOTTOM, MASK = 0x0FFFFFFFFFFFFFFF, 0xFFFF, 0x0FFFFFFF00000000
    out = _Bits()
    n = len(data)
    start = 0
    signed = lambda v: v - (1 << 64) if v >> 63 else v
    while start < n:
        end = min(start + chunk, n)
        rng, low = TOP, 0
        lr = log_range
        while lr > 8 and (1 << lr) > end - start:
            lr -= 1
        f = [0] * 257
        for b in data[start:end]:
            f[b] += 1
        alphabet = _normalize(f, end - start, 1 << lr)
        _encode_alphabet(out, alphabet)
        if len(alphabet) > 0:
            out.write(lr - 8, 3)
            chk = 8 if len(alphabet) >= 64 else 6
            llr = 3
            while (1 << llr) <= lr:
                llr += 1
            for i in range(1, len(alphabet), chk):
                endj = min(i + chk, len(alphabet))
                mx = max(f[alphabet[j]] - 1 for j in range(i, endj))
                log_max = 0
                while (1 << log_max) <= mx:
                    log_max += 1
                out.write(log_max, llr)
                if log_max == 0:
                    continue
                for j in range(i, endj):
                    out.write(f[alphabet[j]] - 1, log_max)
        if len(alphabet) <= 1:
            start = end
            continue
        cum = [0] * 257
        for i in range(256):
            cum[i + 1] = cum[i] + f[i]
        for b in data[start:end]:
            rng >>= lr
            low = (low + cum[b] * rng) & M64
            rng = (rng * (cum[b + 1] - cum[b])) & M64
            while True:
                if ((low ^ ((low + rng) & M64)) & MASK) != 0:
                    if signed(rng) > BOTTOM:
                        break
                    rng = (-low) & BOTTOM
                out.write(low >> 32, 28)
                rng = (rng << 28) & M64
                low = (low << 28) & M64
        out.write(low, 60)
        start = end
    return out.bytes()


def test_range_encoder_agrees_with_the_oracle():
    import corpus
    from kanzi_b200 import synth
    r = np.random.default_rng(13)
    cases = [synth.text(70000, 3).tobytes(), synth.exe_like(40000, 4).tobytes(), bytes(50000), bytes(r.integers(0, 256, 33000, dtype=np.uint8)),
             bytes(r.integers(0, 2, 9000, dtype=np.uint8)), b"abc" * 11, b"x", bytes(r.integers(0, 70, 300, dtype=np.uint8)), corpus.fibonacci_chunk()[:40000],
             bytes(r.integers(0, 256, 1 << 15, dtype=np.uint8)) + b"tail"]
    for d in cases:
        ref, ref_bits = O.entropy_encode("RANGE", d)
        got, bits = range_encode(d)
        assert bits == ref_bits and got == ref, (len(d), bits, ref_bits)
        out, rr, used = O.entropy_decode("RANGE", ref, ref_bits, len(d))
        assert rr == len(d) and out == d and used == ref_bits


# ---- LZ / LZX inverse: K/transform/LZCodec.java:605-756 (inverseV6), readLength :241-258 ---------------------------------------------
def lz_inverse(src, dst_end):
    """-> (ok, out); ok None where the Java code would throw.  dst_end = dst.array.length"""
    count = len(src)
    if count == 0:
        return True, b""
    if count < 13:
        return False, b""
    le = lambda i: int.from_bytes(src[i:i + 4], "little", signed=True)
    tk_len, m_idx_len, m_len_len = le(0), le(4), le(8)
    if tk_len < 0 or m_idx_len < 0 or m_len_len < 0:
        return False, b""
    if tk_len < 13 or tk_len > count or m_idx_len > count - tk_len or m_len_len > count - tk_len - m_idx_len:
        return False, b""
    tk = tk_len
    m_idx = tk + m_idx_len
    m_len_idx = m_idx + m_len_len
    src_end, lit_end = tk - 13, tk
    max_dist = (1 << 16) - 2 if (src[12] & 1) == 0 else (1 << 24) - 2
    min_match = ((src[12] >> 1) & 7) + 2
    si = 13
    dst = bytearray(dst_end)
    di = 0
    repd0 = repd1 = count

    def read_length(i):
        res = src[i]
        i += 1
        if res < 254:
            return res, i
        if res == 254:
            return res + (src[i] << 8) + src[i + 1], i + 2
        return res + (src[i] << 16) + (src[i + 1] << 8) + src[i + 2], i + 3

    try:
        while True:
            token = src[tk]
            tk += 1
            if token >= 32:
                if token >= 0xE0:
                    ln, si = read_length(si)
                    lit = 7 + ln
                else:
                    lit = token >> 5
                if lit > dst_end - di or lit > lit_end - si:
                    return False, bytes(dst[:di])
                if si + lit < src_end:                            # emitLiterals (:945-950): eight bytes at a time, up to 7 past the run on both sides
                    padded = (lit + 7) & ~7
                    if si + padded > len(src) or di + padded > dst_end:
                        raise IndexError
                elif si + lit > len(src):
                    raise IndexError
                dst[di:di + lit] = src[si:si + lit]
                si += lit
                di += lit
                if si >= src_end:
                    break
            f = token & 0x18
            if f == 0:
                ml = token & 3
                if ml == 3:
                    ln, m_len_idx = read_length(m_len_idx)
                    ml += min_match + ln
                else:
                    ml += min_match
                dist = repd0 if (token & 4) == 0 else repd1
            else:
                ml = token & 7
                if ml == 7:
                    ln, m_len_idx = read_length(m_len_idx)
                    ml += min_match + ln
                else:
                    ml += min_match
                dist = src[m_idx]
                m_idx += 1
                if f == 0x18:
                    dist = (dist << 16) | (src[m_idx] << 8) | src[m_idx + 1]
                    m_idx += 2
                elif f == 0x10:
                    dist = (dist << 8) | src[m_idx]
                    m_idx += 1
            repd1, repd0 = repd0, dist
            m_end = di + ml
            ref = di - dist
            if ref < 0 or dist > max_dist or m_end > dst_end:
                return False, bytes(dst[:di])
            if dist >= 16:
                while True:                                       # sixteen bytes at a time, past mEnd if need be (the array must hold them)
                    if di + 16 > dst_end:
                        raise IndexError
                    dst[di:di + 16] = dst[ref:ref + 16]
                    ref += 16
                    di += 16
                    if di >= m_end:
                        break
            else:
                for i in range(ml):
                    dst[di + i] = dst[ref + i]
            di = m_end
    except IndexError:
        return None, b""
    return si == src_end + 13, bytes(dst[:di])


@pytest.mark.parametrize("name", ["LZ", "LZX"])
def test_lz_inverse_agrees_with_the_oracle(name):
    import corpus
    from kanzi_b200 import synth
    r = np.random.default_rng(23)
    cases = [synth.text(30000, 5).tobytes(), synth.exe_like(40000, 6).tobytes(), synth.records(25000, 7).tobytes(), (b"0123456789abcdef" * 7 + b"Z") * 300,
             corpus.sparse_with_repeats(60000, 14), bytes(70000), synth.text(70000, 9).tobytes() * 2, b"ACGT" * 4000 + b"TTGACA" * 900]
    done = 0
    for d in cases:
        ok, enc, _, _ = O.transform(name, d, dst_cap=len(d) + len(d) // 64 + 1100, ctx=[7, max(len(d), 1024), len(d), 1, 0, 0])
        if ok != 1:
            continue
        done += 1
        for cap in (len(d) + 16, len(d) + 1000, len(d), len(d) - 1):      # (the 16-byte copies want slack after the last match)
            o = O.transform(name, enc, inverse=True, dst_cap=cap, dst_len=cap, ctx=[7, max(len(d), 1024), len(d), 1, 0, 0])
            p = lz_inverse(enc, cap)
            assert o[0] == (-1 if p[0] is None else int(p[0])), (name, len(d), cap, o[0], p[0])
            if o[0] == 1:
                assert o[1] == p[1] == d
        for k in range(12):                                        # damaged streams: refused or accepted alike, same bytes when accepted
            bad = bytearray(enc)
            bad[int(r.integers(0, len(bad)))] ^= 1 << int(r.integers(0, 8))
            o = O.transform(name, bytes(bad), inverse=True, dst_cap=len(d) + 16, dst_len=len(d) + 16, ctx=[7, max(len(d), 1024), len(d), 1, 0, 0])
            p = lz_inverse(bytes(bad), len(d) + 16)
            assert o[0] == (-1 if p[0] is None else int(p[0])), (name, len(d), "flip", k, o[0], p[0])
            if o[0] == 1:
                assert o[1] == p[1]
    assert done >= 6


# ---- rANS decoder: K/entropy/ANSRangeDecoder.java:148-196 (decode), :257-331 (decodeChunkV2), :333-408 (decodeHeader); bsVersion >= 4 ----
class _BitsIn:
    def __init__(self, data, nbits):
        self.v = int.from_bytes(data, "big")
        self.total = 8 * len(data)
        self.nbits, self.pos = nbits, 0

    def read(self, n):
        if n == 0:
            return 0
        if self.pos + n > ((self.nbits + 7) & ~7):
            raise EOFError
        r = (self.v >> (self.total - self.pos - n)) & ((1 << n) - 1)
        self.pos += n
        return r

    def varint(self):          # EntropyUtils.readVarInt (K/entropy/EntropyUtils.java:284-300)
        value = self.read(8)
        res = value & 0x7F
        shift = 7
        while value >= 128 and shift <= 28:
            value = self.read(8)
            res |= (value & 0x7F) << shift
            shift += 7
        return res


def _decode_alphabet(b):       # EntropyUtils.decodeAlphabet (:86-122)
    if b.read(1) == 0:
        return [] if b.read(1) == 1 else list(range(256))
    last = b.read(5)
    out = []
    for i in range(last + 1):
        m = b.read(8)
        out += [(i << 3) + j for j in range(8) if m & (1 << j)]
    return out


def ans_decode(payload, nbits, n, order, chunk=None, b=None):
    """-> (return value, bytes, bits consumed); a BitStreamException / index error of the Java code shows as return value None.
    `b`: an open bit reader to continue from (the ROLZ codec decodes several arrays from one stream, with its own chunk size)."""
    b = _BitsIn(payload, nbits) if b is None else b
    out = bytearray(n)
    if n <= 32:
        for i in range(n):
            out[i] = b.read(8)
        return n, bytes(out), b.pos
    dim = 255 * order + 1
    chunk = (16384 if chunk is None else chunk) << (8 * order)
    freqs = [[0] * 256 for _ in range(dim)]
    f2s = [[] for _ in range(dim)]
    sym = [[(0, 0)] * 256 for _ in range(dim)]                   # (cumFreq, freq)
    start = 0
    try:
        while start < n:
            end = min(start + chunk, n)
            lr = 8 + b.read(3)
            scale = 1 << lr
            total_alpha = 0
            alphabet = []
            for k in range(dim):
                alphabet = _decode_alphabet(b)
                if not alphabet:
                    continue
                llr = 3
                while (1 << llr) <= lr:
                    llr += 1
                f = freqs[k]
                if len(alphabet) != 256:
                    for i in range(256):
                        f[i] = 0
                if len(f2s[k]) < scale:
                    f2s[k] = [0] * scale
                chk = 8 if len(alphabet) >= 64 else 6
                s = 0
                for i in range(1, len(alphabet), chk):
                    log_max = b.read(llr)
                    if (1 << log_max) > scale:
                        return None, b"", b.pos
                    for j in range(i, min(i + chk, len(alphabet))):
                        fr = 1 if log_max == 0 else 1 + b.read(log_max)
                        if fr <= 0 or fr >= scale:
                            return None, b"", b.pos
                        f[alphabet[j]] = fr
                        s += fr
                if scale <= s:
                    return None, b"", b.pos
                f[alphabet[0]] = scale - s
                s = 0
                for i in range(256):
                    if f[i] == 0:
                        continue
                    f2s[k][s:s + f[i]] = [i] * f[i]
                    sym[k][i] = (s, scale - 1 if f[i] >= scale else f[i])      # Symbol.reset (:576-579) "mirrors the encoder"
                    s += f[i]
                total_alpha += len(alphabet)
            if total_alpha == 0:
                return start, bytes(out), b.pos
            if order == 0 and total_alpha == 1:
                out[start:end] = bytes([alphabet[0]]) * (end - start)
                start = end
                continue
            sz = b.varint()
            if sz >= 1 << 27:
                break
            st = [b.read(32) for _ in range(4)]                  # st0 .. st3
            buf = bytearray(max(2 * (end - start), 256))
            for i in range(sz):
                v = b.read(8)
                if i >= len(buf):
                    raise IndexError
                buf[i] = v
            pos = 0
            mask = scale - 1

            def step(k, ctx):
                nonlocal pos
                cur = f2s[ctx][st[k] & mask]
                cum, fr = sym[ctx][cur]
                x = (fr * (st[k] >> lr) + (st[k] & mask) - cum) & 0xFFFFFFFF
                if x >= 1 << 31:
                    x -= 1 << 32                                  # Java int: the comparison below is signed
                if x < (1 << 15):
                    x = ((x << 8) | buf[pos]) & 0xFFFFFFFF
                    x = ((x << 8) | buf[pos + 1]) & 0xFFFFFFFF
                    pos += 2
                st[k] = x & 0xFFFFFFFF
                return cur

            end4 = start + ((end - start) & -4)
            if order == 0:
                for i in range(start, end4, 4):
                    for j, k in enumerate((3, 2, 1, 0)):
                        out[i + j] = step(k, 0)
            else:
                quarter = (end4 - start) >> 2
                prv = [0, 0, 0, 0]
                for i in range(quarter):
                    for k in (3, 2, 1, 0):
                        cur = step(k, prv[k])
                        out[start + k * quarter + i] = cur
                        prv[k] = cur
            for i in range(end4, end):
                out[i] = buf[pos]
                pos += 1
            if pos != sz:
                break
            start = end
    except (EOFError, IndexError):
        return None, b"", b.pos
    return n, bytes(out), b.pos


@pytest.mark.parametrize("kind,order", [("ANS0", 0), ("ANS1", 1)])
def test_ans_decoder_agrees_with_the_oracle(kind, order):
    from kanzi_b200 import synth
    r = np.random.default_rng(29)
    cases = [synth.text(40000, 3).tobytes(), synth.exe_like(30000, 4).tobytes(), bytes(20000), bytes(r.integers(0, 256, 17000, dtype=np.uint8)),
             bytes(r.integers(0, 2, 9000, dtype=np.uint8)), b"abc" * 11, b"x" * 33, b"hello", bytes(r.integers(0, 70, 300, dtype=np.uint8)), b"ab" * 8200 + b"xyz"]
    for d in cases:
        enc, bits = O.entropy_encode(kind, d)
        o = O.entropy_decode(kind, enc, bits, len(d))
        p = ans_decode(enc, bits, len(d), order)
        assert o[1] == len(d) and o[0] == d and p[0] == len(d) and p[1] == d and p[2] == o[2] == bits, (kind, len(d))
        for k in range(10):                                        # damaged payloads: same verdict, same bytes, same position
            bad = bytearray(enc)
            bad[int(r.integers(0, len(bad)))] ^= 1 << int(r.integers(0, 8))
            o = O.entropy_decode(kind, bytes(bad), bits, len(d))
            p = ans_decode(bytes(bad), bits, len(d), order)
            assert o[1] == (-1 if p[0] is None else p[0]), (kind, len(d), "flip", k, o[1], p[0])
            if p[0] is not None:
                assert o[0][:max(p[0], 0)] == p[1][:max(p[0], 0)] or p[0] == len(d)


# ---- FPAQ decoder: K/entropy/FPAQDecoder.java:88-176 (decode), :201-222 (decodeBitV2), :225-238 (read); bsVersion >= 4 -------------
def fpaq_decode(payload, nbits, n, chunk=4 << 20):
    """-> (return value, bytes, bits consumed); None as return value where the bitstream runs out (Java: BitStreamException)"""
    TOP = 0x00FFFFFFFFFFFFFF
    b = _BitsIn(payload, nbits)
    out = bytearray(n)
    if n == 0:
        return 0, b"", 0
    low, high = 0, TOP
    probs = [[65536 >> 1] * 256 for _ in range(4)]
    p = probs[0]
    start = 0
    try:
        while start < n:
            sz = b.varint()
            if sz >= 2 * n:
                return 0, bytes(out), b.pos
            current = b.read(56)
            buf = bytes(b.read(8) for _ in range(sz)) + bytes(max(sz + (sz >> 2), 1024) - sz)
            idx = 0
            end = start + min(chunk, n - start)
            p = probs[0]
            for i in range(start, end):
                ctx = 1
                for _ in range(8):
                    split = ((((high - low) >> 8) * p[ctx]) >> 8) + low
                    if split >= current:
                        high = split
                        p[ctx] -= (p[ctx] - 65536 + 64) >> 6
                        ctx = (ctx << 1) + 1
                    else:
                        low = split + 1
                        p[ctx] -= p[ctx] >> 6
                        ctx <<= 1
                    while ((low ^ high) & 0x00FFFFFFFF000000) == 0:
                        low = (low << 32) & TOP
                        high = ((high << 32) | 0xFFFFFFFF) & TOP
                        if idx + 4 > sz:
                            current = (current << 32) & TOP
                            idx = sz + 1
                        else:
                            current = ((current << 32) | int.from_bytes(buf[idx:idx + 4], "big")) & TOP
                            idx += 4
                out[i] = ctx & 0xFF
                if idx > sz:
                    return 0, bytes(out), b.pos
                p = probs[(ctx & 0xFF) >> 6]
            start = end
    except EOFError:
        return None, b"", b.pos
    return n, bytes(out), b.pos


def test_fpaq_decoder_agrees_with_the_oracle():
    from kanzi_b200 import synth
    r = np.random.default_rng(31)
    cases = [synth.text(30000, 3).tobytes(), synth.exe_like(20000, 4).tobytes(), bytes(9000), bytes(r.integers(0, 256, 7000, dtype=np.uint8)), b"a", b"hello world" * 3]
    for d in cases:
        enc, bits = O.entropy_encode("FPAQ", d)
        o = O.entropy_decode("FPAQ", enc, bits, len(d))
        p = fpaq_decode(enc, bits, len(d))
        assert o[1] == len(d) and o[0] == d and p[0] == len(d) and p[1] == d and p[2] == o[2] == bits, len(d)
        for k in range(8):
            bad = bytearray(enc)
            bad[int(r.integers(0, len(bad)))] ^= 1 << int(r.integers(0, 8))
            o = O.entropy_decode("FPAQ", bytes(bad), bits, len(d))
            p = fpaq_decode(bytes(bad), bits, len(d))
            assert o[1] == (-1 if p[0] is None else p[0]), (len(d), "flip", k, o[1], p[0])
            if p[0] == len(d):
                assert o[0] == p[1]
    # two chunks: probabilities and the interval carry over, `current` is read afresh (a small chunk size stands in for 4 MiB)
    d = synth.text(5000, 9).tobytes()
    assert fpaq_decode(*fpaq_encode(d, chunk=2048), len(d), chunk=2048)[1] == d


# ---- Huffman decoder, by definition: stream layout of K/entropy/HuffmanDecoder.java:262-294 (decodeV6), :110-148 (readLengths),
#      :297-360 (decodeChunk: four fragment bit strings, then the bytes that do not fill a fragment); canonical codes as the encoder's.
#      The reference decodes through a 12-bit table and a 56-bit window; on a well-formed stream a bit-by-bit prefix decoder must give
#      the same bytes and end at the same bit, which is what this checks (damaged streams are the GPU tests' business).
def huffman_decode(payload, nbits, n, chunk=16384):
    b = _BitsIn(payload, nbits)
    out = bytearray(n)
    start = 0
    while start < n:
        size = min(chunk, n - start)
        if size < 32:
            for i in range(size):
                out[start + i] = b.read(8)
            start += size
            continue
        alphabet = _decode_alphabet(b)
        assert alphabet
        sizes = {}
        cur = 2
        for s in alphabet:
            if b.read(1) == 0:                                   # ExpGolombDecoder.decodeByte, signed (:41-56)
                lg = 1
                while b.read(1) == 0:
                    lg += 1
                res = b.read(lg + 1)
                sgn = res & 1
                res = (res >> 1) + (1 << lg) - 1
                cur += -res if sgn else res
            assert 0 < cur <= 12
            sizes[s] = cur
        if len(alphabet) == 1:
            out[start:start + size] = bytes([alphabet[0]]) * size
            start += size
            continue
        order = sorted(alphabet, key=lambda s: (sizes[s], s))
        table = {}
        code, cs = 0, sizes[order[0]]
        for s in order:
            code <<= sizes[s] - cs
            cs = sizes[s]
            table[(cs, code)] = s
            code += 1
        frag_bits = [b.varint() for _ in range(4)]
        frag = size // 4
        for k in range(4):
            fb = _BitsIn(bytes((b.read(8) if 8 * i + 8 <= frag_bits[k] else b.read(frag_bits[k] - 8 * i) << (8 - (frag_bits[k] - 8 * i)))
                               for i in range((frag_bits[k] + 7) // 8)), frag_bits[k])
            for i in range(frag):
                ln, code = 0, 0
                while (ln, code) not in table:
                    code = (code << 1) | fb.read(1)
                    ln += 1
                    assert ln <= 12
                out[start + k * frag + i] = table[(ln, code)]
            assert fb.pos == frag_bits[k]
        for i in range(4 * frag, size):
            out[start + i] = b.read(8)
        start += size
    return bytes(out), b.pos


def test_huffman_decoder_by_definition_agrees_with_the_oracle():
    from kanzi_b200 import synth
    r = np.random.default_rng(37)
    cases = [synth.text(40000, 3).tobytes(), synth.exe_like(33001, 4).tobytes(), bytes(20000), bytes(r.integers(0, 256, 17003, dtype=np.uint8)),
             bytes(r.integers(0, 2, 9000, dtype=np.uint8)), b"abc" * 11, b"x" * 31, b"hello", bytes(r.integers(0, 70, 300, dtype=np.uint8)),
             synth.skewed(50000, 5, 3.0).tobytes(), synth.text(16384 + 31, 8).tobytes(), synth.text(16384 + 32, 8).tobytes()]
    for d in cases:
        enc, bits = O.entropy_encode("HUFFMAN", d)
        o = O.entropy_decode("HUFFMAN", enc, bits, len(d))
        got, used = huffman_decode(enc, bits, len(d))
        assert o[1] == len(d) and o[0] == d and got == d and used == o[2] == bits, len(d)


# ---- SRT inverse: K/transform/SRT.java:178-257, decodeHeader :335-353 ----------------------------------------------------------
def srt_inverse(src, dst_len):
    """-> (ok, out); ok None where the Java code would throw (an index past the input array)"""
    if len(src) == 0:
        return True, b""
    try:
        k = 0
        freqs = []
        for _ in range(256):
            val = src[k]
            k += 1
            res, shift = val & 0x7F, 7
            while val >= 128:
                val = src[k]
                k += 1
                res |= (val & 0x7F) << shift
                if shift > 21:
                    break
                shift += 7
            res &= 0xFFFFFFFF
            freqs.append(res - (1 << 32) if res >> 31 else res)
        count = len(src) - k
        if count > dst_len:
            return False, b""
        symbols = _srt_order(freqs)
        n_sym = len(symbols)
        buckets, ends, r2s = [0] * 256, [0] * 256, [0] * 256
        pos = 0
        for c in symbols:
            if k + pos < 0 or k + pos >= len(src):
                return False, b""
            r2s[src[k + pos]] = c
            buckets[c] = pos + 1
            pos += freqs[c]
            ends[c] = pos
        c = r2s[0]
        out = bytearray(max(count, 0))
        for i in range(count):
            out[i] = c
            if buckets[c] < ends[c]:
                r = src[k + buckets[c]]
                buckets[c] += 1
                if r == 0:
                    continue
                r2s[0:r] = r2s[1:r + 1]
                r2s[r] = c
                c = r2s[0]
            else:
                if n_sym == 1:
                    continue
                n_sym -= 1
                r2s[0:n_sym] = r2s[1:n_sym + 1]
                c = r2s[0]
        return True, bytes(out)
    except IndexError:
        return None, b""


def test_srt_inverse_agrees_with_the_oracle():
    r = np.random.default_rng(43)
    for d in _cases():
        if len(d) == 0:
            continue
        ok, enc, _, _ = O.transform("SRT", d)
        assert ok == 1
        o = O.transform("SRT", enc, inverse=True, dst_cap=len(d), dst_len=len(d))
        p = srt_inverse(enc, len(d))
        assert o[0] == 1 and p[0] is True and o[1] == p[1] == d, len(d)
        assert O.transform("SRT", enc, inverse=True, dst_cap=len(d) - 1, dst_len=len(d) - 1)[0] == int(srt_inverse(enc, len(d) - 1)[0]) == 0
        for k in range(6):                                        # damaged ranks / header bytes: same verdict, same bytes
            bad = bytearray(enc)
            bad[int(r.integers(0, len(bad)))] = int(r.integers(0, 256))
            o = O.transform("SRT", bytes(bad), inverse=True, dst_cap=len(d) + 300, dst_len=len(d) + 300)
            p = srt_inverse(bytes(bad), len(d) + 300)
            assert o[0] == (-1 if p[0] is None else int(p[0])), (len(d), k, o[0], p[0])
            if o[0] == 1:
                assert o[1] == p[1]


# ---- ROLZ inverse (ROLZCodec1): K/transform/ROLZCodec.java:696-960, readLength :962-980, emitCopy :162-180 --------------------------
def rolz_inverse(src, dst_len):
    """-> (ok, out); ok None where the Java code would throw.  bsVersion >= 4; dst_len = dst slice length = array length"""
    count = len(src)
    if count == 0:
        return True, b""
    try:
        sz = int.from_bytes(src[0:4], "big", signed=True) - 4
        if sz <= 0 or sz > dst_len:
            return False, b""
        size_chunk = min(sz, 16 << 20)
        lit_cap, len_cap, idx_cap, tk_cap = size_chunk, size_chunk // 5 + 4, size_chunk // 4, size_chunk // 4
        counters = [0] * 65536
        flags = src[4]
        lit_order = flags & 1
        mm, dt = 3, 2
        log_checks = flags >> 4
        if log_checks < 2 or log_checks > 8:
            return False, b""
        mask = (1 << log_checks) - 1
        sel = flags & 0x0E
        if sel == 2:
            mm, dt = 4, 8
        elif sel == 4:
            mm, dt = 7, 8
        elif sel == 8:
            dt = 3
        dst = bytearray()
        si = 5
        start = 0
        while start < sz:
            matches = [0] * (65536 << log_checks)
            end_chunk = min(start + size_chunk, sz)
            size_chunk = end_chunk - start
            base0 = len(dst)                                      # output.index
            b = _BitsIn(src[si:], 8 * (count - si))
            lit_len, tk_len, m_len_len, m_idx_len = (b.read(32) for _ in range(4))
            if any(v >= 1 << 31 for v in (lit_len, tk_len, m_len_len, m_idx_len)):
                return False, bytes(dst)
            if lit_len > lit_cap or tk_len > tk_cap or m_len_len > len_cap - 4 or m_idx_len > idx_cap:
                return False, bytes(dst)
            if lit_len < min(size_chunk, 8) or lit_len > size_chunk or (tk_len == 0 and m_idx_len != 0) or (tk_len > 0 and m_idx_len + 1 != tk_len):
                return False, bytes(dst)
            r = ans_decode(None, 0, lit_len, lit_order, b=b)
            if r[0] is None:
                return None, b""
            lit = r[1]
            parts = []
            for n in (tk_len, m_len_len, m_idx_len):
                r = ans_decode(None, 0, n, 0, chunk=32768, b=b)
                if r[0] is None:
                    return None, b""
                parts.append(r[1])
            tk, lens, midx = parts
            lens = lens + bytes(4)
            si += (b.pos + 7) >> 3
            if tk_len == 0:
                if lit_len != size_chunk:
                    return False, bytes(dst)
                dst += lit[:size_chunk]
                start = end_chunk
                continue
            li = ti = ni = mi = 0

            def read_length():
                nonlocal ni
                nxt = lens[ni]; ni += 1
                length = nxt & 0x7F
                for _ in range(3):
                    if not nxt & 0x80:
                        break
                    nxt = lens[ni]; ni += 1
                    length = (length << 7) | (nxt & 0x7F)
                return length

            for _ in range(min(sz - len(dst), 8)):
                dst.append(lit[li]); li += 1
            while len(dst) < end_chunk:
                token = tk[ti]; ti += 1
                match_len = token & 7
                if match_len == 7:
                    if ni >= m_len_len:
                        return False, bytes(dst)
                    match_len = read_length() + 7
                if token < 0xF8:
                    ll = token >> 3
                else:
                    if ni >= m_len_len:
                        return False, bytes(dst)
                    ll = read_length() + 31
                if ll > 0:
                    n0 = len(dst) - base0
                    if li + ll > len(lit) or len(dst) + ll > dst_len:
                        raise IndexError
                    dst += lit[li:li + ll]
                    j, src_inc = 0, 0
                    while j < ll:
                        key = _rolz_key(bytes(dst[base0 + n0 + j - dt:base0 + n0 + j - dt + 8]).ljust(8, b"\0"), 0, mm if mm == 3 else 7)
                        counters[key] = (counters[key] + 1) & mask
                        matches[(key << log_checks) + counters[key]] = n0 + j
                        j += src_inc >> 6
                        src_inc += 1
                        j += 1
                    li += ll
                    if len(dst) >= end_chunk:
                        if len(dst) == end_chunk:
                            break
                        return False, bytes(dst)
                if len(dst) + match_len + mm > sz:
                    return False, bytes(dst)
                p0 = len(dst)
                key = _rolz_key(bytes(dst[p0 - dt:p0 - dt + 8]).ljust(8, b"\0"), 0, mm if mm == 3 else 7)
                base = key << log_checks
                m_idx = midx[mi]; mi += 1
                ref = base0 + matches[base + ((counters[key] - m_idx) & mask)]
                for k in range(match_len + mm):
                    dst.append(dst[ref + k])
                counters[key] = (counters[key] + 1) & mask
                matches[base + counters[key]] = p0 - base0
            if ti != tk_len or mi != m_idx_len or li != lit_len or ni != m_len_len:
                return False, bytes(dst)
            start = end_chunk
        if len(dst) + 4 > dst_len or count - si != 4:
            return False, bytes(dst)
        dst += src[si:si + 4]
        return True, bytes(dst)
    except (IndexError, EOFError):
        return None, b""


def test_rolz_inverse_agrees_with_the_oracle():
    import corpus
    from kanzi_b200 import synth
    r = np.random.default_rng(47)
    dna = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[r.integers(0, 4, 6000)])
    cases = [synth.text(30000, 5).tobytes(), synth.exe_like(40000, 6).tobytes(), synth.records(25000, 7).tobytes(), (b"0123456789abcdef" * 7 + b"Z") * 300,
             corpus.sparse_with_repeats(60000, 14), synth.text(70000, 9).tobytes() * 2, dna + dna[1000:3000] + dna[:2500], bytes(r.integers(0, 256, 5000, dtype=np.uint8)), b"xyz" * 40]
    done = 0
    for d in cases:
        ok, enc, _, _ = O.transform("ROLZ", d, dst_cap=max(len(d) + 64, 1024) + 1024, ctx=[7, max(len(d), 1024), len(d), 1, 0, 0])
        if ok != 1:
            continue
        done += 1
        for cap in (len(d), len(d) + 100, len(d) - 1):
            o = O.transform("ROLZ", enc, inverse=True, dst_cap=cap, dst_len=cap, ctx=[7, max(len(d), 1024), len(d), 1, 0, 0])
            p = rolz_inverse(enc, cap)
            assert o[0] == (-1 if p[0] is None else int(p[0])), (len(d), cap, o[0], p[0])
            if o[0] == 1:
                assert o[1] == p[1] == d
    assert done >= 7


# ---- stream parsing: K/io/CompressedInputStream.java:359-475 (readHeader, version 7), :1025-1095 (readBlockHeader),
#      :1113-1340 (decodeBlock), Sequence.inverse K/transform/Sequence.java:137-207 -------------------------------------------------
def parse_and_decode_stream(stream, bwt_bounds=1):
    """A .knz walked the way CompressedInputStream does, with the oracle's codecs as black boxes for the payloads: header fields and
    check, per block the length prefix, mode byte, skip flags, length bytes, header check byte, block checksum, entropy stage or copy,
    inverse transforms in reverse order under the skip flags.  -> (decoded bytes, header dict)"""
    names = {v: k for k, v in O.T.items()}
    enames = {v: k for k, v in O.E.items()}
    b = _BitsIn(stream, 8 * len(stream))
    H = 0x1E35A7BD
    assert b.read(32) == 0x4B414E5A
    version = b.read(4)
    assert version == 7
    chk = b.read(2)
    assert chk != 3
    entropy = b.read(5)
    ttype = b.read(48)
    block_size = b.read(28) << 4
    sz_mask = b.read(2)
    out_size = b.read(16 * sz_mask) if sz_mask else 0
    b.read(15)
    crc = b.read(24)
    c = (H * ((0x01030507 * version) & 0xFFFFFFFF)) & 0xFFFFFFFF
    for v in [chk, entropy, ttype >> 32, ttype & 0xFFFFFFFF, block_size] + ([out_size >> 32, out_size & 0xFFFFFFFF] if sz_mask else []):
        c = _mix32(c, H, v)
    assert crc == ((c >> 23) ^ (c >> 3)) & 0xFFFFFF
    slots = [(ttype >> (42 - 6 * i)) & 63 for i in range(8)]
    nbtr = max(sum(1 for t in slots if t != 0), 1)
    fns = [t for i, t in enumerate(slots[:nbtr]) if t != 0 or i == 0]          # TransformFactory.newFunction (:240-264)
    out = bytearray()
    while True:
        lr = b.read(5) + 3
        read = b.read(lr)
        if read == 0:
            break
        # the record's bits as bytes of their own: Java copies them into data.array before it parses them
        recv = (b.v >> (b.total - b.pos - read)) & ((1 << read) - 1)
        b.pos += read
        rb = (recv << ((-read) % 8)).to_bytes((read + 7) // 8, "big")
        r = _BitsIn(rb, read)
        mode = r.read(8)
        copy = (mode & 0x80) != 0
        has_skip, transformed_copy, skip = False, False, 0
        if copy:
            if mode & 0x10:
                transformed_copy = True
                if len(fns) > 4:
                    has_skip = True
                else:
                    skip = ((mode << 4) | 0x0F) & 0xFF
        elif mode & 0x10:
            has_skip = True
        else:
            skip = ((mode << 4) | 0x0F) & 0xFF
        if has_skip:
            skip = r.read(8)
        data_size = 1 + ((mode >> 5) & 3)
        pre = 0
        for _ in range(data_size):
            pre = (pre << 8) | r.read(8)
        check = r.read(8)
        c = (H * 0x01030507) & 0xFFFFFFFF
        for v in (mode, skip, pre, read >> 32, read & 0xFFFFFFFF):
            c = _mix32(c, H, v)
        assert check == ((c >> 23) ^ (c >> 3)) & 0xFF
        assert 0 < pre <= min(max(block_size + block_size // 2, 2048), 1 << 30)
        cks = r.read(32) if chk == 1 else (r.read(64) if chk == 2 else None)
        raw_copy = copy and not transformed_copy
        ent = "NONE" if (raw_copy or transformed_copy) else enames[entropy]
        payload = rb[r.pos // 8:]
        assert r.pos % 8 == 0
        cur, ret, _ = O.entropy_decode(ent, payload, read - r.pos, pre)
        assert ret == pre
        if not raw_copy and skip != 0xFF:
            for i in range(len(fns) - 1, -1, -1):
                if skip & (1 << (7 - i)):
                    continue
                cap = max(block_size, len(cur)) + 1024
                ok, cur, _, _ = O.transform(names[fns[i]], cur, inverse=True, dst_cap=cap, dst_len=cap, src_cap=len(cur) + 16, ctx=[7, block_size, len(cur), 1, 0, bwt_bounds])
                assert ok == 1, (names[fns[i]], i)
        if chk == 1:
            assert cks == O.xxhash32(cur, 0x4B414E5A) & 0xFFFFFFFF
        elif chk == 2:
            assert cks == O.xxhash64(cur, 0x4B414E5A) & M64
        out += cur
    return bytes(out), dict(entropy=entropy, transforms=fns, block_size=block_size, out_size=out_size, checksum=chk)


@pytest.mark.parametrize("transforms,entropy,bs,checksum", [(["LZ"], "HUFFMAN", 65536, 0), (["BWT", "RANK", "ZRLT"], "ANS0", 32768, 32), (["ROLZ"], "NONE", 65536, 64),
                                                            (["NONE"], "FPAQ", 4096, 0), (["LZX"], "ANS1", 1 << 20, 32), (["LZ", "RANK", "ZRLT", "SRT", "MTFT"], "HUFFMAN", 16384, 0),
                                                            (["LZP", "ZRLT"], "ANS0", 65536, 0), (["RLT", "ROLZX"], "RANGE", 131072, 64)])
def test_stream_parsing_agrees_with_the_oracle(transforms, entropy, bs, checksum):
    from kanzi_b200 import synth
    r = np.random.default_rng(5)
    data = synth.pasted(100000, 3).tobytes() + bytes(r.integers(0, 256, 70000, dtype=np.uint8)) + synth.records(40000 + 9, 4).tobytes() + bytes(30000)
    for d in (data, data[:bs + 7], data[:12]):
        knz = O.compress(d, transforms, entropy, bs, checksum=checksum)
        got, hdr = parse_and_decode_stream(knz)
        assert got == d == O.decompress(knz, len(d) + 64)
        assert hdr["block_size"] == bs and hdr["out_size"] == len(d) and hdr["checksum"] == {0: 0, 32: 1, 64: 2}[checksum]
        assert hdr["transforms"] == [O.T[t] for t in transforms if t != "NONE"] or transforms == ["NONE"]


# ---- BWTS by definition (K/transform/BWTS.java:60-160; Gil & Scott's bijective BWT): factor the input into Lyndon words, sort every rotation of
#      every factor by its infinite repetition, emit the byte before each rotation's start.  Not on the CUDA path; the oracle carries it.
def bwts_by_definition(d):
    n = len(d)
    factors = []
    i = 0
    while i < n:                                                   # Duval
        j, k = i + 1, i
        while j < n and d[k] <= d[j]:
            k = i if d[k] < d[j] else k + 1
            j += 1
        while i <= k:
            factors.append(d[i:i + j - k])
            i += j - k
    rots = []
    for w in factors:
        for r in range(len(w)):
            rots.append((w[r:] + w[:r], w[r - 1]))
    from functools import cmp_to_key

    def cmp(a, b):
        u, v = a[0], b[0]
        m = len(u) + len(v)
        x, y = (u * (m // len(u) + 1))[:m], (v * (m // len(v) + 1))[:m]
        return (x > y) - (x < y)

    return bytes(last for _, last in sorted(rots, key=cmp_to_key(cmp)))


def test_bwts_matches_its_definition():
    from kanzi_b200 import synth
    r = np.random.default_rng(53)
    cases = [b"banana", b"mississippi", b"abracadabra", b"aaaaaa", b"ab", b"ba", b"cba", b"abcabcabc", b"zyxzyxzy", synth.text(1500, 2).tobytes(),
             bytes(r.integers(0, 3, 900, dtype=np.uint8)), bytes(r.integers(0, 256, 1200, dtype=np.uint8)), bytes(100), b"ab" * 300 + b"a"]
    for d in cases:
        ok, out, _, _ = O.transform("BWTS", d)
        assert ok == 1 and out == bwts_by_definition(d), (len(d), out[:20])
        assert O.transform("BWTS", out, inverse=True, dst_cap=len(d), dst_len=len(d))[:2] == (1, d)
    assert O.transform("BWTS", b"banana")[1] == b"annbaa"             # the textbook example
