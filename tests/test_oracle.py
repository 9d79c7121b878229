"""CPU tests of the oracle (the C++ restatement under oracle/): known answers the reference itself holds
 and the round-trip cases of the reference's own tests."""
import numpy as np
import pytest
import oracle_lib as O
import corpus

# ---- known answers ---------------------------------------------------------------------------------------------------
HEADER_KATS = [  # SURVEY.md Appendix D (stream headers derived from COS:236-313)
    (["NONE"], "HUFFMAN", 65536, 1048576, "4b414e5a70200000000000000020010008000000006b6263"),
    (["LZ"], "ANS0", 4 << 20, 211957760, "4b414e5a70a180000000000008000106511c0000009b5920"),
    (["BWT", "RANK", "ZRLT"], "ANS1", 8 << 20, 10 ** 8, "4b414e5a710090300000000010000102faf0800000b3e871"),
    (["BWT", "SRT", "ZRLT"], "FPAQ", 32 << 20, 10 ** 9, "4b414e5a70409a30000000004000011dcd65000000e05ded"),
    (["ROLZ"], "ANS0", 16 << 20, 8 << 30, "4b414e5a70a58000000000002000018001000000000000489998"),
]


@pytest.mark.parametrize("tr,ent,bs,size,hexs", HEADER_KATS)
def test_stream_header_kat(tr, ent, bs, size, hexs):
    assert O.stream_header(tr, ent, bs, size).hex() == hexs


def test_whole_stream_kat():
    # the stream T/test/TestCompressedStream.java:178-187 builds: 8 bytes, NONE&NONE, block 1024, size unknown
    s = O.compress(bytes([1, 2, 3, 4, 5, 6, 7, 8]), ["NONE"], "NONE", 1024, input_size=0)
    assert s.hex() == "4b414e5a700000000000000000008000000ea77b2588708410102030405060708000"


def test_block_header_checksum_formula():
    # T/test/TestCompressedStream.java:488-504 restates mix32 + the block header checksum; same arithmetic here
    def mix32(c, h, v):
        c ^= (h * (~v & 0xFFFFFFFF)) & 0xFFFFFFFF
        c = ((c << 13) | (c >> 19)) & 0xFFFFFFFF
        return (c * 5 + 0x52DCE729) & 0xFFFFFFFF
    H = 0x1E35A7BD
    c = (H * 0x01030507) & 0xFFFFFFFF
    for v in (0x87, 0, 8, 0, 88):
        c = mix32(c, H, v)
    assert ((c >> 23) ^ (c >> 3)) & 0xFF == 0x41       # byte 24 of the KAT stream above


def test_bwt_mississippi_kat():   # K/transform/BWT.java:45-50
    ok, out, pi = O.bwt_forward(b"mississippi")
    assert ok == 1 and out == b"ipssmpissii" and pi[0] == 5


def test_expgolomb_kat():   # ExpGolombEncoder.java:53,69-70: +1 -> 0100, -1 -> 0101
    import ctypes as C
    bits = C.c_uint32(0)
    assert O.lib().kzo_expgolomb_signed(1, C.byref(bits)) == 4 and bits.value == 0b0100
    assert O.lib().kzo_expgolomb_signed(-1, C.byref(bits)) == 4 and bits.value == 0b0101
    assert O.lib().kzo_expgolomb_signed(0, C.byref(bits)) == 1 and bits.value == 1
    assert O.lib().kzo_expgolomb_signed(-128, C.byref(bits)) == 16 and bits.value == 259


# ---- round trips (the reference's test strategy, §4) -------------------------------------------------------------------
CASES = corpus.small_cases()


@pytest.mark.parametrize("ent", ["NONE", "HUFFMAN", "ANS0", "ANS1", "FPAQ", "RANGE"])
def test_entropy_roundtrip(ent):
    for name, d in list(CASES.items()) + [(f"lit{i}", x) for i, x in enumerate(corpus.ENTROPY_LITERALS)] + [("fib", corpus.fibonacci_chunk())]:
        pay, bits = O.entropy_encode(ent, d)
        out, r, used = O.entropy_decode(ent, pay, bits, len(d))
        assert r == len(d) and out == d and used == bits, (ent, name)


@pytest.mark.parametrize("tr", ["LZ", "LZX", "ROLZ", "ZRLT", "RANK", "MTFT", "SRT", "BWT", "LZP", "RLT", "ROLZX", "BWTS"])
def test_transform_roundtrip(tr):
    applied = 0
    for name, d in CASES.items():
        ctx = [7, max(len(d), 1024), len(d), 1, 0, 0]
        ok, out, used, _ = O.transform(tr, d, dst_cap=len(d) + len(d) // 64 + 1100, ctx=ctx)
        assert ok in (0, 1), (tr, name)
        if ok != 1:
            continue
        applied += 1
        ok2, back, _, _ = O.transform(tr, out, inverse=True, dst_cap=len(d) + 512, src_cap=len(out) + 16, ctx=ctx)
        assert ok2 == 1 and back == d, (tr, name)
    assert applied > 5


def test_bwt_raw_both_inverses():
    for d in corpus.BWT_LITERALS + [CASES["text64k"], CASES["runs"], CASES["zeros80k"][:3000], bytes(np.arange(70001, dtype=np.uint32).astype(np.uint8))]:
        ok, b, pi = O.bwt_forward(d)
        assert ok == 1
        assert O.bwt_inverse(b, pi, 0) == (1, d)
        if len(d) >= 2:
            assert O.bwt_inverse(b, pi, 1) == (1, d)
        if len(d) >= 256:
            assert O.bwt_inverse(b, pi, 2) == (1, d)


def test_bwt_blockcodec_as_reference_never_fires():
    # SURVEY.md E-1 / DESIGN.md: as written, BWTBlockCodec.forward returns false for the stream's slices
    d = CASES["text64k"]
    ok, _, _, _ = O.transform("BWT", d, dst_cap=len(d) + 33, ctx=[7, 65536, len(d), 1, 0, 1])
    assert ok == 0
    ok, out, _, _ = O.transform("BWT", d, dst_cap=len(d) + 33, ctx=[7, 65536, len(d), 1, 0, 0])
    assert ok == 1 and len(out) == len(d) + 1 + 8 * 2


STREAM_CFGS = [(["NONE"], "HUFFMAN", 65536), (["LZ"], "ANS0", 1 << 20), (["BWT", "RANK", "ZRLT"], "ANS1", 1 << 20),
               (["BWT", "SRT", "ZRLT"], "FPAQ", 1 << 20), (["ROLZ"], "ANS0", 1 << 20), (["LZX"], "HUFFMAN", 1 << 18), (["NONE"], "NONE", 1 << 16)]


@pytest.mark.parametrize("tr,ent,bs", STREAM_CFGS)
@pytest.mark.parametrize("bwt_bounds", [1, 0])
def test_stream_roundtrip(tr, ent, bs, bwt_bounds):
    if bwt_bounds == 0 and "BWT" not in tr:
        pytest.skip("bounds switch only matters for BWT chains")
    from kanzi_b200 import synth
    d = (synth.text(1_300_000, 3).tobytes() + synth.noise(200_000, 4).tobytes() + bytes(70000) + synth.exe_like(500_007, 5).tobytes() + b"tail!")
    s = O.compress(d, tr, ent, bs, bwt_bounds=bwt_bounds)
    assert O.decompress(s, len(d) + 1024, bwt_bounds=bwt_bounds) == d


def test_stream_tiny_inputs():
    for n in (0, 1, 8, 15, 16, 17, 100):
        d = bytes(range(n))
        s = O.compress(d, ["LZ"], "ANS0", 1024)
        assert O.decompress(s, 2048) == d


def test_mt_block_encoder_matches_stream():
    from kanzi_b200 import synth
    d = synth.silesia_like(3_000_001, 2)
    bs = 1 << 20
    s = O.compress(d, ["LZ"], "ANS0", bs)
    recs, off, bits = O.encode_blocks_mt(d, ["LZ"], "ANS0", bs, 4)
    # re-assemble the stream from the per-block records and compare with the single-threaded emulator
    hdr = O.stream_header(["LZ"], "ANS0", bs, len(d))
    acc = int.from_bytes(hdr, "big")
    nbits = len(hdr) * 8
    for o, b in zip(off, bits):
        nb = (int(b) + 7) // 8
        v = int.from_bytes(recs[int(o): int(o) + nb].tobytes(), "big") >> (nb * 8 - int(b))
        acc = (acc << int(b)) | v
        nbits += int(b)
    acc <<= 8
    nbits += 8
    pad = (-nbits) % 8
    assert (acc << pad).to_bytes((nbits + pad) // 8, "big") == s
    back = O.decode_blocks_mt(recs, off, bits, ["LZ"], "ANS0", bs, 4, len(d))
    assert back.tobytes() == d.tobytes()


def test_xxhash_published_vectors():
    """XXHash32 is the published XXH32 (K/util/hash/XXHash32.java:94-142): the reference vectors of the xxHash project pin it.
    XXHash64 matches the published XXH64 below 32 bytes; from 32 bytes on Kanzi folds its accumulators with a 32-bit rotation
    idiom on 64-bit values (XXHash64.java:127-128), so its result is Kanzi's own and only the short vectors apply."""
    assert O.xxhash32(b"", 0) == 0x02CC5D05
    assert O.xxhash32(b"abc", 0) == 0x32D153FF
    assert O.xxhash32(b"Nobody inspects the spammish repetition", 0) == 0xE2293B2F
    assert O.xxhash64(b"", 0) == 0xEF46DB3751D8E999
    assert O.xxhash64(b"abc", 0) == 0x44BC2CF5AD770999
    assert O.xxhash64(b"Nobody inspects the spammish repetition", 0) != 0xFBCEA83C8A378BF1      # the published XXH64: not what Kanzi computes


def test_checksummed_streams_round_trip_and_detect_corruption():
    d = bytes((i * 131 + (i >> 7)) & 0xFF for i in range(200_000))
    for ck in (32, 64):
        knz = O.compress(d, ["LZ"], "ANS0", 1 << 16, checksum=ck)
        plain = O.compress(d, ["LZ"], "ANS0", 1 << 16)
        assert len(knz) == len(plain) + (ck // 8) * 4            # 4 blocks, ck/8 bytes each (the lw field keeps its width here)
        assert O.decompress(knz, len(d) + 64) == d
        bad = bytearray(knz)
        bad[len(bad) // 2] ^= 0x04
        with pytest.raises(RuntimeError):
            O.decompress(bytes(bad), len(d) + 64)


def test_none_tokens_are_dropped_from_a_transform_list():
    # TransformFactory.getType (K/transform/TransformFactory.java:140-153): "-t BWT+NONE+RANK+ZRLT" is the type of "BWT+RANK+ZRLT"
    d = corpus.small_cases()["text64k"]
    assert O.compress(d, ["BWT", "NONE", "RANK", "ZRLT"], "ANS1", 32768, bwt_bounds=0) == O.compress(d, ["BWT", "RANK", "ZRLT"], "ANS1", 32768, bwt_bounds=0)
    assert O.compress(d, ["NONE", "LZ"], "ANS0", 32768) == O.compress(d, ["LZ"], "ANS0", 32768)
    assert O.stream_header(["NONE", "NONE"], "HUFFMAN", 65536, 100) == O.stream_header(["NONE"], "HUFFMAN", 65536, 100)
