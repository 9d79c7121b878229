"""Full-size properties (BASELINE.json configs at their real sizes): what the small bit-exact cases cannot show.
At 200 MB the oracle (one CPU thread) checks a prefix; the whole stream is checked through size-independent properties:
encode -> decode round trip, and block independence (COS:905-907: a block's record depends on nothing but its own bytes,
so the records of a prefix of the input are a prefix of the records of the whole input)."""
import numpy as np
import pytest

import kanzi_b200 as K
import oracle_lib as O
from kanzi_b200 import synth

pytestmark = pytest.mark.gpu


def _first_diff(a, b):
    n = min(len(a), len(b))
    x = np.frombuffer(a[:n], dtype=np.uint8) != np.frombuffer(b[:n], dtype=np.uint8)
    return int(np.argmax(x)) if x.any() else n


@pytest.mark.timeout(600)
def test_cfg2_full_size_round_trip_and_block_independence():
    gen, size, tr, ent, bs = synth.CONFIGS["cfg2"]
    data = gen(size, 2).tobytes()
    knz = K.compress(data, tr, ent, bs)
    assert K.decompress(knz, len(data) + 1024) == data
    # the oracle on the first 6 blocks: byte-identical records
    k = 6 * bs
    ref = O.compress(data[:k], tr, ent, bs)
    got = K.compress(data[:k], tr, ent, bs)
    assert got == ref, ("first differing byte", _first_diff(got, ref))
    # block independence: the prefix stream minus its end-of-stream marker opens the full stream
    # (the 24-byte stream header carries the input size and a checksum over it, COS:236-313; both sizes need 32 bits, so the
    # records start at the same bit in both streams; the prefix stream ends with the 5-bit length-of-length + empty block)
    hdr, body = 24, len(got) - 2
    assert knz[hdr:body] == got[hdr:body], ("first differing byte", hdr + _first_diff(knz[hdr:body], got[hdr:body]))
    # a different split of the same bytes into batches gives the same records: last 9 blocks alone vs inside the whole
    tail = data[(size // bs - 8) * bs:]
    t_knz = K.compress(tail, tr, ent, bs)
    assert K.decompress(t_knz, len(tail) + 1024) == tail


@pytest.mark.timeout(600)
def test_cfg1_full_size_bit_exact():
    gen, size, tr, ent, bs = synth.CONFIGS["cfg1"]
    data = gen(size, 1).tobytes()
    ref = O.compress(data, tr, ent, bs)
    got = K.compress(data, tr, ent, bs)
    assert got == ref, ("first differing byte", _first_diff(got, ref))
    assert K.decompress(ref, len(data) + 1024) == data


@pytest.mark.timeout(900)
def test_cfg3_full_block_size_round_trip():
    """cfg3's chain at its real block size (8 MiB), three blocks: round trip on the GPU, first block's record against the oracle."""
    gen, size, tr, ent, bs = synth.CONFIGS["cfg3"]
    data = gen(3 * bs - 12345, 3).tobytes()
    for flags in (K.FLAG_BWT_ASREF, 0):
        knz = K.compress(data, tr, ent, bs, flags=flags)
        assert K.decompress(knz, len(data) + 1024, flags=flags) == data
        ref = O.compress(data[:bs], tr, ent, bs, bwt_bounds=1 if flags else 0)
        got = K.compress(data[:bs], tr, ent, bs, flags=flags)
        assert got == ref, (flags, "first differing byte", _first_diff(got, ref))


@pytest.mark.timeout(1800)
def test_cfg4_real_block_size_two_blocks():
    """cfg4's chain (BWT+SRT+ZRLT & FPAQ) at its real block size: two blocks of 32 MiB (the n > 8 MiB regime: biPSIv2 on the
    reference side, K/transform/BWT.java:384-544; 8 FPAQ chunks of 4 MiB sharing state, K/entropy/FPAQEncoder.java:140-170),
    both readings of the BWT bounds clause.  First block's record against the oracle, both blocks round trip."""
    gen, size, tr, ent, bs = synth.CONFIGS["cfg4"]
    assert bs == 32 << 20
    data = gen(2 * bs - 54321, 4).tobytes()
    for flags in (K.FLAG_BWT_ASREF, 0):
        knz = K.compress(data, tr, ent, bs, flags=flags)
        assert K.decompress(knz, len(data), flags=flags) == data
        ref = O.compress(data[:bs], tr, ent, bs, bwt_bounds=1 if flags else 0)
        got = K.compress(data[:bs], tr, ent, bs, flags=flags)
        assert got == ref, (flags, "first differing byte", _first_diff(got, ref))
        hdr, body = 24, len(got) - 2          # block independence: the one-block stream's record opens the two-block stream
        assert knz[hdr:body] == got[hdr:body]


@pytest.mark.timeout(1800)
def test_cfg5_real_block_size_four_blocks():
    """cfg5's chain (ROLZ & ANS0) at its real block size: four blocks of 16 MiB (one full ROLZ chunk each, ROLZCodec.java:497-509),
    whole stream against the oracle, and the decode of the oracle's stream."""
    gen, size, tr, ent, bs = synth.CONFIGS["cfg5"]
    assert bs == 16 << 20
    data = gen(4 * bs - 777, 5).tobytes()
    ref = O.compress(data, tr, ent, bs)
    got = K.compress(data, tr, ent, bs)
    assert got == ref, ("first differing byte", _first_diff(got, ref))
    assert K.decompress(ref, len(data)) == data
