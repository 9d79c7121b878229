"""The per-block path the Java host really drives (one ByteTransform.forward + EntropyEncoder.encode per block and pool thread,
K/io/CompressedOutputStream.java:792-916), from C with 16 pthreads (tests/native/kzg_blocks_mt.c), with the library's call
coalescing off and on: every block round-trips, and coalesced calls produce the same bits as lone ones."""
import json
import os
import subprocess
import pytest
import numpy as np
import kanzi_b200 as K
import oracle_lib as O
from kanzi_b200 import build as kb

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("xf,ent,bs", [(3, 5, 1 << 20), (11, 5, 1 << 18), (8, 1, 1 << 16)])
def test_per_block_threads_round_trip(xf, ent, bs):
    exe = kb.build_native()
    p = subprocess.run([exe, "16", "48", str(bs), str(xf), str(ent), "64", "300"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, (p.stdout, p.stderr)
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["failures"] == 0 and line["coalesced_bits_equal"]
    assert line["coalesced"]["requests"] >= 2 * 48 and line["coalesced"]["batches"] < line["coalesced"]["requests"]


def test_coalesced_calls_match_the_oracle():
    """Python threads through ctypes (the GIL is released inside the library): coalesced per-block results == oracle == lone calls."""
    import threading
    from kanzi_b200 import synth
    blocks = [synth.text(200_000 + 1000 * i, 70 + i).tobytes() for i in range(12)]
    ref = [O.transform("LZ", b)[1] for b in blocks]
    L = K.lib()
    L.kzg_set_coalescing.argtypes = [__import__("ctypes").c_int, __import__("ctypes").c_int]
    L.kzg_set_coalescing(16, 500)
    got = [None] * len(blocks)

    def work(i):
        ok, out, used = K.transform_forward("LZ", blocks[i], {"blockSize": len(blocks[i]), "size": len(blocks[i]), "flags": 0})
        got[i] = out if ok else None
    try:
        th = [threading.Thread(target=work, args=(i,)) for i in range(len(blocks))]
        [t.start() for t in th]
        [t.join() for t in th]
    finally:
        L.kzg_set_coalescing(0, 0)
    assert got == ref
