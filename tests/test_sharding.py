"""N > 1 host logic on CPU: two gloo ranks each encode their share of the blocks (with the oracle standing in for the
device path), all_gather the per-block sizes, and the joint stream assembled from the gathered offsets must be the
single-process .knz."""
import os
import sys
import socket
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from kanzi_b200 import sharding


def test_block_partition_is_a_partition():
    for nb in (1, 2, 7, 51, 512):
        for world in (1, 2, 4, 8):
            seen = sorted(b for r in range(world) for b in sharding.blocks_of_rank(nb, world, r))
            assert seen == list(range(nb))
            assert all(sharding.owner_of_block(b, world) == r for r in range(world) for b in sharding.blocks_of_rank(nb, world, r))


def test_place_bits_matches_big_int_concat():
    rng = np.random.default_rng(5)
    recs = [rng.integers(0, 256, size=int(rng.integers(1, 40)), dtype=np.uint8) for _ in range(30)]
    bits = [int(len(r) * 8 - rng.integers(0, 8)) for r in recs]
    offs, end = sharding.stream_bit_offsets(13, bits)
    out = np.zeros((end + 7) // 8 + 2, dtype=np.uint8)
    acc = 0
    for r, b, o in zip(recs, bits, offs):
        sharding.place_bits(out, o, r, b)
        acc = (acc << b) | (int.from_bytes(r.tobytes(), "big") >> (len(r) * 8 - b))
    pad = (-end) % 8
    assert out[: (end + pad) // 8].tobytes() == (acc << pad).to_bytes((end + pad) // 8, "big")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import oracle_lib as O
    from kanzi_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bs = 1 << 18
        data = synth.silesia_like(7 * bs + 12345, 2)
        nb = (len(data) + bs - 1) // bs
        mine = sharding.blocks_of_rank(nb, world, rank)
        # every rank encodes only its own blocks (the oracle's block encoder stands in for the device path on CPU)
        recs, bits = [], []
        for b in mine:
            chunk = data[b * bs: min(len(data), (b + 1) * bs)]
            out, off, nbits = O.encode_blocks_mt(chunk, ["LZ"], "ANS0", bs, 1)
            recs.append(out[: (int(nbits[0]) + 7) // 8].copy())
            bits.append(int(nbits[0]))
        all_bits = sharding.gather_block_bits(bits, nb, world, rank)
        hdr = O.stream_header(["LZ"], "ANS0", bs, len(data))
        offs, end = sharding.stream_bit_offsets(len(hdr) * 8, all_bits)
        # the test gathers the payloads too, to check the offsets; the product leaves them where they are
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine, recs))
        if rank == 0:
            stream = np.zeros((end + 8 + 7) // 8, dtype=np.uint8)
            stream[: len(hdr)] = np.frombuffer(hdr, dtype=np.uint8)
            for ids, rs in gathered:
                for b, r in zip(ids, rs):
                    sharding.place_bits(stream, offs[b], r, all_bits[b])
            ref = O.compress(data, ["LZ"], "ANS0", bs)        # ends with the 8-bit zero-length marker
            q.put(("ok", stream.tobytes() == ref, len(ref)))
    except Exception as e:                                     # pragma: no cover
        q.put(("err", repr(e), 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_two_ranks_gloo_joint_stream():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    kind, ok, n = q.get(timeout=200)
    for p in ps:
        p.join(timeout=60)
    assert kind == "ok", ok
    assert ok and n > 0
