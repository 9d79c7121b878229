"""Block sharding across the GPUs of one box.

Blocks of a `.knz` stream are independent (`EncodingTask` rebuilds every codec per block,
K/io/CompressedOutputStream.java:905-907), so rank r of N owns blocks r, r+N, r+2N, ... (1-based block id b
-> GPU (b-1) % N, as the survey's partitioning rule says) and no block data crosses ranks.  The only exchange
is the per-block encoded size: every rank needs all of them to know at which bit of the joint stream its
records start (K/io/CompressedOutputStream.java:1024-1035 concatenates records bit-granularly in block order).
That is one `all_gather` of int64 — NCCL on GPUs, gloo in the CPU tests.
"""
import numpy as np


def blocks_of_rank(n_blocks, world, rank):
    """0-based ids of the blocks rank `rank` of `world` encodes/decodes."""
    return list(range(rank, n_blocks, world))


def owner_of_block(block, world):
    return block % world


def gather_block_bits(local_bits, n_blocks, world, rank, device="cpu", group=None):
    """all_gather of the encoded bit lengths.  `local_bits[i]` belongs to block blocks_of_rank(...)[i].
    Returns a list of n_blocks ints in block order (identical on every rank)."""
    import torch
    import torch.distributed as dist
    per = (n_blocks + world - 1) // world
    mine = torch.zeros(per, dtype=torch.int64, device=device)
    if len(local_bits):
        mine[: len(local_bits)] = torch.as_tensor(list(local_bits), dtype=torch.int64, device=device)
    if world == 1:
        parts = [mine]
    else:
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
    out = [0] * n_blocks
    for r in range(world):
        ids = blocks_of_rank(n_blocks, world, r)
        vals = parts[r].tolist()
        for i, b in enumerate(ids):
            out[b] = int(vals[i])
    return out


def stream_bit_offsets(header_bits, block_bits):
    """Bit offset of every block record in the joint stream, and the offset of the end marker."""
    offs, acc = [], int(header_bits)
    for b in block_bits:
        offs.append(acc)
        acc += int(b)
    return offs, acc


def place_bits(stream, bit_off, rec, nbits):
    """OR the first `nbits` bits of `rec` (MSB-first) into `stream` (uint8 array, zero where written) at bit `bit_off`."""
    if nbits <= 0:
        return
    nb = (nbits + 7) // 8
    a = np.zeros(nb + 1, dtype=np.uint8)
    a[:nb] = np.frombuffer(bytes(rec[:nb]), dtype=np.uint8)
    if nbits & 7:
        a[nb - 1] &= (0xFF << (8 - (nbits & 7))) & 0xFF
    sh = bit_off & 7
    if sh:
        w = a.astype(np.uint16)
        a = ((w >> sh) | (np.concatenate(([0], w[:-1])) << (8 - sh))).astype(np.uint8)
    first = bit_off >> 3
    total = (sh + nbits + 7) // 8
    stream[first: first + total] |= a[:total]
