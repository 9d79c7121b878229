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


def extract_bits(stream, bit_off, nbits):
    """The `nbits` bits of `stream` (uint8 array) starting at bit `bit_off`, left-aligned in a fresh uint8 array (tail bits zero)."""
    nb = (nbits + 7) // 8
    first, sh = bit_off >> 3, bit_off & 7
    a = np.zeros(nb + 1, dtype=np.uint16)
    src = np.asarray(stream[first: first + nb + 1], dtype=np.uint16)
    a[: len(src)] = src
    out = (((a[:-1] << sh) | (a[1:] >> (8 - sh))) & 0xFF).astype(np.uint8) if sh else a[:-1].astype(np.uint8)
    if nbits & 7:
        out[nb - 1] &= (0xFF << (8 - (nbits & 7))) & 0xFF
    return out


def shard_of_rank(data, block_size, world, rank):
    """The bytes rank `rank` encodes: its blocks (b = rank, rank + world, ...) of `data`, concatenated in block order.  Every
    block but the stream's last is full, so the concatenation splits back into the same blocks (the stream's last block, if
    short, is the last block of its owner's shard)."""
    n = len(data)
    nb = (n + block_size - 1) // block_size
    parts = [data[b * block_size: min(n, (b + 1) * block_size)] for b in blocks_of_rank(nb, world, rank)]
    return np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)


def assemble_joint_stream(header, shard_streams, shard_index, n_blocks, world):
    """The joint .knz from every rank's own stream: `shard_streams[r]` is rank r's .knz (uint8 array), `shard_index[r]` =
    (bit offsets, bit lengths) of its records.  Records are placed in block order after `header` (bytes), then the end marker
    (5 + 3 zero bits, CompressedOutputStream.java:491-492).  Test / verification helper: the product never moves payloads."""
    all_bits = [0] * n_blocks
    for r in range(world):
        for i, b in enumerate(blocks_of_rank(n_blocks, world, r)):
            all_bits[b] = int(shard_index[r][1][i])
    offs, end = stream_bit_offsets(len(header) * 8, all_bits)
    out = np.zeros((end + 8 + 7) // 8, dtype=np.uint8)
    out[: len(header)] = np.frombuffer(bytes(header), dtype=np.uint8)
    for r in range(world):
        for i, b in enumerate(blocks_of_rank(n_blocks, world, r)):
            rec = extract_bits(shard_streams[r], int(shard_index[r][0][i]), all_bits[b])
            place_bits(out, offs[b], rec, all_bits[b])
    return out
