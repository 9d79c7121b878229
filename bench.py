#!/usr/bin/env python
"""bench.py — the hot path of BASELINE.json on synthetic blocks: encode + decode MB/s, bit-exact .knz.

One step = one pass of the hot path over one batch: every block of the workload is encoded
(transform chain + entropy coder + container assembly -> .knz) and the .knz is decoded back.
  value : whole-job MB/s (10^6 original bytes / step time) with inputs already resident in HBM
  e2e   : the same through the host-buffer C ABI (pinned H2D of the input / D2H of the .knz inside the
          timed region on the way in, H2D of the .knz / D2H of the decoded bytes on the way back)
  roofline      : the top kernel (by CUDA-event time, measured live with the library's per-kernel events) vs the
                  measured HBM copy bandwidth (MEASURED_PEAKS.json), next to the dominant stage
  cpu_baseline  : the C++ restatement of the reference (oracle/) on the box's host cores, bounded sample
`--impl reference` times that CPU path as the reference arm (no JVM exists in the image: SURVEY.md §0.2).

Multi-GPU (N > 1, torchrun, one rank per GPU): ONE joint stream of 51*N blocks (N silesia-shaped segments) is dealt
round-robin, block b -> rank b % N (kanzi_b200/sharding.py).  Every rank encodes and decodes only its own blocks; the one
exchange inside the timed step is the NCCL all-gather of the per-block record bit lengths, from which every rank computes
every record's bit offset in the joint stream (CompressedOutputStream.java:1024-1035).  Outside the timed region rank 0
encodes the whole joint stream on its own GPU and checks every rank's records (length + CRC) against it.  Per-GPU work is
fixed as N grows ("scaling": "weak"); the extra key `strong_scaling` times a fixed 400-block stream (8 segments of 50 whole
blocks) dealt over the N ranks."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
# the LZ forward runs block groups on up to 32 side streams: more hardware work queues than the default 8 (read at CUDA context creation)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--scale", type=float, default=None, help="fraction of the config's full size (default: full, capped for cfg4/cfg5)")
    ap.add_argument("--cpu-sample-mb", type=float, default=256.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the short lines for the other BASELINE configs (N=1 default run only)")
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed-total (400 blocks) strong-scaling measurement")
    ap.add_argument("--bwt-fixed", action="store_true", help="BWT bounds 'fixed' instead of as the reference is written")
    return ap.parse_args()


DEFAULT_SCALE = {"cfg1": 1.0, "cfg2": 1.0, "cfg3": 1.0, "cfg4": 0.1, "cfg5": 1.0 / 8}
SEEDS = {"cfg1": 1, "cfg2": 2, "cfg3": 3, "cfg4": 4, "cfg5": 5}
STRONG_SEGMENTS = 8          # 8 segments x 50 whole blocks of cfg2 = 400 blocks


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc, self.marks = gpu, [], None, []

    def mark(self):
        self.marks.append(len(self.rows))

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        # samples inside the resident timed region when it was long enough to hold any; otherwise everything sampled while the
        # benchmark was running (warm-up, timed steps, e2e steps: the same workload throughout)
        region = self.rows[self.marks[0]:self.marks[1]] if len(self.marks) >= 2 else []
        scope = "timed region" if region else "warm-up + timed + e2e steps"
        self.rows = region or self.rows
        self.scope = scope
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(self.rows), "scope": self.scope}


def config_dict(cfg, transforms, entropy, bs, n, gen_name, flags, world):
    """The SAME dict in both arms (the driver compares them)."""
    return {"workload": f"{cfg}: {'+'.join(transforms)}&{entropy} -b {bs}, {n} bytes ({(n + bs - 1) // bs} blocks) per GPU, synthetic {gen_name}",
            "sharding": "single GPU" if world == 1 else f"one joint stream of {world} segments, block b -> rank b % {world}, per-block bit lengths all-gathered",
            "l2": "flushed between steps (256 MiB write); inputs also exceed L2", "bwt_bounds": "asref" if flags else "fixed",
            "resident_decode_note": "the resident decode gets the .knz in HBM plus an untimed host copy of it for the container walk (record lengths are a serial chain the host walks)"}


def cpu_arm(data, transforms, entropy, bs, sample_mb, flags, steps=1, warmup=0):
    """The reference's CPU path for this chain = the oracle port, one EncodingTask/DecodingTask per host thread."""
    import oracle_lib as O
    cores = os.cpu_count() or 1
    nblk = max(1, int(sample_mb * 1e6) // bs)
    sample = np.ascontiguousarray(data[: min(len(data), nblk * bs)])
    enc_t, dec_t = [], []
    # buffers allocated (and touched) outside the timed region, as the GPU arm's are
    obuf = np.ones(len(sample) + len(sample) // 4 + (1 << 16) + 1200 * (len(sample) // bs + 1), dtype=np.uint8)
    bbuf = np.ones(len(sample) + 64, dtype=np.uint8)
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        recs, off, bits = O.encode_blocks_mt(sample, transforms, entropy, bs, cores, bwt_bounds=1 if flags else 0, out=obuf)
        t1 = time.perf_counter()
        back = O.decode_blocks_mt(recs, off, bits, transforms, entropy, bs, cores, len(sample), bwt_bounds=1 if flags else 0, out=bbuf)
        t2 = time.perf_counter()
        if it >= warmup:
            enc_t.append(t1 - t0)
            dec_t.append(t2 - t1)
    assert back.tobytes() == sample.tobytes(), "oracle round trip failed"
    te, td = sum(enc_t) / len(enc_t), sum(dec_t) / len(dec_t)
    mb = len(sample) / 1e6
    return {"value": mb / (te + td), "unit": "MB/s", "cores": cores, "kind": "port",
            "sample": f"first {len(sample)} bytes ({(len(sample) + bs - 1) // bs} blocks) of the workload, {cores} threads, one block per thread",
            "encode_MBps": mb / te, "decode_MBps": mb / td, "ms_per_step": 1e3 * (te + td)}


class Codec:
    """Device-resident and host-buffer entry points of libkanzi_b200 over one (shard of a) workload."""

    def __init__(self, torch, K, dev, data, transforms, entropy, bs, flags):
        self.torch, self.K, self.L, self.dev = torch, K, K.lib(), dev
        self.data, self.n, self.bs, self.flags = data, len(data), bs, flags
        n = self.n
        self.h_in = torch.from_numpy(data).pin_memory()
        self.d_in = torch.zeros(n + 256, dtype=torch.uint8, device=dev)
        self.d_in[:n].copy_(self.h_in)
        self.cap = int(K.compress_bound(n, bs))
        self.d_knz = torch.zeros(self.cap + 256, dtype=torch.uint8, device=dev)
        self.d_back = torch.zeros(n + 256, dtype=torch.uint8, device=dev)
        self.h_knz = torch.zeros(self.cap, dtype=torch.uint8).pin_memory()
        self.h_back = torch.zeros(n + 256, dtype=torch.uint8).pin_memory()
        self.ids = (C.c_int32 * 8)(*([K.T[t] for t in transforms] + [0] * (8 - len(transforms))))
        self.nT, self.eid = len(transforms), K.E[entropy]
        self.tim_e, self.tim_d = (C.c_float * 3)(), (C.c_float * 3)()
        self.u8p = C.POINTER(C.c_uint8)
        torch.cuda.synchronize()

    def enc_dev(self):
        r = self.L.kzg_compress_dev(self.d_in.data_ptr(), self.n, self.ids, self.nT, self.eid, self.bs, self.flags, self.d_knz.data_ptr(), self.cap, self.tim_e)
        if r < 0:
            raise self.K.KzgError(r, "kzg_compress_dev")
        return r

    def dec_dev(self, knz_len):
        r = self.L.kzg_decompress_dev(self.d_knz.data_ptr(), knz_len, C.cast(self.h_knz.data_ptr(), self.u8p), self.flags, self.d_back.data_ptr(), self.n, self.tim_d)
        if r < 0:
            raise self.K.KzgError(r, "kzg_decompress_dev")
        return r

    def enc_host(self):
        k = self.L.kzg_compress(C.cast(self.h_in.data_ptr(), self.u8p), self.n, self.ids, self.nT, self.eid, self.bs, self.flags, C.cast(self.h_knz.data_ptr(), self.u8p), self.cap)
        if k < 0:
            raise self.K.KzgError(k, "kzg_compress")
        return k

    def dec_host(self, k):
        r = self.L.kzg_decompress(C.cast(self.h_knz.data_ptr(), self.u8p), k, self.flags, C.cast(self.h_back.data_ptr(), self.u8p), self.n)
        if r != self.n:
            raise self.K.KzgError(r, "kzg_decompress")
        return r

    def gate(self):
        """device round trip; leaves the .knz on the host (h_knz) too.  -> knz_len"""
        torch = self.torch
        knz_len = self.enc_dev()
        self.h_knz[:knz_len].copy_(self.d_knz[:knz_len])
        torch.cuda.synchronize()
        assert self.dec_dev(knz_len) == self.n
        assert torch.equal(self.d_back[:self.n], self.d_in[:self.n]), "GPU round trip mismatch"
        return knz_len


def joint_segments(base, segments):
    """The joint input of an N-rank run: N silesia-shaped segments = the base workload rotated by a segment-specific odd amount
    (every block of every segment differs from every other, same statistics; generating N independent 212 MB inputs would
    cost minutes of numpy time per rank)."""
    return [base if s == 0 else np.roll(base, s * 1000003) for s in range(segments)]


def shard_blocks(segs, bs, world, rank):
    """Blocks b % world == rank of the joint stream (the concatenation of `segs`), concatenated: the bytes this rank owns."""
    from kanzi_b200 import sharding
    n_seg = len(segs[0])
    nb_seg = (n_seg + bs - 1) // bs
    assert n_seg % bs == 0 or len(segs) == 1, "segments must be whole blocks (only the joint stream's last block may be short)"
    parts = []
    for b in sharding.blocks_of_rank(nb_seg * len(segs), world, rank):
        s, bb = divmod(b, nb_seg)
        parts.append(segs[s][bb * bs: min(n_seg, (bb + 1) * bs)])
    return np.ascontiguousarray(np.concatenate(parts))


def record_crcs(knz, K):
    """(bit lengths, CRC32 of every block record's bits) of a .knz held in a numpy array"""
    from kanzi_b200 import sharding
    hb, off, bits = K.stream_index(knz)
    crcs = [zlib.crc32(sharding.extract_bits(knz, int(o), int(b)).tobytes()) for o, b in zip(off, bits)]
    return hb, np.asarray(bits, dtype=np.int64), np.asarray(crcs, dtype=np.int64)


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from kanzi_b200 import synth
    gen, full, transforms, entropy, bs = synth.CONFIGS[a.config]
    scale = a.scale if a.scale is not None else DEFAULT_SCALE[a.config]
    n = max(bs, int(full * scale))
    if world > 1 or a.config != "cfg2":
        n = max(bs, n // bs * bs) if (world > 1) else n          # joint streams are made of whole-block segments
    flags = 0 if a.bwt_fixed else 1
    seed = SEEDS[a.config]
    cfgd = config_dict(a.config, transforms, entropy, bs, n, gen.__name__, flags, max(world, a.gpus))

    if a.impl == "reference":
        if rank != 0:
            return 0
        data = gen(min(n, int(a.cpu_sample_mb * 1e6) + bs), seed)
        r = cpu_arm(data, transforms, entropy, bs, a.cpu_sample_mb, flags, steps=max(1, a.steps), warmup=min(a.warmup, 1))
        line = {"metric": "encode+decode MB/s", "value": round(r["value"], 2), "unit": "MB/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": round(r["ms_per_step"], 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "impl": "reference", "config": cfgd,
                "note": "CPU restatement (oracle/) of the reference Java path, one block per host thread; no JVM in the image",
                "cpu_baseline": {k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items() if k != "ms_per_step"},
                "e2e": {"value": round(r["value"], 2), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "encode_MBps": round(r["encode_MBps"], 2), "decode_MBps": round(r["decode_MBps"], 2), "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import kanzi_b200 as K
    from kanzi_b200 import sharding
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    K.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L = K.lib()
    L.kzg_set_profiling.argtypes = [C.c_int]
    L.kzg_profile_json.restype = C.c_char_p
    kstream = torch.cuda.ExternalStream(L.kzg_stream(), device=dev)

    base = gen(n, seed)                                   # every rank derives the same joint stream from the same base segment
    segs = joint_segments(base, world)
    nb_seg = (n + bs - 1) // bs
    nb_joint = nb_seg * world
    data = shard_blocks(segs, bs, world, rank) if world > 1 else base
    cod = Codec(torch, K, dev, data, transforms, entropy, bs, flags)
    my_blocks = sharding.blocks_of_rank(nb_joint, world, rank)

    # ---- correctness gates before any timing ------------------------------------------------------------------------------
    knz_len = cod.gate()
    if rank == 0:
        # this rank's .knz against the oracle's, byte for byte (bounded: the oracle is one CPU thread)
        import oracle_lib as O
        chk = len(data) if len(data) <= (256 << 20) else 8 * bs
        ref = O.compress(data[:chk], transforms, entropy, bs, bwt_bounds=flags)
        got = cod.h_knz[:knz_len].numpy().tobytes() if chk == len(data) else K.compress(data[:chk], transforms, entropy, bs, flags=flags)
        assert got == ref, "GPU .knz differs from the oracle's"
    joint_check = None
    if world > 1:
        # every rank: bit length + CRC of each of its records; rank 0: the same from the whole joint stream encoded on its own GPU
        _, my_bits, my_crc = record_crcs(cod.h_knz[:knz_len].numpy(), K)
        assert len(my_bits) == len(my_blocks)
        per = (nb_joint + world - 1) // world
        mine = torch.zeros(2 * per, dtype=torch.int64, device=dev)
        mine[:len(my_bits)] = torch.from_numpy(my_bits).to(dev)
        mine[per:per + len(my_crc)] = torch.from_numpy(my_crc).to(dev)
        parts = [torch.zeros_like(mine) for _ in range(world)]
        torch.distributed.all_gather(parts, mine)
        if rank == 0:
            joint = np.ascontiguousarray(np.concatenate(segs))
            jk = np.frombuffer(K.compress(joint, transforms, entropy, bs, flags=flags), dtype=np.uint8)
            hb, jbits, jcrc = record_crcs(jk, K)
            assert len(jbits) == nb_joint
            bad = 0
            for r in range(world):
                pr = parts[r].cpu().numpy()
                for i, b in enumerate(sharding.blocks_of_rank(nb_joint, world, r)):
                    if pr[i] != jbits[b] or pr[per + i] != jcrc[b]:
                        bad += 1
            assert bad == 0, f"{bad} block records of the sharded encode differ from the single-GPU joint stream"
            offs, end = sharding.stream_bit_offsets(hb, [int(x) for x in jbits])
            assert (end + 8 + 7) // 8 == len(jk), "joint stream length from the gathered bit lengths differs from the single-GPU stream"
            joint_check = {"blocks": int(nb_joint), "joint_knz_bytes": int(len(jk)), "records_match_single_gpu_stream": True}
            del joint, jk
        torch.distributed.barrier()
    del segs

    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def make_gather(nblocks_mine, nblocks_total):
        per = (nblocks_total + world - 1) // world
        bits_dev = torch.zeros(per, dtype=torch.int64, device=dev)
        outs = [torch.zeros_like(bits_dev) for _ in range(world)]

        def gather():
            """the path's one exchange: all ranks learn every record's bit length -> every record's offset in the joint stream"""
            bl = K.last_block_bits()
            bits_dev[:len(bl)] = torch.from_numpy(bl).to(dev, non_blocking=True)
            torch.distributed.all_gather(outs, bits_dev)
            return outs
        return gather

    gather_main = make_gather(len(my_blocks), nb_joint) if world > 1 else None

    def step_resident(c, gather):
        l2_flush.zero_()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(kstream)
        k = c.enc_dev()
        if gather is not None:
            gather()
        e1.record(kstream)
        c.dec_dev(k)
        e2.record(kstream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), e1.elapsed_time(e2), list(c.tim_e), list(c.tim_d)

    def step_e2e(c, gather):
        l2_flush.zero_()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(kstream)
        k = c.enc_host()
        if gather is not None:
            gather()
        e1.record(kstream)
        c.dec_host(k)
        e2.record(kstream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), e1.elapsed_time(e2), k

    def rmax(x):          # max over ranks of a per-rank scalar
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wait = time.time()
    while not sampler.rows and time.time() - t_wait < 3.0:      # nvidia-smi takes a moment to print its first line
        time.sleep(0.02)
    for _ in range(a.warmup):
        step_resident(cod, gather_main)
    barrier()
    sampler.mark()
    K.launch_count(reset=True)
    t_enc, t_dec, stage_e, stage_d = [], [], [], []
    for _ in range(a.steps):
        te, td, se, sd = step_resident(cod, gather_main)
        t_enc.append(te); t_dec.append(td); stage_e.append(se); stage_d.append(sd)
    barrier()
    launches = K.launch_count()
    sampler.mark()
    # e2e (host buffers), fewer repetitions of the same workload
    for _ in range(min(a.warmup, 1)):
        step_e2e(cod, gather_main)
    e_enc, e_dec = [], []
    for _ in range(max(1, min(a.steps, 3))):
        te, td, k = step_e2e(cod, gather_main)
        e_enc.append(te); e_dec.append(td)
    assert np.array_equal(cod.h_back[:cod.n].numpy(), data), "e2e round trip mismatch"
    assert k == knz_len and torch.equal(cod.h_knz[:k], cod.d_knz[:k].cpu()), "the host-buffer entry's .knz differs from the device-resident one (which equals the oracle's)"
    barrier()
    clocks = sampler.stop()
    # per-kernel events: one extra (untimed) step with the library's profiling on
    # (block groups off for that step: one launch per kernel over all blocks, nothing else on the GPU while it runs, so an event
    #  pair brackets the kernel alone; in the timed steps the groups overlap and a kernel's wall time is not its own)
    os.environ["KZG_LZ_GROUPS"] = "1"
    os.environ["KZG_DEC_GROUPS"] = "1"
    cod.enc_dev()
    L.kzg_set_profiling(1)
    kk = cod.enc_dev()
    cod.dec_dev(kk)
    prof = json.loads(L.kzg_profile_json().decode())
    L.kzg_set_profiling(0)
    del os.environ["KZG_LZ_GROUPS"], os.environ["KZG_DEC_GROUPS"]

    enc_ms, dec_ms = rmax(sum(t_enc) / len(t_enc)), rmax(sum(t_dec) / len(t_dec))
    e2e_enc_ms, e2e_dec_ms = rmax(sum(e_enc) / len(e_enc)), rmax(sum(e_dec) / len(e_dec))

    # ---- strong scaling: a fixed 408-block stream (8 segments of cfg2) dealt over the N ranks ---------------------------------
    strong = None
    if a.config == "cfg2" and scale == 1.0 and not a.no_strong and STRONG_SEGMENTS % world == 0:
        if world == STRONG_SEGMENTS:
            strong = {"total_blocks": nb_joint, "same_as_main": True}
        else:
            del cod
            torch.cuda.empty_cache()
            segs8 = joint_segments(base[: n // bs * bs], STRONG_SEGMENTS)      # whole-block segments: 8 x 50 blocks
            sdata = shard_blocks(segs8, bs, world, rank)
            del segs8
            scod = Codec(torch, K, dev, sdata, transforms, entropy, bs, flags)
            scod.gate()
            nb8 = (n // bs) * STRONG_SEGMENTS
            sg = make_gather(len(sdata) // bs, nb8) if world > 1 else None
            step_resident(scod, sg)
            st = [step_resident(scod, sg) for _ in range(max(1, min(a.steps, 3)))]
            s_enc, s_dec = rmax(sum(x[0] for x in st) / len(st)), rmax(sum(x[1] for x in st) / len(st))
            tot8 = (n // bs * bs) * STRONG_SEGMENTS
            strong = {"total_blocks": nb8, "total_bytes": int(tot8), "ms_per_step": round(s_enc + s_dec, 3),
                      "value": round(tot8 / 1e6 / ((s_enc + s_dec) * 1e-3), 2), "unit": "MB/s",
                      "note": "fixed total work: efficiency(N) = value(N) / (N * value(1)) over these keys"}
            del scod
    if strong and strong.get("same_as_main"):
        strong.update({"total_bytes": int(n * world), "ms_per_step": round(enc_ms + dec_ms, 3), "value": round(world * n / 1e6 / ((enc_ms + dec_ms) * 1e-3), 2), "unit": "MB/s"})

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return 0
    total_mb = world * n / 1e6
    step_ms = enc_ms + dec_ms
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    xf_ms = sum(s[0] for s in stage_e) / len(stage_e)
    ent_ms = sum(s[1] for s in stage_e) / len(stage_e)
    asm_ms = sum(s[2] for s in stage_e) / len(stage_e)
    dxf_ms = sum(s[0] for s in stage_d) / len(stage_d)
    dent_ms = sum(s[1] for s in stage_d) / len(stage_d)
    stages = {"enc_transform_ms": xf_ms, "enc_entropy_ms": ent_ms, "enc_container_ms": asm_ms, "dec_entropy_ms": dent_ms, "dec_transform_ms": dxf_ms}
    dom = max(stages, key=stages.get)
    # algorithmic bytes: a stage / kernel reads its input once and writes its output once.  The transform stage's
    # output is the pre-entropy length; it is not visible from outside the library, so it is bounded by the .knz size below and
    # n above: the stage figure uses n + knz (a lower bound of the bytes, i.e. the fraction is not flattered)
    nloc = len(data)
    alg_stage = {"enc_transform_ms": nloc + knz_len, "enc_entropy_ms": nloc + knz_len, "enc_container_ms": 2 * knz_len,
                 "dec_entropy_ms": knz_len + nloc, "dec_transform_ms": knz_len + nloc}
    # per-kernel roofline: the top kernel by summed CUDA-event time of one profiled step
    enc_side = lambda name: name.startswith("lzf_") or "encode" in name
    kern = sorted(prof.items(), key=lambda kv: -kv[1][1])
    ktable = []
    for name, (cnt, ms) in kern[:6]:
        alg = nloc if enc_side(name) else (knz_len + nloc)
        ktable.append({"kernel": name, "launches": cnt, "total_ms": round(ms, 3), "algorithmic_bytes": int(alg), "achieved_GBps": round(alg / (ms * 1e-3) / 1e9, 3) if ms > 0 else None})
    top = ktable[0] if ktable else None
    traffic, traffic_src = None, None
    try:      # DRAM bytes of the top kernel per launch set, from the committed ncu launch list of this command (profiles/)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic_cfg2.json")))
        if a.config == "cfg2" and scale == 1.0 and top and top["kernel"] in tj:
            traffic, traffic_src = int(tj[top["kernel"]]), tj.get("_source")
    except Exception:
        pass
    achieved = top["achieved_GBps"] if top else 0.0
    line = {"metric": "encode+decode MB/s", "value": round(total_mb / (step_ms * 1e-3), 2), "unit": "MB/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": round(step_ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": cfgd,
            "stream": {"knz_bytes_rank0": int(knz_len), "ratio": round(nloc / max(knz_len, 1), 3), "blocks_rank0": len(my_blocks), "blocks_total": nb_joint},
            "encode_MBps": round(total_mb / (enc_ms * 1e-3), 2), "decode_MBps": round(total_mb / (dec_ms * 1e-3), 2),
            "stages_ms": {k: round(v, 3) for k, v in stages.items()},
            "e2e": {"value": round(total_mb / ((e2e_enc_ms + e2e_dec_ms) * 1e-3), 2), "unit": "MB/s",
                    "h2d_bytes_per_step": int(nloc + knz_len), "d2h_bytes_per_step": int(knz_len + nloc),
                    "encode_MBps": round(total_mb / (e2e_enc_ms * 1e-3), 2), "decode_MBps": round(total_mb / (e2e_dec_ms * 1e-3), 2)},
            "roofline": {"bound": "hbm", "kernel": top["kernel"] if top else None, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 6) if achieved else None, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": top["algorithmic_bytes"] if top else None,
                         "how": "CUDA events around every launch of the kernel on its own stream during one extra profiled step (kzg_set_profiling); launches of one "
                                "step summed (the blocks are dealt into groups, one launch per group and round); algorithmic bytes = the bytes the kernel's step must read + write once",
                         "kernels": ktable,
                         "decode": (lambda dms: {"ms": round(dms, 3), "algorithmic_bytes": int(knz_len + nloc), "achieved": round((knz_len + nloc) / (dms * 1e-3) / 1e9, 3),
                                                 "frac": round((knz_len + nloc) / (dms * 1e-3) / 1e9 / peak, 6),
                                                 "top_kernel": next((k["kernel"] for k in ktable if not enc_side(k["kernel"])), None),
                                                 "note": "whole decode (entropy + inverse transform) of this rank's shard: coded bytes in + original bytes out over its time"})(dent_ms + dxf_ms),
                         "stage": {"name": dom, "ms": round(stages[dom], 3), "algorithmic_bytes": int(alg_stage[dom]),
                                   "achieved": round(alg_stage[dom] / (stages[dom] * 1e-3) / 1e9, 3) if stages[dom] > 0 else None,
                                   "frac": round(alg_stage[dom] / (stages[dom] * 1e-3) / 1e9 / peak, 6) if stages[dom] > 0 else None}},
            "clocks": clocks, "gpu_launches": int(launches)}
    if joint_check:
        line["joint_stream_check"] = joint_check
    if strong:
        line["strong_scaling"] = strong
    if world == 1 and a.config == "cfg2" and scale == 1.0 and not a.no_others:
        try:
            line["other_configs"] = other_configs(torch, K, dev, kstream, flags, a)
        except Exception as e:       # never lose the headline line to a side measurement
            line["other_configs"] = {"error": str(e)[:300]}
    if not a.no_cpu_baseline:
        try:
            r = cpu_arm(data, transforms, entropy, bs, a.cpu_sample_mb, flags)
            line["cpu_baseline"] = {k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items() if k != "ms_per_step"}
        except Exception as e:
            line["cpu_baseline"] = {"error": str(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


# passes of the other BASELINE.json configs (parity-test cases, not the headline): value / e2e / CPU arm on a sample of the same data.
# Their codecs are one dependent chain per block or chunk (inverse RANK / SRT, FPAQ, ROLZ, the four rANS states of a 4 MiB order-1
# chunk): a lone warp runs such a chain at about four cycles per instruction, so what these lines measure is how many blocks are in
# flight.  cfg3 runs whole (12 blocks); cfg5 runs 128 of its 512 blocks (2 GiB: a 256 MiB base from the generator, rotated by odd
# amounts so every block differs; the generator itself costs 7 s per 256 MiB); cfg4 (FPAQ: ~25 s per 32 MiB block each way) codes
# ONE short block of 4 MiB under -b 32M; full 32 MiB blocks are exercised by tests/test_gpu_fullsize.py and `--config cfg4`.
OTHER = {"cfg1": 1.0, "cfg3": 1.0, "cfg4": 0.0042, "cfg5": 0.25}
CFG5_BASE = 256 << 20


def other_configs(torch, K, dev, kstream, flags, a):
    from kanzi_b200 import synth
    out = {}
    for cfg, sc in OTHER.items():
        gen, full, transforms, entropy, bs = synth.CONFIGS[cfg]
        n = max(bs, int(full * sc) // bs * bs) if cfg != "cfg4" else (4 << 20)
        if cfg == "cfg3":
            n = int(full * sc)
        if cfg == "cfg5" and n > CFG5_BASE:
            n = n // CFG5_BASE * CFG5_BASE
            base5 = gen(CFG5_BASE, SEEDS[cfg])
            data = np.concatenate([base5 if i == 0 else np.roll(base5, i * 1000003) for i in range(n // CFG5_BASE)])
            del base5
        else:
            data = gen(n, SEEDS[cfg])
        c = Codec(torch, K, dev, data, transforms, entropy, bs, flags)
        t0 = time.time()
        k = c.gate()
        gate_s = time.time() - t0
        # chains whose codecs are one dependent chain per block (FPAQ, inverse RANK / SRT, ANS1) take seconds per pass at these
        # block counts: one timed pass each is all the default run can afford
        reps = 1 if gate_s > 2.0 else 3
        res, e2e = [], []
        for _ in range(reps):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            torch.cuda.synchronize()
            e0.record(kstream); kk = c.enc_dev(); e1.record(kstream); c.dec_dev(kk); e2.record(kstream)
            torch.cuda.synchronize()
            res.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
        for _ in range(reps if gate_s <= 25.0 else 0):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(kstream); kk = c.enc_host(); e1.record(kstream); c.dec_host(kk); e2.record(kstream)
            torch.cuda.synchronize()
            e2e.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
        if e2e:
            assert np.array_equal(c.h_back[:n].numpy(), data)
        em, dm = sum(x[0] for x in res) / reps, sum(x[1] for x in res) / reps
        eem, edm = (sum(x[0] for x in e2e) / len(e2e), sum(x[1] for x in e2e) / len(e2e)) if e2e else (None, None)
        mb = n / 1e6
        entry = {"workload": f"{cfg}: {'+'.join(transforms)}&{entropy} -b {bs}, {n} bytes ({(n + bs - 1) // bs} blocks), scale {sc:.4g} of the config",
                 "value": round(mb / ((em + dm) * 1e-3), 2), "encode_MBps": round(mb / (em * 1e-3), 2), "decode_MBps": round(mb / (dm * 1e-3), 2),
                 "e2e": round(mb / ((eem + edm) * 1e-3), 2) if e2e else None, "knz_bytes": int(k), "unit": "MB/s",
                 "hbm_frac_stream": round((n + k) / ((em + dm) * 1e-3) / 1e9 / 6549.1, 6)}
        if not a.no_cpu_baseline:
            # CPU arm: one block per host thread where the config has that many (bounded: about 30 s of CPU work)
            cores = os.cpu_count() or 1
            r = cpu_arm(data, transforms, entropy, bs, min(max(64.0, cores * bs / 1e6), 600.0), flags)
            entry["cpu_reference_MBps"] = round(r["value"], 2)
            entry["cpu_sample"] = r["sample"]
        out[cfg] = entry
        del c, data
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    sys.exit(main())
