#!/usr/bin/env python
"""bench.py — the hot path of BASELINE.json on synthetic blocks: encode + decode MB/s, bit-exact .knz.

One step = one pass of the hot path over one batch: every block of the workload is encoded
(transform chain + entropy coder + container assembly -> .knz) and the .knz is decoded back.
  value : whole-job MB/s (10^6 original bytes / step time) with inputs already resident in HBM
  e2e   : the same through the host-buffer C ABI (pinned H2D of the input / D2H of the .knz inside the
          timed region on the way in, H2D of the .knz / D2H of the decoded bytes on the way back)
  roofline      : dominant kernel vs the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  : the C++ restatement of the reference (oracle/) on the box's host cores, bounded sample
`--impl reference` times that CPU path as the reference arm (no JVM exists in the image: SURVEY.md §0.2).
Launch: python bench.py [--gpus N --steps K --warmup W]; N > 1 via torchrun (one rank per GPU, blocks sharded
per rank, NCCL only for the per-block size gather)."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

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
    ap.add_argument("--bwt-fixed", action="store_true", help="BWT bounds 'fixed' instead of as the reference is written")
    return ap.parse_args()


DEFAULT_SCALE = {"cfg1": 1.0, "cfg2": 1.0, "cfg3": 1.0, "cfg4": 0.1, "cfg5": 1.0 / 32}


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


def cpu_arm(data, transforms, entropy, bs, sample_mb, flags, steps=1, warmup=0):
    """The reference's CPU path for this chain = the oracle port, one EncodingTask/DecodingTask per host thread."""
    import oracle_lib as O
    cores = os.cpu_count() or 1
    nblk = max(1, int(sample_mb * 1e6) // bs)
    sample = np.ascontiguousarray(data[: min(len(data), nblk * bs)])
    enc_t, dec_t = [], []
    # buffers allocated (and touched) outside the timed region, as the GPU arm's are
    obuf = np.ones(len(sample) + len(sample) // 4 + (1 << 16) + 64 * (len(sample) // bs + 1), dtype=np.uint8)
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


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from kanzi_b200 import synth
    gen, full, transforms, entropy, bs = synth.CONFIGS[a.config]
    scale = a.scale if a.scale is not None else DEFAULT_SCALE[a.config]
    n = max(bs, int(full * scale))
    flags = 0 if a.bwt_fixed else 1
    workload = f"{a.config}: {'+'.join(transforms)}&{entropy} -b {bs}, {n} bytes ({(n + bs - 1) // bs} blocks) per GPU, synthetic {gen.__name__}"
    seed = {"cfg1": 1, "cfg2": 2, "cfg3": 3, "cfg4": 4, "cfg5": 5}[a.config]

    if a.impl == "reference":
        if rank != 0:
            return 0
        data = gen(min(n, int(a.cpu_sample_mb * 1e6) + bs), seed)
        r = cpu_arm(data, transforms, entropy, bs, a.cpu_sample_mb, flags, steps=max(1, a.steps), warmup=min(a.warmup, 1))
        line = {"metric": "encode+decode MB/s", "value": round(r["value"], 2), "unit": "MB/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": round(r["ms_per_step"], 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "impl": "reference",
                "config": {"workload": workload, "note": "CPU restatement (oracle/) of the reference Java path; no JVM in the image"},
                "cpu_baseline": {k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items() if k != "ms_per_step"},
                "e2e": {"value": round(r["value"], 2), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "encode_MBps": round(r["encode_MBps"], 2), "decode_MBps": round(r["decode_MBps"], 2), "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import kanzi_b200 as K
    from kanzi_b200 import binding
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    K.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L = K.lib()
    kstream = torch.cuda.ExternalStream(L.kzg_stream(), device=dev)

    data = gen(n, seed + 1000 * rank)                       # each rank owns its shard of blocks (weak scaling)
    h_in = torch.from_numpy(data).pin_memory()
    d_in = torch.zeros(n + 256, dtype=torch.uint8, device=dev)
    d_in[:n].copy_(h_in)
    cap = int(K.compress_bound(n, bs))
    d_knz = torch.zeros(cap + 256, dtype=torch.uint8, device=dev)
    d_back = torch.zeros(n + bs + 256, dtype=torch.uint8, device=dev)
    h_knz = torch.zeros(cap, dtype=torch.uint8).pin_memory()
    h_back = torch.zeros(n + 256, dtype=torch.uint8).pin_memory()
    ids = (C.c_int32 * 8)(*([K.T[t] for t in transforms] + [0] * (8 - len(transforms))))
    nT, eid = len(transforms), K.E[entropy]
    u8p = C.POINTER(C.c_uint8)
    tim_e = (C.c_float * 3)()
    tim_d = (C.c_float * 3)()
    torch.cuda.synchronize()

    def enc_dev():
        r = L.kzg_compress_dev(d_in.data_ptr(), n, ids, nT, eid, bs, flags, d_knz.data_ptr(), cap, tim_e)
        if r < 0:
            raise K.KzgError(r, "kzg_compress_dev")
        return r

    def dec_dev(knz_len):
        r = L.kzg_decompress_dev(d_knz.data_ptr(), knz_len, C.cast(h_knz.data_ptr(), u8p), flags, d_back.data_ptr(), n + bs, tim_d)
        if r < 0:
            raise K.KzgError(r, "kzg_decompress_dev")
        return r

    # correctness gate before any timing: device round trip and bit-exactness of one block record vs the oracle
    knz_len = enc_dev()
    h_knz[:knz_len].copy_(d_knz[:knz_len])
    torch.cuda.synchronize()
    assert dec_dev(knz_len) == n
    assert torch.equal(d_back[:n], d_in[:n]), "GPU round trip mismatch"
    if rank == 0:
        # the whole .knz the timed path produces must be the oracle's, byte for byte (bounded: the oracle is one CPU thread)
        import oracle_lib as O
        chk = n if n <= (256 << 20) else 8 * bs
        ref = O.compress(data[:chk], transforms, entropy, bs, bwt_bounds=flags)
        got = h_knz[:knz_len].numpy().tobytes() if chk == n else K.compress(data[:chk], transforms, entropy, bs, flags=flags)
        assert got == ref, "GPU .knz differs from the oracle's"

    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    sizes = torch.zeros(1, dtype=torch.int64, device=dev)

    def step_resident(timed):
        l2_flush.zero_()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(kstream)
        k = enc_dev()
        e1.record(kstream)
        dec_dev(k)
        e2.record(kstream)
        if world > 1:            # the path's one exchange: gather the per-shard encoded sizes (block offsets in the joint stream)
            sizes[0] = k
            out = [torch.zeros_like(sizes) for _ in range(world)]
            torch.distributed.all_gather(out, sizes)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), e1.elapsed_time(e2), list(tim_e), list(tim_d)

    def step_e2e():
        l2_flush.zero_()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(kstream)
        k = L.kzg_compress(C.cast(h_in.data_ptr(), u8p), n, ids, nT, eid, bs, flags, C.cast(h_knz.data_ptr(), u8p), cap)
        if k < 0:
            raise K.KzgError(k, "kzg_compress")
        e1.record(kstream)
        r = L.kzg_decompress(C.cast(h_knz.data_ptr(), u8p), k, flags, C.cast(h_back.data_ptr(), u8p), n)
        if r != n:
            raise K.KzgError(r, "kzg_decompress")
        e2.record(kstream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), e1.elapsed_time(e2), k

    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wait = time.time()
    while not sampler.rows and time.time() - t_wait < 3.0:      # nvidia-smi takes a moment to print its first line
        time.sleep(0.02)
    for _ in range(a.warmup):
        step_resident(False)
    barrier()
    sampler.mark()
    K.launch_count(reset=True)
    t_enc, t_dec, stage_e, stage_d = [], [], [], []
    for _ in range(a.steps):
        te, td, se, sd = step_resident(True)
        t_enc.append(te); t_dec.append(td); stage_e.append(se); stage_d.append(sd)
    barrier()
    launches = K.launch_count()
    sampler.mark()
    # e2e (host buffers), fewer repetitions of the same workload
    for _ in range(min(a.warmup, 1)):
        step_e2e()
    e_enc, e_dec = [], []
    for _ in range(max(1, min(a.steps, 3))):
        te, td, k = step_e2e()
        e_enc.append(te); e_dec.append(td)
    assert np.array_equal(h_back[:n].numpy(), data), "e2e round trip mismatch"
    assert k == knz_len and torch.equal(h_knz[:k], d_knz[:k].cpu()), "the host-buffer entry's .knz differs from the device-resident one (which equals the oracle's)"
    barrier()
    clocks = sampler.stop()

    def rmax(x):          # max over ranks of a per-rank scalar
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    enc_ms, dec_ms = rmax(sum(t_enc) / len(t_enc)), rmax(sum(t_dec) / len(t_dec))
    e2e_enc_ms, e2e_dec_ms = rmax(sum(e_enc) / len(e_enc)), rmax(sum(e_dec) / len(e_dec))
    if rank != 0:
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
    # dominant kernel: the transform stage of encoding (lz_forward for cfg2); algorithmic bytes = stage input + output
    xf_ms = sum(s[0] for s in stage_e) / len(stage_e)
    ent_ms = sum(s[1] for s in stage_e) / len(stage_e)
    asm_ms = sum(s[2] for s in stage_e) / len(stage_e)
    dxf_ms = sum(s[0] for s in stage_d) / len(stage_d)
    dent_ms = sum(s[1] for s in stage_d) / len(stage_d)
    stages = {"enc_transform_ms": xf_ms, "enc_entropy_ms": ent_ms, "enc_container_ms": asm_ms, "dec_entropy_ms": dent_ms, "dec_transform_ms": dxf_ms}
    dom = max(stages, key=stages.get)
    # per-block algorithmic bytes: stage kernels read their input once and write their output once
    post_bytes = int(knz_len)      # entropy output ~ .knz size; transform output is between n and knz size: use n + knz as the chain's ideal traffic
    alg_bytes = {"enc_transform_ms": n + post_bytes, "enc_entropy_ms": n + post_bytes, "enc_container_ms": 2 * post_bytes,
                 "dec_entropy_ms": post_bytes + n, "dec_transform_ms": post_bytes + n}[dom]
    achieved = alg_bytes / (stages[dom] * 1e-3) / 1e9 if stages[dom] > 0 else 0.0
    traffic, traffic_src = None, None
    try:      # DRAM bytes of the stage's kernels per invocation, from the committed ncu launch list (full cfg2 only)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic_cfg2.json")))
        if a.config == "cfg2" and scale == 1.0:
            traffic, traffic_src = int(tj[dom]), tj.get("_source")
    except Exception:
        pass
    line = {"metric": "encode+decode MB/s", "value": round(total_mb / (step_ms * 1e-3), 2), "unit": "MB/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": round(step_ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": workload, "l2": "flushed between steps (256 MiB write); inputs also exceed L2", "bwt_bounds": "asref" if flags else "fixed",
                       "knz_bytes": int(knz_len), "ratio": round(n / max(knz_len, 1), 3)},
            "encode_MBps": round(total_mb / (enc_ms * 1e-3), 2), "decode_MBps": round(total_mb / (dec_ms * 1e-3), 2),
            "stages_ms": {k: round(v, 3) for k, v in stages.items()},
            "e2e": {"value": round(total_mb / ((e2e_enc_ms + e2e_dec_ms) * 1e-3), 2), "unit": "MB/s",
                    "h2d_bytes_per_step": int(n + knz_len), "d2h_bytes_per_step": int(knz_len + n),
                    "encode_MBps": round(total_mb / (e2e_enc_ms * 1e-3), 2), "decode_MBps": round(total_mb / (e2e_dec_ms * 1e-3), 2)},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 3), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 6),
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes),
                         "note": "the dominant 'kernel' is a stage (all LZ forward kernels of one encode, timed with CUDA events on the library stream); its groups overlap on side streams"},
            "clocks": clocks, "gpu_launches": int(launches)}
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


if __name__ == "__main__":
    sys.exit(main())
