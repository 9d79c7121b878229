"""Developer aid: encode + decode passes of one BASELINE config over `nbytes` of its synthetic workload, device-resident, with host
wall-clock and CUDA-event times per call (run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list).
Usage: python tools/gpu_cfg_pass.py cfg3 16777216 [reps] [--fixed] [--chain LZX HUFFMAN 4194304]"""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
import kanzi_b200 as K
from kanzi_b200 import synth
cfg = sys.argv[1]
gen, full, transforms, entropy, bs = synth.CONFIGS[cfg]
n = int(sys.argv[2]) if len(sys.argv) > 2 else min(full, 4 * bs)
reps = int(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else 2
flags = 0 if "--fixed" in sys.argv else 1
if "--chain" in sys.argv:          # --chain LZX HUFFMAN 4194304: the config's data under another chain / block size
    k = sys.argv.index("--chain")
    transforms, entropy, bs = sys.argv[k + 1].split("+"), sys.argv[k + 2], int(sys.argv[k + 3])
seed = {"cfg1": 1, "cfg2": 2, "cfg3": 3, "cfg4": 4, "cfg5": 5}[cfg]
K.set_device(0)
L = K.lib()
dev = torch.device("cuda", 0)
BASE = 256 << 20
if n > BASE and n % BASE == 0:      # large runs: a 256 MiB base rotated by odd amounts (every block differs, same statistics; the generator is slow)
    import numpy as np
    base = gen(BASE, seed)
    data = np.concatenate([base if i == 0 else np.roll(base, i * 1000003) for i in range(n // BASE)])
else:
    data = gen(n, seed)
d_in = torch.zeros(n + 256, dtype=torch.uint8, device=dev); d_in[:n].copy_(torch.from_numpy(data))
cap = int(K.compress_bound(n, bs))
d_knz = torch.zeros(cap + 256, dtype=torch.uint8, device=dev)
d_back = torch.zeros(n + 256, dtype=torch.uint8, device=dev)
h_knz = torch.zeros(cap, dtype=torch.uint8).pin_memory()
ids = (C.c_int32 * 8)(*([K.T[t] for t in transforms] + [0] * (8 - len(transforms))))
te, td = (C.c_float * 3)(), (C.c_float * 3)()
u8p = C.POINTER(C.c_uint8)
for rep in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    k = L.kzg_compress_dev(d_in.data_ptr(), n, ids, len(transforms), K.E[entropy], bs, flags, d_knz.data_ptr(), cap, te)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    assert k > 0, k
    h_knz[:k].copy_(d_knz[:k]); torch.cuda.synchronize()
    t2 = time.perf_counter()
    r = L.kzg_decompress_dev(d_knz.data_ptr(), k, C.cast(h_knz.data_ptr(), u8p), flags, d_back.data_ptr(), n, td)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    assert r == n, r
    print(f"{cfg} n={n} ({(n + bs - 1) // bs} blocks) knz={k}: encode wall {1e3 * (t1 - t0):.3f} ms (xf {te[0]:.3f} ent {te[1]:.3f} cont {te[2]:.3f}) = {n / 1e6 / (t1 - t0):.1f} MB/s; "
          f"decode wall {1e3 * (t3 - t2):.3f} ms (xf {td[0]:.3f} ent {td[1]:.3f}) = {n / 1e6 / (t3 - t2):.1f} MB/s", flush=True)
assert torch.equal(d_back[:n], d_in[:n])
