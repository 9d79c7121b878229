"""Developer aid: per-launch timeline (CUDA events on the launching streams, KZG_PROF sites) of one cfg2 encode + decode with
the block groups overlapping as in production.  Usage: python tools/gpu_timeline.py [scale] > timeline.txt"""
import sys, os, json, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["KZG_TIMELINE"] = "1"
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
import kanzi_b200 as K
from kanzi_b200 import synth
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
gen, full, transforms, entropy, bs = synth.CONFIGS["cfg2"]
n = max(bs, int(full * scale))
K.set_device(0)
L = K.lib()
L.kzg_set_profiling.argtypes = [C.c_int]
L.kzg_profile_json.restype = C.c_char_p
dev = torch.device("cuda", 0)
data = gen(n, 2)
d_in = torch.zeros(n + 256, dtype=torch.uint8, device=dev); d_in[:n].copy_(torch.from_numpy(data))
cap = int(K.compress_bound(n, bs))
d_knz = torch.zeros(cap + 256, dtype=torch.uint8, device=dev)
d_back = torch.zeros(n + 256, dtype=torch.uint8, device=dev)
h_knz = torch.zeros(cap, dtype=torch.uint8).pin_memory()
ids = (C.c_int32 * 8)(*([K.T[t] for t in transforms] + [0] * (8 - len(transforms))))
te, td = (C.c_float * 3)(), (C.c_float * 3)()
u8p = C.POINTER(C.c_uint8)
for rep in range(3):
    if rep == 2:
        L.kzg_set_profiling(1)
    k = L.kzg_compress_dev(d_in.data_ptr(), n, ids, 1, K.E[entropy], bs, 1, d_knz.data_ptr(), cap, te)
    if rep == 2:
        enc = json.loads(L.kzg_profile_json().decode()); L.kzg_set_profiling(1)
    h_knz[:k].copy_(d_knz[:k]); torch.cuda.synchronize()
    r = L.kzg_decompress_dev(d_knz.data_ptr(), k, C.cast(h_knz.data_ptr(), u8p), 1, d_back.data_ptr(), n, td)
    assert r == n
    if rep == 2:
        dec = json.loads(L.kzg_profile_json().decode()); L.kzg_set_profiling(0)
print(f"encode {te[0]:.2f} {te[1]:.2f} {te[2]:.2f} ms, decode xf {td[0]:.2f} ent {td[1]:.2f} ms")
for nm, tl in (("ENCODE", enc), ("DECODE", dec)):
    print(nm)
    for name, a, b in sorted(tl["_timeline"], key=lambda x: x[1]):
        print(f"  {a:8.3f} -> {b:8.3f}  ({b - a:7.3f})  {name}")
