"""Developer aid: cfg2 encode/decode on the device a few times, printing stage times (and, with KZG_DEBUG=17, the LZ forward
phase timers and stitch statistics on stderr).  Usage: [KZG_DEBUG=17] [KZG_LZ_SEG=n] python tools/gpu_cfg2_dbg.py [scale] [reps]"""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import kanzi_b200 as K
from kanzi_b200 import synth
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = sys.argv[3] if len(sys.argv) > 3 else "cfg2"
gen, full, transforms, entropy, bs = synth.CONFIGS[cfg]
n = max(bs, int(full * scale))
K.set_device(0)
L = K.lib()
dev = torch.device("cuda", 0)
data = gen(n, {"cfg1": 1, "cfg2": 2, "cfg3": 3, "cfg4": 4, "cfg5": 5}[cfg])
d_in = torch.zeros(n + 256, dtype=torch.uint8, device=dev)
d_in[:n].copy_(torch.from_numpy(data))
cap = int(K.compress_bound(n, bs))
d_knz = torch.zeros(cap + 256, dtype=torch.uint8, device=dev)
d_back = torch.zeros(n + bs + 256, dtype=torch.uint8, device=dev)
h_knz = torch.zeros(cap, dtype=torch.uint8).pin_memory()
ids = (C.c_int32 * 8)(*([K.T[t] for t in transforms] + [0] * (8 - len(transforms))))
te, td = (C.c_float * 3)(), (C.c_float * 3)()
u8p = C.POINTER(C.c_uint8)
flags = 1
for i in range(reps):
    k = L.kzg_compress_dev(d_in.data_ptr(), n, ids, len(transforms), K.E[entropy], bs, flags, d_knz.data_ptr(), cap, te)
    assert k > 0, k
    h_knz[:k].copy_(d_knz[:k]); torch.cuda.synchronize()
    r = L.kzg_decompress_dev(d_knz.data_ptr(), k, C.cast(h_knz.data_ptr(), u8p), flags, d_back.data_ptr(), n + bs, td)
    assert r == n, r
    print(f"rep {i}: knz {k}  enc xf/ent/asm {te[0]:.2f} {te[1]:.2f} {te[2]:.2f} ms   dec xf/ent {td[0]:.2f} {td[1]:.2f} ms", flush=True)
assert torch.equal(d_back[:n], d_in[:n])
print("round trip ok")
