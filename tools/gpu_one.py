"""Developer aid: one transform / entropy call (for compute-sanitizer runs)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kanzi_b200 as K, oracle_lib as O, corpus
kind, name = sys.argv[1], sys.argv[2]
d = corpus.small_cases()[name]
cap = len(d) + len(d) // 64 + 1100
okr, ref, _, _ = O.transform(kind, d, dst_cap=cap, ctx=[7, max(len(d), 1024), len(d), 1, 0, 0])
ok, got, used = K.transform_forward(kind, d, {"blockSize": max(len(d), 1024), "size": len(d), "flags": 0}, dst_cap=cap)
print("forward", ok, okr, got == ref, len(got), len(ref))
if okr:
    ok2, back, _ = K.transform_inverse(kind, ref, {"blockSize": max(len(d), 1024), "flags": 0}, dst_cap=len(d) + 512)
    print("inverse", ok2, back == d)
