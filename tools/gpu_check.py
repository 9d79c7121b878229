"""Developer aid: run every parity check without stopping at the first failure and print a compact report."""
import sys, os, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import kanzi_b200 as K
import oracle_lib as O
import corpus
from kanzi_b200 import synth

CASES = corpus.small_cases()
only = set(sys.argv[1:])
fails = 0
FAMILIES = ["NONE", "HUFFMAN", "ANS0", "ANS1", "FPAQ", "LZ", "LZX", "ROLZ", "ZRLT", "RANK", "MTFT", "SRT", "BWT", "STREAM"]
if not only:
    import subprocess
    tot = 0
    for fam in FAMILIES:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), fam], capture_output=True, text=True)
        out = r.stdout
        sys.stdout.write("".join(l + "\n" for l in out.splitlines() if not l.startswith("FAILURES:")))
        if r.returncode != 0 and "FAILURES:" not in out:
            print(f"FAIL family {fam} crashed rc={r.returncode}: {r.stderr[-500:]}")
            tot += 1
        for l in out.splitlines():
            if l.startswith("FAILURES:"):
                tot += int(l.split()[1])
        sys.stdout.flush()
    print("FAILURES:", tot)
    sys.exit(0)

def fd(a, b):
    n = min(len(a), len(b))
    x = np.frombuffer(a[:n], dtype=np.uint8) != np.frombuffer(b[:n], dtype=np.uint8)
    return int(np.argmax(x)) if x.any() else n

def check(label, fn):
    global fails
    t0 = time.time()
    try:
        msg = fn()
        ok = msg is None
    except Exception as e:
        ok, msg = False, f"EXC {type(e).__name__}: {e}"
    if not ok:
        fails += 1
    print(f"{'ok  ' if ok else 'FAIL'} {label} {'' if ok else msg} ({time.time()-t0:.2f}s)", flush=True)

ents = [e for e in ["NONE", "HUFFMAN", "ANS0", "ANS1", "FPAQ"] if not only or e in only]
inputs = list(CASES.items()) + [(f"lit{i}", x) for i, x in enumerate(corpus.ENTROPY_LITERALS)] + [("fib", corpus.fibonacci_chunk())]
for ent in ents:
    for name, d in inputs:
        def f():
            ref, rb = O.entropy_encode(ent, d)
            got, b = K.entropy_encode(ent, d)
            if b != rb: return f"bits {b} != {rb} (first diff byte {fd(got, ref)})"
            if got != ref: return f"payload differs at byte {fd(got, ref)} of {len(ref)}"
            out, r, used = K.entropy_decode(ent, ref, rb, len(d))
            if r != len(d) or used != rb or out != d: return f"decode r={r} used={used}/{rb} diff at {fd(out, d)}"
        check(f"entropy {ent} {name}[{len(d)}]", f)

trs = [t for t in ["LZ", "LZX", "ROLZ", "ZRLT", "RANK", "MTFT", "SRT", "BWT"] if not only or t in only]
for tr in trs:
    for name, d in CASES.items():
        def f():
            cap = len(d) + len(d) // 64 + 1100
            okr, ref, _, octx = O.transform(tr, d, dst_cap=cap, ctx=[7, max(len(d), 1024), len(d), 1, 0, 0])
            kctx = {"blockSize": max(len(d), 1024), "size": len(d), "flags": 0}
            ok, got, used = K.transform_forward(tr, d, kctx, dst_cap=cap)
            if int(ok) != okr: return f"forward returned {ok}, oracle {okr}"
            if not ok: return None
            if got != ref: return f"forward differs at byte {fd(got, ref)} len {len(got)} vs {len(ref)}"
            if kctx["dataType"] != octx[4]: return f"dataType {kctx['dataType']} vs {octx[4]}"
            ok2, back, _ = K.transform_inverse(tr, ref, {"blockSize": max(len(d), 1024), "flags": 0}, dst_cap=len(d) + 512)
            if not ok2 or back != d: return f"inverse ok={ok2} diff at {fd(back, d)} len {len(back)} vs {len(d)}"
        check(f"transform {tr} {name}[{len(d)}]", f)

if not only or "STREAM" in only:
    d = (synth.text(1_300_000, 3).tobytes() + synth.noise(200_000, 4).tobytes() + bytes(70000) + synth.exe_like(500_007, 5).tobytes() + b"tail!")
    cfgs = [(["NONE"], "HUFFMAN", 65536), (["LZ"], "ANS0", 1 << 20), (["LZX"], "HUFFMAN", 1 << 18), (["NONE"], "NONE", 1 << 16), (["LZ"], "NONE", 1 << 18),
            (["BWT", "RANK", "ZRLT"], "ANS1", 1 << 20), (["BWT", "SRT", "ZRLT"], "FPAQ", 1 << 20), (["ROLZ"], "ANS0", 1 << 20), (["ZRLT"], "ANS0", 1 << 16)]
    for tr, ent, bs in cfgs:
        for flags in (1, 0):
            if flags == 0 and "BWT" not in tr: continue
            def f():
                ref = O.compress(d, tr, ent, bs, bwt_bounds=flags)
                got = K.compress(d, tr, ent, bs, flags=flags)
                if got != ref: return f"stream differs at byte {fd(got, ref)} len {len(got)} vs {len(ref)}"
                back = K.decompress(ref, len(d) + 1024, flags=flags)
                if back != d: return f"decompress differs at {fd(back, d)} len {len(back)} vs {len(d)}"
            check(f"stream {'+'.join(tr)}&{ent} bs={bs} flags={flags}", f)
print("FAILURES:", fails)
