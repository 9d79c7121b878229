"""Opcode histogram of the named kernels in the built library (cuobjdump -sass), for profiles/: how many instructions, which memory
instructions at which widths, and whether the asynchronous-copy / TMA mnemonics (LDGSTS, UBLKCP, SYNCS, UTMALDG) are present.
Usage: python tools/sass_summary.py kernel_substring ... > profiles/rNN_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "kanzi_b200", "libkanzi_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append(m.group(1).strip())
print(f"cuobjdump -sass {os.path.relpath(lib, ROOT)} (sm_100a), kernels matching: {' '.join(sys.argv[1:])}\n")
for want in sys.argv[1:]:
    for name, ins in funcs.items():
        if want not in name:
            continue
        ops = collections.Counter()
        for i in ins:
            i = re.sub(r"^@!?U?P\d+\s+", "", i)
            ops[i.split()[0]] += 1
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        print(f"{dem.split('(')[0]}: {len(ins)} SASS instructions")
        mem = {k: v for k, v in ops.items() if re.match(r"(LDG|STG|LD\.|ST\.|LDS|STS|LDL|STL|ATOM|RED|LDGSTS|UBLKCP|UTMA|SYNCS|LDSM|SHFL|VOTE|MATCH|BAR|WARPSYNC|REDUX)", k)}
        print("  memory / sync / warp ops: " + ", ".join(f"{k} {v}" for k, v in sorted(mem.items(), key=lambda kv: -kv[1])))
        top = [f"{k} {v}" for k, v in ops.most_common(14) if k not in mem]
        print("  most frequent others:     " + ", ".join(top))
        print("  async-copy / TMA mnemonics: " + (", ".join(f"{k} {v}" for k, v in ops.items() if re.match(r"(LDGSTS|UBLKCP|UTMA|SYNCS)", k)) or "none"))
        print()
