import sys, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for dbg in ("0", "2", "4", "6"):
    env = dict(os.environ, KZG_DEBUG=dbg)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_one.py"), "LZ", "text16385"], capture_output=True, text=True, env=env)
    print("KZG_DEBUG", dbg, r.stdout.strip().replace("\n", " | "), r.stderr[-200:])
