"""bench.py's reference arm (the CPU restatement timed on the host cores) runs without a GPU: its JSON line must carry the keys the
driver reads and the SAME `config` dict the GPU arm prints for the same arguments (the driver compares the two)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_line(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


def test_reference_arm_line_contract():
    d = _ref_line("--config", "cfg1", "--steps", "1", "--warmup", "0")
    assert d["impl"] == "reference" and d["metric"] == "encode+decode MB/s" and d["unit"] == "MB/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1 and d["dtype"] == "u8" and d["data"] == "synthetic"
    assert d["e2e"] == {"value": d["value"], "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "sample" in cb and cb["value"] == d["value"]
    assert d["gpu_launches"] == 0


def test_reference_arm_config_matches_gpu_arm():
    sys.path.insert(0, ROOT)
    import bench
    from kanzi_b200 import synth
    gen, full, transforms, entropy, bs = synth.CONFIGS["cfg1"]
    want = bench.config_dict("cfg1", transforms, entropy, bs, max(bs, int(full * bench.DEFAULT_SCALE["cfg1"])), gen.__name__, 1, 1)
    got = _ref_line("--config", "cfg1", "--steps", "1", "--warmup", "0")["config"]
    assert got == want
    assert set(want) == {"workload", "sharding", "l2", "bwt_bounds", "resident_decode_note"}
