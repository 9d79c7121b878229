"""CPU tests of the product boundary: the C-ABI library loads, exports every symbol include/kzg.h declares,
answers the pure-host queries, and refuses compute loudly when no CUDA device exists (no CPU fallback)."""
import os
import re
import subprocess
import pytest
import kanzi_b200 as K
from kanzi_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "kzg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kzg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    out = subprocess.check_output(["nm", "-D", "--defined-only", binding.LIB_PATH], text=True)
    exported = set(re.findall(r" T (kzg_[a-z0-9_]+)", out))
    decl = declared_symbols()
    assert len(decl) >= 15
    missing = [s for s in decl if s not in exported]
    assert not missing, f"declared in kzg.h but not exported: {missing}"


def test_library_has_no_torch_or_oracle_dependency():
    out = subprocess.check_output(["ldd", binding.LIB_PATH], text=True)
    assert "torch" not in out and "kzoracle" not in out
    syms = subprocess.check_output(["nm", "-D", binding.LIB_PATH], text=True)
    assert "kzo_" not in syms      # nothing from oracle/ is linked into the product


def test_sm100a_cubin_present():
    out = subprocess.run(["cuobjdump", "-lelf", binding.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_max_encoded_len_matches_reference_formulas():
    # BWTBlockCodec.java:222-224, LZCodec.java:961-964, ROLZCodec.java:1001-1003, SRT.java:364-366
    for n in (1, 100, 512, 513, 1024, 1025, 65536, 4 << 20):
        assert K.transform_max_encoded_len("BWT", n) == n + 33
        assert K.transform_max_encoded_len("LZ", n) == (n + 16 if n <= 1024 else n + n // 64) + 2
        assert K.transform_max_encoded_len("LZX", n) == (n + 16 if n <= 1024 else n + n // 64) + 2
        assert K.transform_max_encoded_len("ROLZ", n) == (n + 64 if n <= 512 else n)
        assert K.transform_max_encoded_len("SRT", n) == n + 1024
        for t in ("RANK", "MTFT", "ZRLT", "NONE"):
            assert K.transform_max_encoded_len(t, n) == n
    # the sibling codecs: LZCodec.java:1283-1285 (LZP), RLT.java:355-357, ROLZCodec.java:1417-1421 (ROLZX); and the oracle agrees on all of them
    import oracle_lib as O
    for n in (1, 100, 512, 513, 1024, 1025, 16384, 16385, 65536, 4 << 20):
        assert K.transform_max_encoded_len("LZP", n) == (n + 16 if n <= 1024 else n + n // 64)
        assert K.transform_max_encoded_len("RLT", n) == (n + 32 if n <= 512 else n)
        assert K.transform_max_encoded_len("ROLZX", n) == (n + 1024 if n <= 16384 else n + n // 32)
        for t in K.T:
            assert K.transform_max_encoded_len(t, n) == O.lib().kzo_transform_max_encoded_len(O.T[t], n), (t, n)
    assert K.transform_max_encoded_len(2, 100) < 0          # BWTS: an id the library has no kernel for is refused, not guessed


def test_host_bitstreams_roundtrip():
    obs = K.OutputBitStream()
    obs.writeBits(0x4B414E5A, 32)
    obs.writeBits(5, 3)
    obs.writeBits(b"\xab\xcd\xef", 0, 20)
    obs.writeBit(1)
    n = obs.written()
    obs.close()
    ibs = K.InputBitStream(obs.toByteArray(), n)
    assert ibs.readBits(32) == 0x4B414E5A and ibs.readBits(3) == 5 and ibs.readBits(20) == 0xABCDE and ibs.readBit() == 1
    assert n == 56


@pytest.mark.skipif(K.device_count() > 0, reason="only meaningful without a GPU")
def test_compute_fails_loudly_without_gpu():
    with pytest.raises(K.KzgError) as e:
        K.entropy_encode("ANS0", b"x" * 1000)
    assert e.value.code == -binding.ERR_NO_DEVICE
    with pytest.raises(K.KzgError):
        K.compress(b"y" * 5000, ["LZ"], "ANS0", 1024)
    with pytest.raises(K.KzgError):
        K.transform_forward("LZ", b"z" * 5000)
