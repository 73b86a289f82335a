"""CPU: the reference's per-object interfaces (include/compat: multifm/fm_demod.h:22-34, pager/pager_pocsag.h:8-56,
pager/pager_flex.h:16-115) compile and link from plain C the way decoder/decoder.c:685-697 and multifm/demod.c:89 use
them, the compat library exports exactly the reference's entry points, and without a GPU every constructor fails loudly
(there is no CPU path behind these names)."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "tsl-sdr_b200")
LIB = os.path.join(PKG, "libtslb200_compat.so")

REFERENCE_ENTRY_POINTS = ["multifm_fm_demod_init", "multifm_fm_demod_process", "multifm_fm_demod_cleanup",
                          "pager_pocsag_new", "pager_pocsag_delete", "pager_pocsag_on_pcm",
                          "pager_flex_new", "pager_flex_delete", "pager_flex_on_pcm"]


def build_compat_decoder(dst):
    exe = os.path.join(dst, "compat_decoder")
    subprocess.run(["gcc", "-std=gnu11", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include", "compat"), "-o", exe,
                    os.path.join(ROOT, "tests", "c", "compat_decoder.c"), "-L", PKG, "-ltslb200_compat", "-ltslb200",
                    f"-Wl,-rpath,{PKG}"], check=True)
    return exe


@pytest.mark.skipif(not os.path.exists(LIB), reason="compat library not built")
def test_compat_library_exports_the_reference_entry_points():
    L = ctypes.CDLL(LIB)
    for name in REFERENCE_ENTRY_POINTS:
        assert hasattr(L, name), name
    # every prototype in the compat headers is one of them (nothing declared that is not exported)
    for hdr in ("fm_demod.h", "pager_pocsag.h", "pager_flex.h"):
        text = open(os.path.join(ROOT, "include", "compat", hdr)).read()
        for line in text.splitlines():
            if line.startswith("aresult_t ") and "(" in line and "(*" not in line:
                assert line.split()[1].split("(")[0] in REFERENCE_ENTRY_POINTS, line


@pytest.mark.skipif(not os.path.exists(LIB), reason="compat library not built")
def test_reference_call_sequence_compiles_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe = build_compat_decoder(str(tmp_path))
    if torch.cuda.is_available():
        pytest.skip("GPU present: parity is covered by tests/test_gpu_compat.py")
    (tmp_path / "in.bin").write_bytes(os.urandom(8192))
    for mode in ("POCSAG", "FLEX", "FM"):
        r = subprocess.run([exe, mode, str(tmp_path / "in.bin"), str(tmp_path / "out.txt")], capture_output=True, text=True, timeout=60)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr and "failed: -5" in r.stderr, (mode, r.stderr)
        assert (tmp_path / "out.txt").read_bytes() == b""
