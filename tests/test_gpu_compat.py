"""GPU: a C program written against the reference's own headers (include/compat) and call sequences
(decoder/decoder.c:635-651,685-697; multifm/demod.c:89) produces what the reference objects produced: the golden
fixtures recorded from them (tests/golden) and the oracle."""
import os
import subprocess

import numpy as np
import pytest

import flexcases
from test_compat_cpu import build_compat_decoder

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run(exe, mode, data, tmp_path, block=1024):
    data.tofile(tmp_path / "in.bin")
    r = subprocess.run([exe, mode, str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), str(block)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-1000:]
    return (tmp_path / "out.bin").read_bytes()


def test_fm_demod_object_matches_reference_fixture(tmp_path):
    exe = build_compat_decoder(str(tmp_path))
    z = np.load(os.path.join(G, "fir_fm.npz"))
    for key, block in (("c1_0", 1024), ("c4_0", 1024), ("c4_0", 7), ("c0_2", 100000)):
        got = np.frombuffer(run(exe, "FM", z[key + "_y"], tmp_path, block), dtype=np.int16)
        assert np.array_equal(got, z[key + "_pcm"]), key


def test_pager_pocsag_object_matches_reference_fixture(tmp_path):
    exe = build_compat_decoder(str(tmp_path))
    z = np.load(os.path.join(G, "pocsag_chain.npz"))
    total = 0
    for c in range(4):
        for block in (1024, 333):
            lines = run(exe, "POCSAG", z[f"ch{c}_res"], tmp_path, block).decode().splitlines()
            exp = []
            for me, tx in zip(z[f"ch{c}_meta"], z[f"ch{c}_text"]):
                kind, baud, cap, fn, ln = (int(v) for v in me)
                exp.append(f"POCSAG {'ALN' if kind else 'NUM'} {baud} {cap} {fn} {ln} {bytes(tx[:ln]).hex()}")
            assert lines == exp, (c, block)
            total += len(lines)
    assert total >= 8


@pytest.mark.parametrize("coding", ["1600/2", "6400/4"])
def test_pager_flex_object_matches_reference_fixture(tmp_path, coding):
    exe = build_compat_decoder(str(tmp_path))
    z = np.load(os.path.join(G, "flex.npz"))
    for t in (0, 3):
        key = f"{coding.replace('/', '_')}_t{t}"
        exp = []
        for kind, baud, cap, phase, ln, aux, text in flexcases.arrays_to_msgs(z[key + "_meta"], z[key + "_text"]):
            if kind == 2:
                exp.append(f"FLEX ALN {baud} {phase} {aux[0]} {aux[1]} {cap} {aux[2]} {aux[3]} {aux[4]} {ln} {text.hex()}")
            elif kind == 3:
                exp.append(f"FLEX NUM {baud} {phase} {aux[0]} {aux[1]} {cap} {ln} {text.hex()}")
            else:
                exp.append(f"FLEX SIV {baud} {phase} {aux[0]} {aux[1]} {cap} {aux[2]} {aux[3]}")
        lines = run(exe, "FLEX", flexcases.case_pcm(coding, t), tmp_path).decode().splitlines()
        assert len(exp) > 5 and lines == exp, key
