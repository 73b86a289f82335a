"""GPU: the kernels' re-formulated discriminator arithmetic (csrc/fm_math.cuh "v2": branch-free fast_atan2f,
FP64-free PCM scaling with a guard band) against literal device transcriptions of multifm/fast_atan2f.c:101-174 and
multifm/fm_demod.c:68-72 -- 2^31 pseudo-random operand pairs per FMA variant, and EVERY float in [-3.2, 3.2]
through the PCM scaling.  Bit-exact or fail."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(pkg, what, first, count, fma):
    out = (C.c_uint64 * 8)()
    rc = pkg._lib.lib().gpuchan_math_selftest(what, first, count, fma, out)
    assert rc == 0, pkg._lib.lib().gpuchan_last_error()
    return list(out)


@pytest.mark.parametrize("fma", [1, 0])
def test_fast_atan2f_v2_matches_the_literal_transcription(pkg, fma):
    out = _run(pkg, 0, 20260925 + fma, 1 << 31, fma)
    assert out[0] == 0, f"arctangent differs: first offender s_im={np.int32(np.uint32(out[4]))} s_re={np.int32(np.uint32(out[5]))} " \
                        f"ref={out[6]:#x} v2={out[7]:#x} ({out[0]} in total)"
    assert out[1] == 0, f"{out[1]} PCM values differ from the FP64 expression"
    assert out[2] > 0           # the guard band does trigger at this sample size ...
    assert out[2] < (1 << 31) // 100000  # ... but rarely


@pytest.mark.parametrize("first", [0x00000000, 0x80000000])
def test_pcm_scaling_exhaustive(pkg, first):
    """All float bit patterns of one sign with |phi| <= 3.2 (1.08e9 values each)."""
    count = 0x404CCCCE
    out = _run(pkg, 1, first, count, 1)
    assert out[1] == 0, f"{out[1]} PCM values differ; first: phi bits {out[4]:#x} ref {np.int32(np.uint32(out[5]))} got {np.int32(np.uint32(out[6]))}"
    assert 0 < out[2] < count // 100000


@pytest.mark.parametrize("fma", [1, 0])
def test_kernel_arithmetic_against_the_reference_fixture_and_the_oracle(pkg, oracle, fma):
    """The fused kernel's packed-pair arithmetic against what the REFERENCE objects computed (tests/golden/atan2.npz,
    recorded from multifm/fast_atan2f.c in both contraction variants) and against the oracle's discriminator
    (multifm/fm_demod.c:66-72) -- not only against our own device transcription."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "atan2.npz"))
    rng = np.random.default_rng(7)
    extra_im = rng.integers(-2**31, 2**31, 200000, dtype=np.int64)
    extra_re = rng.integers(-2**31, 2**31, 200000, dtype=np.int64) >> rng.integers(0, 31, 200000)
    s_im = np.concatenate([z["s_im"], extra_im]).astype(np.int32)
    s_re = np.concatenate([z["s_re"], extra_re]).astype(np.int32)
    n = len(s_im)
    phi = np.zeros(n, np.float32)
    pcm = np.zeros(n, np.int16)
    rc = pkg._lib.lib().gpuchan_math_eval(s_im.ctypes.data, s_re.ctypes.data, n, fma, phi.ctypes.data, pcm.ctypes.data)
    assert rc == 0, pkg._lib.lib().gpuchan_last_error()
    want = z["phi" if fma else "phi_nofma"]
    got = phi[:len(want)]
    same = (got.view(np.uint32) == want.view(np.uint32)) | ((got == 0) & (want == 0))
    assert same.all(), f"{(~same).sum()} of {len(want)} angles differ from the reference fixture"
    exp_phi = np.array([oracle.L.orc_fast_atan2f(np.float32(a), np.float32(b), fma) for a, b in zip(s_im[:20000], s_re[:20000])], np.float32)
    g = phi[:20000]
    assert ((g.view(np.uint32) == exp_phi.view(np.uint32)) | ((g == 0) & (exp_phi == 0))).all()
    exp_pcm = (phi.astype(np.float64) / np.pi * 16384.0).astype(np.float32).astype(np.int32).astype(np.int16)   # fm_demod.c:71-72
    assert np.array_equal(pcm, exp_pcm)
