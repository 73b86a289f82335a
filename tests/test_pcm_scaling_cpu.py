"""CPU: the fused kernel's FP64-free PCM scaling (csrc/fm_math.cuh pcm_from_phi_pair: two-float product of six
multiply-adds, guard band as one multiply-add on the exponent of the rounded value, FP64 quotient by reciprocal with
exact-remainder correction behind it) restated with exact rational arithmetic and checked against the reference's
expression `(int16_t)(float)((double)phi / M_PI * 16384.0)` (multifm/fm_demod.c:71-72).

The exhaustive check -- every float in [-3.2, 3.2] through the device code -- is tests/test_gpu_math.py; this test pins
the *formulas* without a GPU: random angles, and angles constructed to fall next to float rounding boundaries of the
scaled value, where the guard band has to fire."""
from fractions import Fraction

import numpy as np


def rn(x: Fraction, p: int) -> Fraction:
    """x rounded to the nearest binary floating-point number with a p-bit significand (ties to even; normal range)."""
    if x == 0:
        return Fraction(0)
    s, a = (-1 if x < 0 else 1), abs(x)
    e = a.numerator.bit_length() - a.denominator.bit_length()       # 2^(e-1) <= a < 2^(e+1)
    if a < Fraction(2) ** e:
        e -= 1                                                       # now 2^e <= a < 2^(e+1)
    ulp = Fraction(2) ** (e - p + 1)
    q, r = divmod(a, ulp)
    q = int(q)
    if r * 2 > ulp or (r * 2 == ulp and (q & 1)):
        q += 1
    return s * q * ulp


def f32(x): return rn(x, 24)
def f64(x): return rn(x, 53)


C1 = Fraction(float(np.float32(0.3183098733425140380859375)))       # (float)(1 / M_PI)
C2 = Fraction(float(np.float32(1.2841276486597053e-08)))            # (float)(1 / M_PI - (double)c1)
PI_D = Fraction(3.14159265358979323846)                              # M_PI
Y_D = Fraction(0.31830988618379069122)                               # 1 / M_PI rounded to double
GUARD = Fraction(float(np.float32(1.9073486328125e-06)))
SCALE = f32(Fraction(2) ** -24 * (1 - GUARD))                        # PCM_GUARD_SCALE
assert SCALE == Fraction(2) ** -24 - Fraction(2) ** -43


def reference(phi: Fraction) -> int:
    """fm_demod.c:71-72 in IEEE arithmetic: double division, double multiplication, conversion to float, truncation."""
    q = f64(f64(phi / PI_D) * 16384)
    return int(f32(q))          # int() truncates toward zero


def fast(phi: Fraction):
    """pcm_from_phi_pair for one half: returns (trunc(f), margin)."""
    hi = f32(phi * (16384 * C1))
    nlo = f32(hi - phi * (16384 * C1))                               # fma(ph, -c1', hi): the product's rounding error, exact
    assert nlo == hi - phi * (16384 * C1)
    nlo = f32(nlo - phi * (16384 * C2))                              # fma(ph, -c2', nlo)
    f = f32(hi - nlo)                                                # fma(nlo, -1, hi)
    d = f32(f32(hi - f) - nlo)                                       # fma(nlo, -1, fma(f, -1, hi))
    if f == 0:
        p2 = Fraction(0)
    else:
        a = abs(f)
        e = a.numerator.bit_length() - a.denominator.bit_length()
        if a < Fraction(2) ** e:
            e -= 1
        p2 = Fraction(2) ** e                                        # bits(f) & 0x7f800000 as a float
    margin = f32(p2 * SCALE - abs(d))
    return int(f), margin


def exact_path(phi: Fraction) -> int:
    """pcm_from_phi_exact: a / M_PI correctly rounded in FP64 by reciprocal and exact-remainder correction."""
    ad = f32(phi * 16384)
    q = f64(ad * Y_D)
    r = f64(ad - q * PI_D)
    q = f64(q + r * Y_D)
    return int(f32(q))


def _angles():
    rng = np.random.default_rng(20261017)
    out = [Fraction(float(v)) for v in rng.uniform(-3.1416, 3.1416, 6000).astype(np.float32)]
    out += [Fraction(float(v)) for v in (rng.uniform(-1, 1, 1500) * 10.0 ** rng.uniform(-8, 0, 1500)).astype(np.float32)]
    # angles whose scaled value sits next to a float rounding boundary: midpoints between neighbouring floats
    for _ in range(2500):
        t = np.float32(rng.uniform(-16384, 16384))
        mid = (Fraction(float(t)) + Fraction(float(np.nextafter(t, np.float32(np.inf))))) / 2
        phi0 = np.float32(float(mid * PI_D / 16384))
        for k in (-1, 0, 1):
            v = phi0
            for _ in range(abs(k)):
                v = np.nextafter(v, np.float32(np.inf if k > 0 else -np.inf))
            out.append(Fraction(float(v)))
    out += [Fraction(0), Fraction(float(np.float32(np.pi))), Fraction(float(-np.float32(np.pi))), Fraction(float(np.float32(np.pi / 2)))]
    return out + _near_boundary_angles()


def _near_boundary_angles(limit=400):
    """Floats phi in [0.5, 3.2] whose scaled value lies within 2^-15 ulp of a float rounding boundary: found by scanning all
    2.2e7 of them in float64 (vectorised), then handed to the exact arithmetic.  The guard band (2^-20 ulp) sits inside."""
    lo, hi = np.float32(0.5).view(np.uint32), np.float32(3.2).view(np.uint32)
    found = []
    for start in range(int(lo), int(hi), 1 << 22):
        bits = np.arange(start, min(start + (1 << 22), int(hi)), dtype=np.uint32)
        phi = bits.view(np.float32).astype(np.float64)
        x = phi / np.pi * 16384.0
        f = x.astype(np.float32)
        half_ulp = np.spacing(np.abs(f)).astype(np.float64) / 2
        gap = np.abs(half_ulp - np.abs(x - f.astype(np.float64))) / half_ulp       # 0 = exactly on a rounding boundary
        idx = np.flatnonzero(gap < 2.0 ** -15)
        found += [(float(gap[i]), int(bits[i])) for i in idx]
    found.sort()
    out = []
    for _, b in found[:limit]:
        v = float(np.uint32(b).view(np.float32))
        out += [Fraction(v), Fraction(-v)]
    return out


def test_fast_pcm_scaling_equals_the_reference_expression_outside_the_guard_band():
    flagged = wrong_if_unguarded = 0
    angles = _angles()
    for phi in angles:
        want = reference(phi)
        got, margin = fast(phi)
        if margin < 0:
            flagged += 1
            wrong_if_unguarded += got != want
            assert exact_path(phi) == want, f"exact path differs at phi = {float(phi)!r}"
        else:
            assert got == want, f"fast path differs outside the guard band at phi = {float(phi)!r}: {got} vs {want}"
    # the band is narrow (2^-20 ulp), but the scan above finds the angles that fall into it; it must fire there
    assert 0 < flagged < len(angles) // 20, (flagged, len(angles))
    print(f"{len(angles)} angles, {flagged} in the guard band, {wrong_if_unguarded} of those would have been wrong without it")


def test_exact_path_equals_the_reference_expression_everywhere():
    for phi in _angles()[::3]:
        assert exact_path(phi) == reference(phi), float(phi)
