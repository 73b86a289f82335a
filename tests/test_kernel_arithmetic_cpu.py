"""CPU: the fused kernel's re-formulated discriminator arithmetic (csrc/fm_math.cuh) restated operation by operation in
exact rational arithmetic with IEEE rounding, and checked against the reference's semantics:

* the FP64-free PCM scaling (pcm_from_phi_pair: two-float product of six multiply-adds, guard band as one multiply-add
  on the exponent of the rounded value, FP64 quotient by reciprocal with exact-remainder correction behind it) against
  `(int16_t)(float)((double)phi / M_PI * 16384.0)` (multifm/fm_demod.c:71-72);
* the branch-free arctangent (atan2p_stage1..3: division by reciprocal and five multiply-adds, floor by the 2^23 trick,
  octant fix-up as multiply-adds on 0/1 compare results) against the oracle's transcription of
  multifm/fast_atan2f.c:101-174 -- itself pinned to the reference objects by tests/test_oracle_vs_ref.py.

The exhaustive checks -- every float in [-3.2, 3.2] and 2^31 operand pairs through the device code -- are
tests/test_gpu_math.py; these tests pin the *formulas* without a GPU."""
from fractions import Fraction

import numpy as np
import pytest


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


# ---- arctangent ---------------------------------------------------------------------------------------------------

def rz(x: Fraction, p: int) -> Fraction:
    """x rounded toward zero to a p-bit significand (add.rz)."""
    if x == 0:
        return Fraction(0)
    s, a = (-1 if x < 0 else 1), abs(x)
    e = a.numerator.bit_length() - a.denominator.bit_length()
    if a < Fraction(2) ** e:
        e -= 1
    ulp = Fraction(2) ** (e - p + 1)
    return s * int(a // ulp) * ulp


PI_F = Fraction(float(np.float32(3.14159265358979323846)))
HPI_F = Fraction(float(np.float32(1.57079632679489661923)))
TINY = Fraction(float(np.float32(1.0e-30)))
Z_THR = np.float32(0.003921569)
if float(Z_THR) < 0.003921569:
    Z_THR = np.nextafter(Z_THR, np.float32(np.inf))         # host_z_small_thr(): (double)z < 0.003921569 as a float compare
Z_THR = Fraction(float(Z_THR))


def atan2_v3(s_im: int, s_re: int, table, fma: bool, seed_ulps: int) -> Fraction:
    """atan2p_stage1..3 for one half of a pair.  seed_ulps perturbs the reciprocal seed (MUFU.RCP is good to about one ulp;
    the quotient must come out correctly rounded whatever the seed within that)."""
    fy, fx = f32(Fraction(s_im)), f32(Fraction(s_re))            # I2FP.F32.S32
    ya, xa = abs(fy), f32(abs(fx) + TINY)
    num, den = min(ya, xa), max(ya, xa)
    r0 = f32(1 / den)
    r0 += seed_ulps * (r0 - f32(r0 * (1 - Fraction(2) ** -24)) or Fraction(0))  # neighbouring floats of the exact reciprocal
    e = f32(1 - den * r0)
    r = f32(r0 + r0 * e)
    q = f32(num * r)
    rem = f32(num - den * q)
    z = f32(q + r * rem)
    assert z == f32(num / den), (s_im, s_re, seed_ulps)         # the div.rn fast path is exact for these operands
    alpha = f32(z * 255)
    t = rz(alpha + 2 ** 23, 24)
    idx = int(t - 2 ** 23) & 0xff
    frac = f32(alpha - (t - 2 ** 23))
    e_x = Fraction(float(table[idx]))
    e_y = f32(Fraction(float(table[idx + 1])) - e_x)             # the table's slope entry: float difference of neighbours
    ip = f32(e_y * frac + e_x) if fma else f32(e_x + f32(e_y * frac))
    base = z if z < Z_THR else ip
    sb = -base if s_re < 0 else base
    w01 = 1 if xa > ya else 0
    cnx = 1 if fx < -ya else 0
    w = f32(Fraction(w01 * 2 - 1))
    cst = f32(cnx * PI_F + f32(w01 * -HPI_F + HPI_F))
    inner = f32(sb * w + cst)
    return -inner if s_im < 0 else inner


@pytest.mark.parametrize("fma", [1, 0])
def test_branch_free_arctangent_equals_the_oracle(oracle, fma):
    table = oracle.atan_table()
    rng = np.random.default_rng(17 + fma)
    pairs = [(0, 0), (0, 1), (1, 0), (0, -1), (-1, 0), (1, 1), (-1, -1), (1, -1), (-1, 1), (2**31 - 1, 2**31 - 1),
             (-2**31, -2**31), (-2**31, 2**31 - 1), (2**31 - 1, 1), (1, 2**31 - 1), (255, 65025), (1, 255), (1, 254), (1, 256)]
    big = rng.integers(-2**31, 2**31, (1500, 2), dtype=np.int64)
    small = rng.integers(-2**31, 2**31, (1500, 2), dtype=np.int64) >> rng.integers(0, 31, (1500, 2))
    near_diag = rng.integers(-2**20, 2**20, 500, dtype=np.int64)
    pairs += [tuple(int(v) for v in p) for p in big] + [tuple(int(v) for v in p) for p in small]
    pairs += [(int(a), int(a + d)) for a, d in zip(near_diag, rng.integers(-2, 3, 500))]
    for n, (s_im, s_re) in enumerate(pairs):
        want = Fraction(float(oracle.L.orc_fast_atan2f(np.float32(s_im), np.float32(s_re), fma)))
        for seed in ((-1, 0, 1) if n % 8 == 0 else (0,)):
            got = atan2_v3(s_im, s_re, table, bool(fma), seed)
            assert got == want, (s_im, s_re, fma, seed, float(got), float(want))


# ---- integer part -------------------------------------------------------------------------------------------------

def _rq14_ref(a):
    """filter/complex.h:31-34 round_q30_q15 with "Q_15_SHIFT" = 14 (filter/filter.h:16) on a wrapping int32, truncated to int16."""
    a = a.astype(np.int32)
    return ((a >> 14) + ((a >> 13) & 1)).astype(np.int16).astype(np.int64)


def _wrap32(v):
    return ((v + 2**31) % 2**32 - 2**31).astype(np.int64)


def _top16(v):
    return _wrap32(v) >> 16


def test_integer_reformulations_equal_the_reference_rounding():
    """The epilogue folds the "<< 2" and the rounding constant of rq() into the multiply-adds that produce its argument
    (fm_math.cuh top16 / derotate_v2 / rot_step_v2 with the negated increment, tc_engine.cu comb): identities modulo 2^32,
    checked here on a million random operands including the int16 / int32 extremes."""
    rng = np.random.default_rng(5)
    n = 1 << 20
    i16 = lambda: np.concatenate([rng.integers(-32768, 32768, n - 4), [-32768, 32767, -32768, 32767]]).astype(np.int64)
    i32 = lambda: np.concatenate([rng.integers(-2**31, 2**31, n - 4), [-2**31, 2**31 - 1, 0, -1]]).astype(np.int64)
    # rq14(a) == (4a + 0x8000) >> 16 for every wrapping int32 a
    a = i32()
    assert np.array_equal(_top16(4 * a + 0x8000), _rq14_ref(_wrap32(a)))
    # limb recombination: SUM mode (a0 weight 2^8, a1 weight 1), RADIX mode (2^16, 2^8, 1)
    a0, a1, a2 = i32(), i32(), i32()
    assert np.array_equal(_top16(a1 * 4 + (a0 * 1024 + 0x8000)), _rq14_ref(_wrap32(a0 * 256 + a1)))
    assert np.array_equal(_top16(a2 * 4 + (a1 * 1024 + (a0 * 262144 + 0x8000))), _rq14_ref(_wrap32(a0 * 65536 + a1 * 256 + a2)))
    # derotation y = rq(q * rot) (direct_fir.c:162-163) and recurrence rot <- rq(rot * incr) (:166-167)
    q_re, q_im, r_re, r_im, i_re, i_im = i16(), i16(), i16(), i16(), i16(), i16()
    d_re, d_im = _wrap32(q_re * r_re - q_im * r_im), _wrap32(q_re * r_im + q_im * r_re)
    assert np.array_equal(_top16(d_re * 4 + 0x8000), _rq14_ref(d_re)) and np.array_equal(_top16(d_im * 4 + 0x8000), _rq14_ref(d_im))
    i4_re, i4_im, ni4_im = 4 * i_re, 4 * i_im, -4 * i_im
    n_re = r_re * i4_re + (r_im * ni4_im + 0x8000)
    n_im = r_re * i4_im + (r_im * i4_re + 0x8000)
    assert np.array_equal(_top16(n_re), _rq14_ref(_wrap32(r_re * i_re - r_im * i_im)))
    assert np.array_equal(_top16(n_im), _rq14_ref(_wrap32(r_re * i_im + r_im * i_re)))
