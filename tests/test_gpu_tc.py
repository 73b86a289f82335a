"""GPU: the tensor-core (tcgen05 kind::i8) building block and engine."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _as(x, signed):
    return x.astype(np.int8).astype(np.int64) if signed else x.astype(np.uint8).astype(np.int64)


@pytest.mark.parametrize("Kp,R,N,s0,s1,signs", [(224, 66, 64, 0, 1, (1, 1, 0, 0)), (224, 66, 64, 2, 0, (1, 0, 0, 1)),
                                                (32, 64, 64, 0, 0, (0, 0, 1, 1)), (416, 70, 64, 5, 6, (1, 0, 1, 0)),
                                                (64, 40, 32, 7, 8, (0, 1, 1, 1))])
def test_tcgen05_i8_tile_conventions(pkg, Kp, R, N, s0, s1, signs):
    rng = np.random.default_rng(Kp + R)
    A0 = rng.integers(0, 256, (128, Kp), dtype=np.uint8)
    A1 = rng.integers(0, 256, (128, Kp), dtype=np.uint8)
    B0 = rng.integers(0, 256, (R, Kp), dtype=np.uint8)
    B1 = rng.integers(0, 256, (R, Kp), dtype=np.uint8)
    out = np.zeros((128, N), np.int32)
    rc = pkg._lib.lib().gpuchan_tc_selftest(A0.ctypes.data, B0.ctypes.data, A1.ctypes.data, B1.ctypes.data, Kp, R, N, s0, s1,
                                            *signs, out.ctypes.data)
    assert rc == 0
    exp = _as(A0, signs[0]) @ _as(B0[s0:s0 + N], signs[1]).T + _as(A1, signs[2]) @ _as(B1[s1:s1 + N], signs[3]).T
    exp = ((exp + 2**31) % 2**32 - 2**31).astype(np.int32)
    assert np.array_equal(out, exp)
