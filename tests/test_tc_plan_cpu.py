"""CPU: host logic of the tensor-core engine, without a GPU.

The engine computes the reference's complex int16 FIR (filter/direct_fir.c:366-385, int32 accumulators that wrap) as a
program of int8 x int8 -> int32 tile products over a tap image and two byte planes of the samples.  Everything the
host prepares for that -- limb decomposition of the taps (sum of int8 terms, or radix 256), image layout, operand
offsets, operand signedness, accumulator assignment, the split of the program over the two issuing warps -- is
checked here by executing the program with numpy on random samples and comparing the recombined accumulators with the
direct integer FIR, bit for bit.  (What the GPU adds -- tcgen05 itself -- is covered by tests/test_gpu_tc.py.)
"""
import ctypes as C

import numpy as np
import pytest

from tsl_sdr_b200 import synth

TC_OUT, TC_LEAD, TC_N, TC_CH = 64, 16, 80, 64


def plan(pkg, lpf, offs, fs, D, gains=None, smem=0):
    L = pkg._lib.lib()
    lpf = np.ascontiguousarray(lpf, np.float64)
    offs = np.ascontiguousarray(offs, np.int32)
    cfg = pkg._lib.GpuChanCfg()
    cfg.struct_size = C.sizeof(pkg._lib.GpuChanCfg)
    cfg.sample_rate_hz, cfg.decimation, cfg.nr_taps, cfg.nr_channels = int(fs), int(D), len(lpf), len(offs)
    cfg.max_batch_samples = 1 << 16
    cfg.lpf_taps = lpf.ctypes.data_as(C.POINTER(C.c_double))
    cfg.offset_hz = offs.ctypes.data_as(C.POINTER(C.c_int32))
    g = None
    if gains is not None:
        g = np.ascontiguousarray(gains, np.float64)
        cfg.gain = g.ctypes.data_as(C.POINTER(C.c_double))
    info = (C.c_uint32 * 16)()
    rc = L.gpuchan_tc_plan_query(C.byref(cfg), smem, info, None, 0, None, 0)
    if rc != 0:
        return rc, list(info), None, None
    img = np.zeros((info[15] & 0xffff) * info[6], np.uint8)
    prog = np.zeros((info[10], 4), np.uint32)
    rc = L.gpuchan_tc_plan_query(C.byref(cfg), smem, info, img.ctypes.data, img.size, prog.ctypes.data, len(prog))
    assert rc == 0
    return 0, list(info), img, prog


def emulate_tile(info, img_group, prog, x):
    """x: int16 [R, D, 2] (rows of D complex samples).  Returns int32 [128, TC_N]: the recombined accumulators."""
    _, mode, accs, _, _, a_chunks, a_group_bytes, _, _, _, _, _, Kp, Q, R, _ = info
    D = x.shape[1]
    raw = np.zeros((R, Kp), np.int16)
    raw[:, :2 * D] = x.reshape(R, 2 * D)
    hi = (raw >> 8).astype(np.int64)                    # signed high byte
    lo = (raw & 0xff).astype(np.int64)                  # unsigned low byte
    planes = {False: hi, True: lo}
    A = img_group.reshape(a_chunks, 2, 128, 16)         # [chunk][K half][row][16]
    acc = np.zeros((accs, 128, TC_N), np.int64)
    started = [False] * accs
    nslab = Kp // 16
    plane_lo16 = nslab * R
    for a_lo, b_lo, d_acc, idesc in prog:
        a_off16 = int(a_lo) & 0xffff
        assert (int(a_lo) >> 16) == 2048 // 16 and (int(b_lo) >> 16) == R       # LBO fields
        chunk = a_off16 // 256
        b_off16 = int(b_lo) & 0xffff
        is_lo = b_off16 >= plane_lo16
        rem = b_off16 - (plane_lo16 if is_lo else 0)
        slab, q = divmod(rem, R)
        assert slab % 2 == 0 and q < Q
        a_signed, b_signed = (int(idesc) >> 7) & 1, (int(idesc) >> 10) & 1
        assert ((int(idesc) >> 17) & 0x3f) == TC_N // 8 and ((int(idesc) >> 24) & 0x1f) == 128 // 16
        Am = np.concatenate([A[chunk, 0], A[chunk, 1]], axis=1)                 # [128][32]
        Am = Am.view(np.int8).astype(np.int64) if a_signed else Am.astype(np.int64)
        P = planes[is_lo]
        assert b_signed == (0 if is_lo else 1)          # high bytes are signed, low bytes unsigned
        Bm = P[q:q + TC_N, slab * 16:(slab + 2) * 16]   # [N][32]
        which = int(d_acc) & 0xffff
        assert which % TC_N == 0
        w = which // TC_N
        accumulate = int(d_acc) >> 31
        assert accumulate == (1 if started[w] else 0), "first MMA into an accumulator must not accumulate"
        started[w] = True
        acc[w] += Am @ Bm.T
    assert all(started)
    if mode == 0:
        tot = acc[0] * 256 + acc[1]
    else:
        tot = acc[0] * 65536 + acc[1] * 256 + acc[2]
    return ((tot + 2**31) % 2**32 - 2**31).astype(np.int64)


SHAPES = [(64, 127, 100, 2400000, None), (5, 127, 25, 1200000, None), (70, 255, 200, 10000000, None),
          (40, 512, 120, 3000000, None), (33, 63, 16, 1000000, None), (4, 33, 33, 250000, None),
          (6, 127, 100, 2400000, [2.5, 3.98, 1.0, 0.5, 2.0, 1.5]),     # sum of int8 terms, several terms
          (4, 127, 50, 2400000, [1.0, 3.98, 15.0, 40.0])]              # radix mode (|tap| > 508)


@pytest.mark.parametrize("Cn,T,D,fs,gains", SHAPES)
def test_mma_program_equals_direct_integer_fir(pkg, Cn, T, D, fs, gains):
    cut = 200000.0 if gains is not None and max(gains) > 10 else min(9000.0, fs / 8)
    lpf = synth.lowpass_taps(T, cut, fs)
    offs = synth.channel_offsets(Cn, fs)
    rc, info, img, prog = plan(pkg, lpf, offs, fs, D, gains)
    assert rc == 0 and info[0] == 1
    ok, mode, accs, nb, nt, a_chunks, a_group_bytes, b_stage_bytes, smem, copies, prog_len, split, Kp, Q, R, G = info
    G, gpc = G & 0xffff, G >> 16
    assert gpc in (1, 2) and (gpc == 1 or (G % 2 == 0 and nb >= 2 and copies == 16))
    assert smem == gpc * a_group_bytes + nb * b_stage_bytes + 2048 * copies + 128
    maxabs = max(int(np.abs(v).max()) for c in range(Cn)
                 for v in pkg.prepare_taps(lpf, offs[c], fs, 1.0 if gains is None else gains[c]))
    assert mode == (0 if maxabs <= 4 * 127 else 1)      # sum of at most four int8 terms, else radix 256
    assert R == TC_N + Q - 1 and Q == -(-T // D) and Kp % 32 == 0 and Kp >= 2 * D and G == -(-Cn // TC_CH)
    assert 2 <= nb <= 4 and nt == 512 // (accs * TC_N) and smem <= 232448 - 3072 and copies in (1, 16)
    assert 0 < split < prog_len <= 160
    # the two issuing warps own disjoint accumulators
    accs_of = [set((int(d) & 0xffff) // TC_N for d in prog[:split, 2]), set((int(d) & 0xffff) // TC_N for d in prog[split:, 2])]
    assert not (accs_of[0] & accs_of[1]) and len(accs_of[0] | accs_of[1]) == accs
    rng = np.random.default_rng(T * 7 + D)
    x = rng.integers(-32768, 32768, (R, D, 2), dtype=np.int64).astype(np.int16)
    x[rng.integers(0, R, 8), rng.integers(0, D, 8)] = [-32768, 32767]           # extremes
    stream = x.reshape(R * D, 2).astype(np.int64)
    for g in range(G):
        got = emulate_tile(info, img[g * a_group_bytes:(g + 1) * a_group_bytes], prog, x)
        for ch in sorted({0, 1, 15, 16, 47, 63} | {(Cn - 1) % TC_CH}):
            c = g * TC_CH + ch
            s, i = divmod(ch, 16)
            row_re, row_im = 32 * s + i, 32 * s + 16 + i
            if c >= Cn:
                assert not got[row_re].any() and not got[row_im].any()      # padding channels: zero taps
                continue
            c_re, c_im = (v.astype(np.int64) for v in pkg.prepare_taps(lpf, offs[c], fs, 1.0 if gains is None else gains[c]))
            for n in (0, 1, TC_LEAD - 1, TC_LEAD, 40, TC_N - 1):
                seg = stream[n * D:n * D + T]
                e_re = int((c_re * seg[:, 0] - c_im * seg[:, 1]).sum())
                e_im = int((c_im * seg[:, 0] + c_re * seg[:, 1]).sum())
                wrap = lambda v: (v + 2**31) % 2**32 - 2**31
                assert got[row_re, n] == wrap(e_re) and got[row_im, n] == wrap(e_im), (c, n)


def test_plan_limits(pkg):
    """Shapes whose tap image and two sample stages cannot share 227 KB are refused with a reason (the bank then falls
    back to the IMAD engine); a smaller shared-memory budget trades arctangent-table copies and stages."""
    fs = 3000000
    rc, info, _, _ = plan(pkg, synth.lowpass_taps(2047, 9000.0, fs), synth.channel_offsets(3, fs), fs, 2000)
    assert rc == -5 and info[0] == 0
    assert b"tensor-core engine unavailable" in pkg._lib.lib().gpuchan_last_error()
    lpf, offs = synth.lowpass_taps(127, 9000.0, 2400000), synth.channel_offsets(64, 2400000)
    rc, full, _, _ = plan(pkg, lpf, offs, 2400000, 100)
    rc2, tight, _, _ = plan(pkg, lpf, offs, 2400000, 100, smem=140000)
    assert rc == 0 and rc2 == 0
    assert full[9] == 16 and full[3] >= 3          # 16 conflict-free table copies and at least 3 stages on a B200
    assert tight[9] == 1 and 2 <= tight[3] <= full[3] and tight[8] <= 140000


def test_two_channel_groups_per_cta_when_both_tap_images_fit(pkg, monkeypatch):
    """TcPlan::gpc (csrc/tc_engine.cu tc_make_plan): a transformed sample tile feeds two channel groups' MMAs whenever both
    tap images, two sample stages and the 16 arctangent copies fit the 227 KB; one group per CTA otherwise (odd group
    counts, long filters) and under the measurement knob GPUCHAN_TC_GPC=1."""
    def gpc_of(Cn, T, D, fs):
        rc, info, _, _ = plan(pkg, synth.lowpass_taps(T, min(9000.0, fs / 8), fs), synth.channel_offsets(Cn, fs), fs, D)
        assert rc == 0 and info[0] == 1
        return int(info[15]) >> 16, int(info[3]), int(info[9]), int(info[8])
    monkeypatch.delenv("GPUCHAN_TC_GPC", raising=False)
    gpc, nb, copies, smem = gpc_of(256, 127, 100, 2400000)          # north_star's shape
    assert (gpc, nb, copies) == (2, 2, 16) and smem <= 232448 - 3072
    assert gpc_of(256, 127, 25, 1200000)[0] == 2                     # configs[2]
    assert gpc_of(64, 127, 100, 2400000)[0] == 1                     # one group
    assert gpc_of(192, 127, 100, 2400000)[0] == 1                    # three groups
    assert gpc_of(1024, 255, 200, 10000000)[0] == 1                  # configs[3]: two 100 KB tap images do not fit
    assert gpc_of(256, 512, 120, 3000000)[0] == 1                    # configs[4]
    monkeypatch.setenv("GPUCHAN_TC_GPC", "1")
    assert gpc_of(256, 127, 100, 2400000)[0] == 1
