"""GPU parity: the CUDA channel bank (through the C ABI) against the CPU oracle, bit for bit.

Reference semantics under test: filter/direct_fir.c:329-417 (FIR + derotator), multifm/fm_demod.c:36-85 and
multifm/fast_atan2f.c:101-174 (discriminator), multifm/demod.c:205-261 (tap preparation).
north_star allows 1e-4 relative on FM audio; we hold the stronger bar: identical int16 PCM.
"""
import numpy as np
import pytest

from conftest import rand_iq
from tsl_sdr_b200 import synth
from tsl_sdr_b200.gpuchan import GpuChan, GpuChanError, F_ATAN_FMA, F_KEEP_IQ, ENGINE_IMAD, ENGINE_TC

pytestmark = pytest.mark.gpu


ENGINES = [pytest.param(ENGINE_IMAD, id="imad"), pytest.param(ENGINE_TC, id="tc")]


def run_bank(lpf, offs, fs, D, iq, chunks=None, gains=None, flags=F_ATAN_FMA | F_KEEP_IQ, max_batch=None, engine=ENGINE_IMAD):
    n = len(iq) // 2
    chunks = chunks or [n]
    try:
        bank = GpuChan(lpf, offs, fs, D, max_batch or max(chunks), gains=gains, flags=flags, engine=engine)
    except GpuChanError as exc:
        if engine == ENGINE_TC and exc.code == -5 and "tensor-core engine unavailable" in str(exc):
            pytest.skip(str(exc))
        raise
    assert bank.engine == engine
    pcm, yiq = [], []
    pos = 0
    for c in chunks:
        c = min(c, n - pos)
        if c <= 0:
            break
        bank.submit(iq[2 * pos: 2 * (pos + c)])
        pcm.append(bank.collect().copy())
        if flags & F_KEEP_IQ:
            yiq.append(bank.collect_iq().copy())
        pos += c
    launches = bank.kernel_launches
    assert launches > 0
    state = [bank.rot_state(c) for c in range(len(offs))]
    bank.close()
    pcm = np.concatenate(pcm, axis=1)
    yiq = np.concatenate(yiq, axis=1) if yiq else None
    return pcm, yiq, state


def oracle_bank(oracle, lpf, offs, fs, D, iq, gains=None, fma=1):
    outs_iq, outs_pcm = [], []
    for c, off in enumerate(offs):
        g = 1.0 if gains is None else gains[c]
        y, p = oracle.channel(lpf, off, fs, D, iq, gain=g, fma=fma)
        outs_iq.append(y.reshape(-1, 2))
        outs_pcm.append(p)
    return np.stack(outs_pcm), np.stack(outs_iq)


def assert_same(got_pcm, got_iq, exp_pcm, exp_iq):
    assert got_pcm.shape == exp_pcm.shape
    if got_iq is not None:
        bad = np.argwhere(got_iq != exp_iq)
        assert bad.size == 0, f"filtered IQ mismatch at {bad[:5].tolist()} (of {len(bad)})"
    bad = np.argwhere(got_pcm != exp_pcm)
    assert bad.size == 0, f"PCM mismatch at {bad[:5].tolist()} (of {len(bad)}): {got_pcm[tuple(bad[0])]} vs {exp_pcm[tuple(bad[0])]}"


@pytest.mark.parametrize("C,T,D,fs", [(1, 127, 100, 2400000), (5, 127, 25, 1200000), (33, 63, 16, 1000000),
                                      (64, 127, 100, 2400000), (70, 255, 200, 10000000), (3, 512, 120, 3000000),
                                      (40, 512, 120, 3000000), (2, 2, 1, 48000), (4, 33, 33, 250000)])
@pytest.mark.parametrize("engine", ENGINES)
def test_noise_one_shot(oracle, C, T, D, fs, engine):
    n = 40000 + 7 * D + 3
    iq = rand_iq(n, seed=C * 1000 + T)
    lpf = synth.lowpass_taps(T, min(9000.0, fs / 8), fs) if T > 2 else np.array([0.5, 0.5])
    offs = synth.channel_offsets(C, fs)
    pcm, yiq, _ = run_bank(lpf, offs, fs, D, iq, engine=engine)
    exp_pcm, exp_iq = oracle_bank(oracle, lpf, offs, fs, D, iq)
    assert_same(pcm, yiq, exp_pcm, exp_iq)


@pytest.mark.parametrize("engine", ENGINES)
def test_chunked_equals_one_shot(oracle, engine):
    """4096-sample sample_bufs (multifm/file_if.c:18), ragged tails and sub-T chunks give the same stream."""
    C, T, D, fs = 6, 127, 100, 2400000
    n = 50000
    iq = rand_iq(n, seed=11)
    lpf = synth.lowpass_taps(T, 9000.0, fs)
    offs = synth.channel_offsets(C, fs)
    exp_pcm, exp_iq = oracle_bank(oracle, lpf, offs, fs, D, iq)
    for chunks in ([4096] * 13, [1, 50, 126, 127, 128, 4096, 9999, 3, 100000], [7777] * 7):
        pcm, yiq, _ = run_bank(lpf, offs, fs, D, iq, chunks=chunks, max_batch=100000, engine=engine)
        assert_same(pcm, yiq, exp_pcm[:, :pcm.shape[1]], exp_iq[:, :pcm.shape[1]])
        assert pcm.shape[1] == exp_pcm.shape[1]


@pytest.mark.parametrize("engine", ENGINES)
def test_full_scale_and_gain_wraparound(oracle, engine):
    """int32 accumulators wrap, int16 truncation after rq: full-scale input with a +6 'dB' gain."""
    C, T, D, fs = 4, 127, 50, 2400000
    n = 20000
    rng = np.random.default_rng(5)
    iq = rng.choice(np.array([-32768, 32767, -32767, 0, 1, -1], dtype=np.int16), size=2 * n)
    lpf = synth.lowpass_taps(T, 200000.0, fs)
    offs = np.array([0, 600000, -600000, 1], dtype=np.int32)
    gains = np.array([1.0, 3.98, 15.0, 0.5])
    pcm, yiq, _ = run_bank(lpf, offs, fs, D, iq, gains=gains, engine=engine)
    exp_pcm, exp_iq = oracle_bank(oracle, lpf, offs, fs, D, iq, gains=gains)
    assert_same(pcm, yiq, exp_pcm, exp_iq)


@pytest.mark.parametrize("engine", ENGINES)
def test_nofma_variant(oracle, engine):
    C, T, D, fs = 3, 127, 100, 2400000
    iq = rand_iq(30000, seed=3)
    lpf = synth.lowpass_taps(T, 9000.0, fs)
    offs = synth.channel_offsets(C, fs)
    pcm, yiq, _ = run_bank(lpf, offs, fs, D, iq, flags=F_KEEP_IQ, engine=engine)
    exp_pcm, exp_iq = oracle_bank(oracle, lpf, offs, fs, D, iq, fma=0)
    assert_same(pcm, yiq, exp_pcm, exp_iq)


@pytest.mark.parametrize("engine", ENGINES)
def test_long_stream_rot_limit_cycle(oracle, engine):
    """Derotator transient -> limit cycle: the tabulated cycle must equal the recurrence
    after ~5e5 outputs; final rot compared with the oracle's state."""
    C, T, D, fs = 3, 16, 4, 2400000
    n = 2200000
    iq = rand_iq(n, seed=9, amp=8000)
    lpf = synth.lowpass_taps(T, 100000.0, fs)
    offs = np.array([312500, -320000, 25000], dtype=np.int32)
    pcm, yiq, state = run_bank(lpf, offs, fs, D, iq, chunks=[500000] * 5, flags=F_ATAN_FMA, engine=engine)
    for c, off in enumerate(offs):
        re, im = oracle.prepare_taps(lpf, off, fs)
        st = oracle.new_state(off, fs, D)
        _, p = oracle.chan_stream(st, re, im, D, iq, want_iq=False)
        assert np.array_equal(p, pcm[c])
        rot = np.frombuffer(bytes(st)[:4], dtype=np.int16)
        assert tuple(rot) == tuple(state[c][0]), (rot, state[c])
        assert state[c][4] != 0, "no limit cycle found"


@pytest.mark.parametrize("keep_iq", [False, True])
def test_steady_state_window_tables(oracle, keep_iq):
    """Tensor-core engine past every channel's transient: the eight derotator phases of a thread's block come from the
    window tables built from the tabulated limit cycles (filter/direct_fir.c:152-172 is the recurrence they replace).
    Cycle lengths 1, 2, 3, 4, 6, 8, 12, 15, 24, 200, 262, 356, 524, 600, 772, 1676, 1708 (odd, shorter than a block, not a
    multiple of 8), ragged submits so that a submit's first output is not a multiple of 8 outputs into the stream, and the
    launch count shows that the steady-state kernel (one launch per submit, no prepass) really ran."""
    C, T, D, fs = 17, 16, 4, 2400000
    offs = np.array([0, 300000, 150000, 75000, 25000, -320000, 50000, 100000, 200000, 7000, 9000, 13000, 59000, 71000, 79000,
                     29000, 23000], dtype=np.int32)
    chunks = [600001, 250003, 123457, 77777, 200001, 350013]
    n = sum(chunks)
    iq = rand_iq(n, seed=77, amp=9000)
    lpf = synth.lowpass_taps(T, 100000.0, fs)
    flags = F_ATAN_FMA | (F_KEEP_IQ if keep_iq else 0)
    bank = GpuChan(lpf, offs, fs, D, max(chunks), flags=flags, engine=ENGINE_TC)
    pcm, yiq, per_submit, pos = [], [], [], 0
    for c in chunks:
        before = bank.kernel_launches
        bank.submit(iq[2 * pos: 2 * (pos + c)])
        pcm.append(bank.collect().copy())
        if keep_iq:
            yiq.append(bank.collect_iq().copy())
        per_submit.append(bank.kernel_launches - before)
        pos += c
    cycles = [bank.rot_state(c) for c in range(C)]
    bank.close()
    assert per_submit[0] > 1 and per_submit[2:] == [1] * (len(chunks) - 2), per_submit
    assert len({int(st[4]) for st in cycles}) >= 12, "the offsets no longer give a variety of cycle lengths"
    pcm = np.concatenate(pcm, axis=1)
    for c, off in enumerate(offs):
        y, p = oracle.channel(lpf, off, fs, D, iq)
        bad = np.flatnonzero(p != pcm[c])
        assert bad.size == 0, f"channel {c} (offset {off}): first PCM mismatch at output {bad[0]} of {len(p)}"
        if keep_iq:
            assert np.array_equal(np.concatenate(yiq, axis=1)[c], y.reshape(-1, 2))


def test_steady_state_two_channel_groups_per_cta(oracle, pkg):
    """128 channels = two channel groups sharing every transformed sample tile (TcPlan::gpc = 2: each epilogue set owns one
    group), past every channel's derotator transient (phases from the cycle tables, one launch per submit), ragged submits.
    Every channel's PCM against the oracle (filter/direct_fir.c:329-417, multifm/fm_demod.c:36-85)."""
    from test_tc_plan_cpu import plan
    C, T, D, fs = 128, 16, 4, 2400000
    bad = {258750, 191250, 108750, 41250}          # offsets whose derotator transient is longer than the first submit
    offs = np.array([k * 3750 for k in range(-68, 76) if abs(k * 3750) not in bad][:C], dtype=np.int32)
    lpf = synth.lowpass_taps(T, 100000.0, fs)
    rc, info, _, _ = plan(pkg, lpf, offs, fs, D)
    assert rc == 0 and int(info[15]) >> 16 == 2, "this shape is meant to run with two groups per CTA"
    chunks = [500001, 123457, 77777, 200003, 99999]
    iq = rand_iq(sum(chunks), seed=78, amp=9000)
    bank = GpuChan(lpf, offs, fs, D, max(chunks), flags=F_ATAN_FMA, engine=ENGINE_TC)
    pcm, per_submit, pos = [], [], 0
    for c in chunks:
        before = bank.kernel_launches
        bank.submit(iq[2 * pos: 2 * (pos + c)])
        pcm.append(bank.collect().copy())
        per_submit.append(bank.kernel_launches - before)
        pos += c
    mus = [bank.rot_state(c)[3] for c in range(C)]
    bank.close()
    assert max(mus) < chunks[0] // D, "a transient outlasts the first submit"
    assert per_submit[0] > 1 and per_submit[1:] == [1] * (len(chunks) - 1), per_submit
    pcm = np.concatenate(pcm, axis=1)
    for c, off in enumerate(offs):
        _, p = oracle.channel(lpf, off, fs, D, iq)
        bad_at = np.flatnonzero(p != pcm[c])
        assert bad_at.size == 0, f"channel {c} (offset {off}): first PCM mismatch at output {bad_at[0]} of {len(p)}"


def test_taps_match_oracle(oracle, pkg):
    fs, T = 2400000, 127
    lpf = synth.lowpass_taps(T, 9000.0, fs)
    for off in [0, 1, -1, 312500, -320000, 1199999, -1200000, 25000]:
        for g in [1.0, 2.5118864315095806]:
            a = pkg.prepare_taps(lpf, off, fs, g)
            b = oracle.prepare_taps(lpf, off, fs, g)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert np.array_equal(pkg.derot_increment(off, fs, 100), oracle.derot_incr(off, fs, 100))


@pytest.mark.parametrize("engine", ENGINES)
def test_against_reference_objects(ref, oracle, engine):
    """Same input through the reference's own direct_fir + fm_demod (oracle/_ref) and the CUDA bank."""
    C, T, D, fs = 8, 127, 25, 1200000
    n = 4096 * 20
    iq = rand_iq(n, seed=21)
    lpf = synth.lowpass_taps(T, 9000.0, fs)
    offs = synth.channel_offsets(C, fs)
    pcm, yiq, _ = run_bank(lpf, offs, fs, D, iq, engine=engine)
    for c, off in enumerate(offs):
        r_iq, r_pcm = ref.channel(lpf, off, fs, D, iq)
        k = len(r_pcm)
        assert k > 0 and k <= pcm.shape[1]
        assert np.array_equal(r_pcm, pcm[c, :k])
        assert np.array_equal(r_iq.reshape(-1, 2), yiq[c, :k])


@pytest.mark.parametrize("C,T,D,fs,cut,gain", [
    (64, 127, 100, 2400000, 9000.0, 1.0),       # benchmark shape
    (64, 127, 100, 2400000, 4000.0, 1.0),       # narrow filter: every tap fits int8 -> single-limb path
    (64, 127, 100, 2400000, 9000.0, 7.5),       # gain pushes taps into the two-limb path
    (130, 127, 100, 2400000, 9000.0, 1.0),      # three channel groups, the last one partially filled
    (256, 127, 25, 1200000, 9000.0, 2.0),       # Q = 6 block rows, odd decimation (row padding)
    (64, 96, 96, 2400000, 20000.0, 3.0),        # Q = 1, Kp = 192
    (40, 255, 200, 10000000, 9000.0, 1.0),      # long rows (Kp = 416)
])
def test_tensor_core_engine_shapes(oracle, C, T, D, fs, cut, gain):
    n = 63 * D * 9 + T + 5 * D + 17             # several tiles, ragged last tile
    iq = rand_iq(n, seed=C + T + D, amp=9000.0)
    lpf = synth.lowpass_taps(T, cut, fs)
    offs = synth.channel_offsets(C, fs)
    gains = np.full(C, gain)
    pcm, yiq, _ = run_bank(lpf, offs, fs, D, iq, gains=gains, engine=ENGINE_TC, chunks=[n // 3, n - n // 3], max_batch=n)
    sel = sorted(set([0, 1, C // 2, C - 2, C - 1, 63 % C, 64 % C]))
    for c in sel:
        y, p = oracle.channel(lpf, offs[c], fs, D, iq, gain=gain)
        assert pcm.shape[1] == len(p)
        assert np.array_equal(yiq[c], y.reshape(-1, 2)), f"channel {c} IQ"
        assert np.array_equal(pcm[c], p), f"channel {c} PCM"


def test_auto_engine_picks_tensor_cores_whenever_the_plan_fits():
    """AUTO = tensor-core engine for any channel count (measured 5x faster than the IMAD engine even for one channel);
    the IMAD engine remains for shapes whose tap image and sample ring do not fit shared memory."""
    fs, T, D = 2400000, 127, 100
    lpf = synth.lowpass_taps(T, 9000.0, fs)
    a = GpuChan(lpf, synth.channel_offsets(64, fs), fs, D, 1 << 16)
    b = GpuChan(lpf, synth.channel_offsets(3, fs), fs, D, 1 << 16)
    assert a.engine == ENGINE_TC and b.engine == ENGINE_TC
    a.close(); b.close()
    fs, T, D = 3000000, 2047, 2000          # tap image alone exceeds shared memory -> IMAD fallback, or a clean refusal
    try:
        c = GpuChan(synth.lowpass_taps(T, 9000.0, fs), synth.channel_offsets(3, fs), fs, D, 1 << 16)
        assert c.engine == ENGINE_IMAD
        c.close()
    except GpuChanError as exc:
        assert exc.code == -5


@pytest.mark.parametrize("fmt_name", ["cs8", "cu8", "cu8_rtl"])
def test_8bit_formats_widened_on_device(oracle, fmt_name):
    """f3: 8-bit captures cross PCIe as bytes and are widened on the device with the reference's conversions
    (multifm/file_if.c:67-157, multifm/rtl_sdr_if.c:142-147, including the cu8 int8_t-pointer quirk)."""
    from tsl_sdr_b200.gpuchan import FMT_CS8, FMT_CU8, FMT_CU8_RTL
    C, T, D, fs = 40, 127, 50, 2400000
    n = 60000
    raw = np.random.default_rng(21).integers(0, 256, 2 * n, dtype=np.uint8)
    if fmt_name == "cs8":
        fmt, wide = FMT_CS8, raw.view(np.int8).astype(np.int16)
    elif fmt_name == "cu8":
        fmt, wide = FMT_CU8, (raw.view(np.int8).astype(np.int16) - 127).astype(np.int16)
    else:
        fmt, wide = FMT_CU8_RTL, ((raw.astype(np.int16) - 127) << 7).astype(np.int16)
    lpf = synth.lowpass_taps(T, 60000.0, fs)
    offs = synth.channel_offsets(C, fs)
    gains = np.full(C, 8.0)
    exp_pcm, _ = oracle_bank(oracle, lpf, offs, fs, D, wide, gains=gains)
    for engine in (ENGINE_IMAD, ENGINE_TC):
        bank = GpuChan(lpf, offs, fs, D, 25000, gains=gains, flags=F_ATAN_FMA, engine=engine)
        got, pos = [], 0
        for k in (25000, 4096, 1, 25000, 5903):
            bank.submit_bytes(raw[2 * pos: 2 * (pos + k)], fmt)
            got.append(bank.collect().copy())
            pos += k
        bank.close()
        assert pos == n
        got = np.concatenate(got, axis=1)
        assert got.shape == exp_pcm.shape and np.array_equal(got, exp_pcm)


@pytest.mark.parametrize("C,T,D,fs,log2n", [(64, 127, 100, 2400000, 25),        # BASELINE configs[1] at bench.py's batch size
                                            (256, 127, 25, 1200000, 22),         # configs[2] shape
                                            (1024, 255, 200, 10000000, 23),      # configs[3] shape, all channels on one GPU
                                            (256, 512, 120, 3000000, 22)])       # configs[4] shape
def test_full_size_properties(oracle, C, T, D, fs, log2n):
    """Full-size runs, checked through size-independent properties: (1) the tensor-core engine and the int32 CUDA-core
    engine -- two independent implementations -- give the same PCM, bit for bit, for every channel; (2) chunk
    invariance: one submit == three ragged submits; (3) prefix property: the outputs that depend only on the first 2^20
    input samples equal the CPU oracle's for a handful of channels."""
    n = 1 << log2n
    iq = rand_iq(n, seed=log2n * 7 + C)
    lpf = synth.lowpass_taps(T, min(9000.0, fs / 8), fs)
    offs = synth.channel_offsets(C, fs)
    pcm_tc, _, _ = run_bank(lpf, offs, fs, D, iq, flags=F_ATAN_FMA, engine=ENGINE_TC)
    pcm_im, _, _ = run_bank(lpf, offs, fs, D, iq, flags=F_ATAN_FMA, engine=ENGINE_IMAD)
    assert pcm_tc.shape == (C, (n - T) // D + 1)
    assert np.array_equal(pcm_tc, pcm_im), "tensor-core and IMAD engines disagree"
    cuts = [n // 3 + 17, n // 2 + 4099, n]
    chunks = [cuts[0], cuts[1] - cuts[0], cuts[2] - cuts[1]]
    pcm_ck, _, _ = run_bank(lpf, offs, fs, D, iq, chunks=chunks, flags=F_ATAN_FMA, max_batch=max(chunks), engine=ENGINE_TC)
    assert np.array_equal(pcm_ck, pcm_tc), "chunked stream differs from the one-shot stream"
    m = 1 << 20
    k = (m - T) // D + 1
    for c in sorted({0, C // 3, C - 1}):
        _, exp = oracle.channel(lpf, offs[c], fs, D, iq[:2 * m])
        assert np.array_equal(pcm_tc[c, :k], exp[:k]), f"channel {c}: prefix differs from the oracle"
