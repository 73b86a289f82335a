"""CPU, differential: oracle restatement vs the reference's own objects (oracle/_ref), live.
Skipped where oracle/_ref was never built (it needs /root/reference at build time)."""
import numpy as np
import pytest

from conftest import rand_iq
from tsl_sdr_b200 import synth


@pytest.mark.parametrize("T,D,fs,off", [(127, 100, 2400000, 312500), (127, 25, 1200000, -320000), (128, 40, 1000000, 99999),
                                        (255, 200, 10000000, -4000001), (512, 120, 3000000, 777777), (2, 1, 48000, 1000),
                                        (33, 33, 250000, -60000), (64, 7, 250000, 0)])
def test_channel_matches_reference(oracle, ref, T, D, fs, off):
    n = 4096 * 9
    iq = rand_iq(n, seed=T + D)
    lpf = synth.lowpass_taps(T, min(9000.0, fs / 8), fs) if T > 2 else np.array([0.5, 0.5])
    for gain in (1.0, 3.9810717055349722):
        r_iq, r_pcm, st = ref.channel(lpf, off, fs, D, iq, gain=gain, return_state=True)
        o_iq, o_pcm = oracle.channel(lpf, off, fs, D, iq, gain=gain)
        k = len(r_pcm)
        assert k > 0 and np.array_equal(r_iq, o_iq[:2 * k]) and np.array_equal(r_pcm, o_pcm[:k])
        a = ref.prepare_taps(lpf, off, fs, gain)
        b = oracle.prepare_taps(lpf, off, fs, gain)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert np.array_equal(st[1], oracle.derot_incr(off, fs, D))


def test_reference_chunking_is_stream_invariant(ref):
    """The 4096-in / 1024-out chunking of demod_thread_process does not change the stream."""
    T, D, fs, off = 127, 100, 2400000, 312500
    iq = rand_iq(4096 * 12, seed=5)
    lpf = synth.lowpass_taps(T, 9000.0, fs)
    a = ref.channel(lpf, off, fs, D, iq)
    b = ref.channel(lpf, off, fs, D, iq, chunks=[1, 4095, 5000, 3192, 4096 * 3, 100000])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_full_scale_wraparound(oracle, ref):
    T, D, fs = 127, 50, 2400000
    rng = np.random.default_rng(5)
    iq = rng.choice(np.array([-32768, 32767, -32767, 0, 1, -1], dtype=np.int16), size=2 * 4096 * 5)
    lpf = synth.lowpass_taps(T, 200000.0, fs)
    for off, gain in ((0, 1.0), (600000, 3.98), (-600000, 15.0)):
        r = ref.channel(lpf, off, fs, D, iq, gain=gain)
        o = oracle.channel(lpf, off, fs, D, iq, gain=gain)
        k = len(r[1])
        assert np.array_equal(r[0], o[0][:2 * k]) and np.array_equal(r[1], o[1][:k])


def test_fm_demod_and_atan2(oracle, ref):
    rng = np.random.default_rng(3)
    y = rng.integers(-32768, 32768, 2 * 20000).astype(np.int16)
    y[:40] = 0
    exp = ref.fm_demod(y)
    last = (0, 0)
    got = np.zeros(len(y) // 2, np.int16)
    for i in range(len(got)):
        a_re, a_im = int(y[2 * i]), int(y[2 * i + 1])
        s_re = a_re * last[0] + a_im * last[1]
        s_im = a_im * last[0] - a_re * last[1]
        s_re = (s_re + 2**31) % 2**32 - 2**31
        s_im = (s_im + 2**31) % 2**32 - 2**31
        got[i] = oracle.L.orc_fm_pcm(s_re, s_im, 1)
        last = (a_re, a_im)
    assert np.array_equal(exp, got)


@pytest.mark.parametrize("I,D,nt", [(4, 5, 97), (16, 25, 821), (192, 125, 2305), (1, 3, 31), (5, 1, 40)])
def test_resampler(oracle, ref, I, D, nt):
    rng = np.random.default_rng(I * 100 + D)
    pcm = np.clip(np.round(rng.normal(0, 9000, 1024 * 7)), -32768, 32767).astype(np.int16)
    taps = np.round(synth.lowpass_taps(nt, 0.45 * min(1.0 / I, 1.0 / D), 1.0) * I * 16384).astype(np.int16)
    exp = ref.resample(taps, I, D, pcm)
    got, _ = oracle.resample(taps, I, D, pcm)
    assert len(exp) > 50 and np.array_equal(got[:len(exp)], exp)


def test_bch_all_single_double_and_random(oracle, ref):
    rng = np.random.default_rng(1)
    cw = synth.pocsag_codeword(0x155555)
    base = int(f"{cw:032b}"[::-1], 2) & 0x7fffffff
    words = [base]
    for a in range(31):
        words.append(base ^ (1 << a))
        for b in range(a + 1, 31):
            words.append(base ^ (1 << a) ^ (1 << b))
    words += [int(x) for x in rng.integers(0, 1 << 31, 3000)]
    for w in words:
        assert oracle.bch_decode(w) == ref.bch_decode(w)
    assert oracle.bch_decode(base) == (0, base)


@pytest.mark.parametrize("baud,spb", [(512, 75), (1200, 32), (2400, 16)])
def test_pocsag_decoder(oracle, ref, baud, spb):
    """Clean and noisy NRZ at 38400 Hz straight into the decoders; tuples must be identical."""
    msgs = [(1234567, 3, "alpha", "HELLO B200 TEST 42"), (2007, 1, "numeric", "0123456789"),
            (8, 0, "alpha", "x" * 60), (123, 2, "numeric", "12-34 [5]U")]
    bits = synth.pocsag_bitstream(msgs)
    nrz = np.repeat(1 - 2 * bits.astype(np.int32), spb)
    rng = np.random.default_rng(baud)
    for sigma in (0.0, 0.35):
        pcm = np.concatenate([np.zeros(500), nrz, np.zeros(3000), nrz, np.zeros(777)]) * 3000.0
        pcm = pcm + rng.normal(0, 3000.0 * sigma, len(pcm))
        pcm = np.clip(np.round(pcm), -32768, 32767).astype(np.int16)
        exp = ref.pocsag(pcm)
        assert oracle.pocsag(pcm) == exp
        assert oracle.pocsag(pcm, chunk=777) == exp
        assert ref.pocsag(pcm, chunk=1000) == exp
        if sigma == 0.0:
            assert len(exp) >= 8 and exp[0][1] == baud


def test_random_pcm_into_pocsag(oracle, ref):
    """Noise-only input exercises false syncs, uncorrectable words and the SEARCH_SYNCWORD fallbacks."""
    rng = np.random.default_rng(77)
    pcm = rng.integers(-2000, 2000, 400000).astype(np.int16)
    # inject bare sync words followed by noise so that batches full of BCH failures get processed
    sync = np.array([(0x7CD215D8 >> (31 - b)) & 1 for b in range(32)])
    burst = np.repeat(1 - 2 * sync, 32) * 3000
    for pos in range(5000, 390000, 40000):
        pcm[pos:pos + len(burst)] = burst
    assert oracle.pocsag(pcm) == ref.pocsag(pcm)


def test_dc_blocker(oracle, ref):
    """decoder -b: resampler output through filter/dc_blocker.h (state carried across 1024-sample blocks)."""
    rng = np.random.default_rng(12)
    pcm = np.clip(np.round(rng.normal(900, 5000, 1024 * 8)), -32768, 32767).astype(np.int16)
    taps = np.round(synth.lowpass_taps(97, 0.09, 1.0) * 4 * 16384).astype(np.int16)
    for pole in (0.9999, 0.99):
        exp = ref.resample(taps, 4, 5, pcm, use_dc=True, pole=pole)
        got = oracle.dc_block(oracle.resample(taps, 4, 5, pcm)[0][:len(exp)], pole)
        assert len(exp) > 1000 and np.array_equal(got, exp)


def test_flex_decoder(oracle, ref):
    """a8: every coding, clean / bit-flipped / noisy frames, whole-buffer and chunked: identical callback tuples."""
    import flexcases
    total = 0
    for coding, trial, pcm in flexcases.all_cases():
        exp = ref.flex(pcm)
        assert oracle.flex(pcm) == exp, (coding, trial)
        assert oracle.flex(pcm, chunk=777) == exp, (coding, trial)
        total += len(exp)
        if trial == 0:
            kinds = {m[0] for m in exp}
            assert kinds == {2, 3, 4} and all(m[1] == int(coding.split("/")[0]) for m in exp)
    assert total > 300


def test_random_pcm_into_flex(oracle, ref):
    """Noise and stray sync patterns: false bit-syncs, unknown A words, uncorrectable FIWs."""
    from tsl_sdr_b200 import flexsynth
    for seed in range(6):
        rng = np.random.default_rng(seed)
        pcm = np.clip(np.round(rng.normal(0, 4000, 300000)), -32768, 32767).astype(np.int16)
        lv = flexsynth.frame_levels("1600/2", 1, 2, {})[:10 * (40 + 32 + 32 + 16 + 32 + 20)]      # sync 1 with a truncated FIW
        burst = np.round(lv * 5000).astype(np.int16)
        for pos in range(3000, 280000, 50000):
            pcm[pos:pos + len(burst)] = burst
        assert oracle.flex(pcm) == ref.flex(pcm)


def mm_signal(seed, spb=32.0, nbits=2500, jitter=0.002):
    """NRZ at a slightly wrong symbol rate, smoothed, with noise: something for the timing loop to track."""
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, nbits)
    t = np.arange(int(nbits * spb * (1 + jitter)) - 64) / (spb * (1 + jitter))
    x = (1 - 2 * bits[t.astype(np.int64)]).astype(np.float64) * 6000
    x = np.convolve(x, np.ones(9) / 9, mode="same") + rng.normal(0, 400, len(x))
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


MM_GAINS = [(1e-6, 1e-7), (3e-5, 2e-6), (1e-4, 1e-5), (0.0, 0.0)]


@pytest.mark.parametrize("kw,km", MM_GAINS)
def test_mueller_muller(oracle, kw, km):
    """f4: pager/mueller_muller.c against the restatement, with and without GNU C's FMA contraction, chunked too."""
    import pyoracle
    if not pyoracle.ref_available():
        pytest.skip("oracle/_ref not built")
    pcm = mm_signal(3)
    spb = 32.0
    for variant, fma in (("fma", 1), ("nofma", 0)):
        r = pyoracle.Ref(variant)
        exp, st = r.mm(pcm, kw, km, spb, spb * 0.9, spb * 1.1)
        got, gst = oracle.mm(pcm, kw, km, spb, spb * 0.9, spb * 1.1, fma=fma)
        assert len(exp) > 2000 and np.array_equal(exp, got) and np.array_equal(st, gst)
        got, gst = oracle.mm(pcm, kw, km, spb, spb * 0.9, spb * 1.1, chunk=1000, fma=fma)
        exp2, st2 = r.mm(pcm, kw, km, spb, spb * 0.9, spb * 1.1, chunk=1000)
        assert np.array_equal(exp2, got) and np.array_equal(st2, gst) and np.array_equal(exp2, exp)
