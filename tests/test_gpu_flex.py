"""GPU parity for the FLEX decoder (row a8): the flex kernel behind include/tslb200_gpupager.h against the oracle and
the golden fixture recorded from the reference's pager_flex objects.  Integer / bitwise work: identical tuples or fail.

Reference under test: pager/pager_flex.c (Sync 1 :295-458, Sync 2 :460-525, block :1200-1310, phase walk :1088-1198,
vector decoders :527-1033), decoder/decoder.c:581-673 (resampler -> [dc blocker] -> pager_flex_on_pcm)."""
import os

import numpy as np
import pytest

import flexcases
from tsl_sdr_b200 import flexsynth, synth
from tsl_sdr_b200.gpuchan import GpuChan, F_ATAN_FMA
from tsl_sdr_b200.gpupager import GpuPager, F_DC_BLOCK, F_INVERT, F_KEEP_PCM, F_NO_RESAMPLE, DECODER_FLEX

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_flex_bank(pcms, chunks=None, flags=F_NO_RESAMPLE, via_callbacks=False, **kw):
    """pcms: list of equal-length int16 arrays, one per channel -> per-channel list of oracle-shaped tuples"""
    C = len(pcms)
    n = len(pcms[0])
    x = np.stack(pcms)
    chunks = chunks or [n]
    bank = GpuPager(C, max(chunks), flags=flags, decoder=DECODER_FLEX, **kw)
    got = [[] for _ in range(C)]
    pos = 0
    for k in chunks:
        k = min(k, n - pos)
        if k <= 0:
            break
        bank.feed(x[:, pos:pos + k])
        for ch, m in (bank.dispatch_flex() if via_callbacks else bank.poll_full()):
            got[ch].append(m)
        pos += k
    assert bank.dropped == 0 and bank.kernel_launches > 0
    bank.close()
    return got


@pytest.mark.parametrize("coding", flexcases.CODINGS)
def test_flex_matches_oracle_and_golden(oracle, coding):
    """six trials of one coding run as six channels of one bank; whole feed, 1024-sample feeds and ragged feeds"""
    z = np.load(os.path.join(G, "flex.npz"))
    pcms = [flexcases.case_pcm(coding, t) for t in range(flexcases.TRIALS)]
    n = max(len(p) for p in pcms)
    pcms = [np.concatenate([p, np.full(n - len(p), p[-1], np.int16)]) for p in pcms]
    exp = [oracle.flex(p) for p in pcms]
    for t in range(flexcases.TRIALS):
        key = f"{coding.replace('/', '_')}_t{t}"
        assert exp[t] == flexcases.arrays_to_msgs(z[key + "_meta"], z[key + "_text"])
    assert sum(len(e) for e in exp) > 40
    for chunks, cb in (([n], False), ([1024] * (n // 1024 + 1), True), ([1, 9, 4999, 12345, 7, n], False)):
        got = run_flex_bank(pcms, chunks, via_callbacks=cb)
        for t in range(flexcases.TRIALS):
            assert got[t] == exp[t], (coding, t, chunks[:3])


def test_flex_noise_and_false_syncs(oracle):
    pcms = []
    for seed in range(8):
        rng = np.random.default_rng(seed)
        pcm = np.clip(np.round(rng.normal(0, 4000, 200000)), -32768, 32767).astype(np.int16)
        lv = flexsynth.frame_levels("3200/4", 1, 2, {})[:10 * (40 + 32 + 32 + 16 + 32 + 20)]
        burst = np.round(lv * 5000).astype(np.int16)
        for pos in range(3000, 180000, 50000):
            pcm[pos:pos + len(burst)] = burst
        pcms.append(pcm)
    exp = [oracle.flex(p) for p in pcms]
    got = run_flex_bank(pcms, [65536] * 4)
    assert got == exp


def test_flex_dc_blocker_and_invert(oracle):
    """decoder -b and -i: DC blocker (filter/dc_blocker.h) in front of the FLEX state machine; inverted input"""
    pcm = flexcases.case_pcm("1600/2", 1)
    off = np.clip(pcm.astype(np.int32) + 1500, -32768, 32767).astype(np.int16)
    exp_dc = oracle.flex(oracle.dc_block(off, 0.999))
    got_dc = run_flex_bank([off], [4096] * (len(off) // 4096 + 1), flags=F_NO_RESAMPLE | F_DC_BLOCK, dc_pole=0.999)[0]
    assert got_dc == exp_dc and len(exp_dc) > 5
    inv = (-pcm.astype(np.int32)).astype(np.int16)
    exp = oracle.flex(pcm)
    got = run_flex_bank([inv], [5000] * (len(inv) // 5000 + 1), flags=F_NO_RESAMPLE | F_INVERT)[0]
    assert got == exp and len(exp) > 5


def test_flex_through_resampler(oracle):
    """25 kHz FM audio -> 16/25 polyphase resampler (decoder -I 16 -D 25, etc/resampler_filter.json shape) -> FLEX"""
    I, D, nt = 16, 25, 821
    taps = np.round(synth.lowpass_taps(nt, 0.45 / D, 1.0) * I * 16384).astype(np.int16)
    pcms25 = []
    for coding in ("1600/2", "6400/4"):
        lv = flexcases.case_levels(coding, 0)
        t = np.arange(int(len(lv) * 25 / 16)) * (16 / 25)
        lv25 = lv[np.minimum(t.astype(np.int64), len(lv) - 1)]
        pcms25.append(flexsynth.pcm_from_levels(lv25, amplitude=6000.0, smooth=3))
    n = min(len(p) for p in pcms25)
    pcms25 = [p[:n] for p in pcms25]
    exp = [oracle.flex(oracle.resample(taps, I, D, p)[0]) for p in pcms25]
    assert all(len(e) > 5 for e in exp)
    bank = GpuPager(2, 8192, taps, I, D, decoder=DECODER_FLEX)
    got = [[], []]
    x = np.stack(pcms25)
    for pos in range(0, n, 8192):
        bank.feed(x[:, pos:pos + 8192])
        for ch, m in bank.poll_full():
            got[ch].append(m)
    bank.close()
    assert got == exp


def test_flex_full_chain_config5_shape(oracle):
    """BASELINE config 5 shape: 3 MS/s cs16 IQ, 512-tap LPF, decimate by 120 (25 kHz), 16/25 resampler, FLEX.
    IQ -> channel bank -> pager bank chained on the device vs the oracle's CPU chain, message for message."""
    fs, T, Dd = 3_000_000, 512, 120
    offs = np.array([-1_000_000, -250_000, 125_000, 987_500], dtype=np.int32)
    codings = ["1600/2", "3200/2", "3200/4", "6400/4"]
    lpf = synth.lowpass_taps(T, 9000.0, fs)
    rt = np.round(synth.lowpass_taps(821, 0.45 / 25, 1.0) * 16 * 16384).astype(np.int16)
    levels = [flexcases.case_levels(c, 0)[: 16000 * 2 + 4000] for c in codings]        # one frame each
    n = int(max(len(l) for l in levels) * fs / 16000) + 20000
    tt = np.arange(n, dtype=np.float64)
    acc = np.zeros(n, dtype=np.complex128)
    rng = np.random.default_rng(5)
    for off, lv in zip(offs, levels):
        idx = np.minimum((tt * (16000.0 / fs)).astype(np.int64), len(lv) - 1)
        wave = np.where(tt * (16000.0 / fs) < len(lv), lv[idx], -1.0)
        phase = 2 * np.pi * (float(off) / fs) * tt + 2 * np.pi * 4800.0 / fs * np.cumsum(wave)
        acc += 1800.0 * np.exp(1j * phase)
    acc += 30.0 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    iq = synth.to_cs16(acc)
    exp = []
    for off in offs:
        _, pcm = oracle.channel(lpf, off, fs, Dd, iq)
        exp.append(oracle.flex(oracle.resample(rt, 16, 25, pcm)[0]))
    assert sum(len(e) for e in exp) >= 20
    batch = 1 << 21
    bank = GpuChan(lpf, offs, fs, Dd, batch, flags=F_ATAN_FMA)
    pager = GpuPager(len(offs), batch // Dd + 8, rt, 16, 25, decoder=DECODER_FLEX)
    got = [[] for _ in offs]
    for pos in range(0, n, batch):
        k = min(batch, n - pos)
        bank.submit(iq[2 * pos: 2 * (pos + k)])
        bank.sync()
        ptr, pitch, cnt = bank.device_pcm()
        pager.feed_device(ptr, pitch, cnt)
        for ch, m in pager.poll_full():
            got[ch].append(m)
        bank.discard()
    bank.close()
    pager.close()
    assert got == exp
