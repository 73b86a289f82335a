"""GPU parity at full size for the two decoder configurations of BASELINE.json.

configs[2]: 256-channel channeliser (etc/pocsag_rtlsdr.json shape: 1.2 MS/s, decimate by 25 -> 48 kHz) + FM + 4/5
  resampler + POCSAG, IQ -> channel bank -> pager bank chained on the device.  EVERY channel's resampled PCM and message
  tuple list is compared with the oracle's CPU chain (multifm/demod.c:49-121 -> decoder/decoder.c:581-673 ->
  pager/pager_pocsag.c:434-543): 32 channels carry POCSAG-1200/2400 bursts, their neighbours (4.2 kHz away) see the same
  bursts off-centre, the rest see noise.
configs[4]: mixed FLEX / POCSAG receiver at the etc/flex_25khz_lpf_3mhz.json shape (3 MS/s, 512 taps, decimate by 120 ->
  25 kHz): the reference runs one `decoder` process per channel and picks protocol and resampler per process
  (decoder/decoder.c:685-697): FLEX channels use 16/25 -> 16 kHz, POCSAG channels 192/125 -> 38.4 kHz.  Here: two pager
  banks with channel maps over ONE channel bank's device PCM."""
import numpy as np
import pytest

import flexcases
from tsl_sdr_b200 import synth
from tsl_sdr_b200.gpuchan import GpuChan, F_ATAN_FMA
from tsl_sdr_b200.gpupager import GpuPager, F_KEEP_PCM, DECODER_FLEX, DECODER_POCSAG

pytestmark = pytest.mark.gpu


def orc_tuple(c, m):
    return (c, m[0], m[1], m[2] & 0xffffffff, m[3], m[4], m[6])


def test_config3_256_channel_pocsag_chain_every_channel(oracle):
    fs, T, D, C = 1_200_000, 127, 25, 256
    n = 1_760_000                                           # 1664 bits at 1200 baud + the latest burst start
    offs = synth.channel_offsets(C, fs)
    lpf = synth.lowpass_taps(T, 9000.0, fs)
    rtaps = np.round(synth.lowpass_taps(97, 14000.0, 192000.0) * 4 * 16384).astype(np.int16)   # decoder -I 4 -D 5 prototype
    active = list(range(3, C, 8))                           # 32 carriers, 33.75 kHz apart
    msgs = [None] * C
    for i, c in enumerate(active):
        msgs[c] = [(1000 + c, c & 3, "alpha", f"CH{c:04d} TEST MESSAGE {i}"), (2000 + c, (c >> 2) & 3, "numeric", f"{c:03d}-555 [{i}]")]
    rng = np.random.default_rng(3)
    tt = np.arange(n, dtype=np.float64)
    acc = 40.0 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    for i, c in enumerate(active):
        baud = 1200 if i % 2 == 0 else 2400
        wave = synth.nrz_waveform(synth.pocsag_bitstream(msgs[c]), baud, fs, n, start=2000 + 997 * i)
        acc += 900.0 * np.exp(1j * (2.0 * np.pi * (float(offs[c]) / fs) * tt + rng.uniform(0, 2 * np.pi)
                                    + synth.fsk_phase(wave, 4500.0, fs)))
    iq = synth.to_cs16(acc)
    del acc, tt

    batch = 1 << 19
    chan = GpuChan(lpf, offs, fs, D, batch, flags=F_ATAN_FMA)
    pager = GpuPager(C, batch // D + 8, rtaps, 4, 5, flags=F_KEEP_PCM)
    res, got = [], []
    for s in range(0, n, batch):
        chan.submit(iq[2 * s: 2 * min(n, s + batch)])
        ptr, pitch, k = chan.device_pcm()
        chan.sync()
        pager.feed_device(ptr, pitch, k)
        chan.discard()
        res.append(pager.collect_pcm(batch * 4 // (5 * D) + 64).copy())
        got += pager.dispatch()
    res = np.concatenate(res, axis=1)
    assert chan.engine == 2 and pager.dropped == 0
    chan.close(); pager.close()

    exp, decoded_channels = [], set()
    for c in range(C):
        _, pcm = oracle.channel(lpf, offs[c], fs, D, iq)
        r = oracle.resample(rtaps, 4, 5, pcm)[0]
        assert res.shape[1] == len(r) and np.array_equal(res[c], r), f"channel {c}: resampled PCM differs"
        for m in oracle.pocsag(r):
            exp.append(orc_tuple(c, m))
            decoded_channels.add(c)
    assert set(active) <= decoded_channels and len(exp) >= 2 * len(active)
    for c in range(C):          # per channel: same tuples in the same (decode) order
        assert [g for g in got if g[0] == c] == [e for e in exp if e[0] == c], f"channel {c}: messages differ"
    assert len(got) == len(exp)


def test_config5_mixed_flex_pocsag_banks_over_one_channel_bank(oracle):
    fs, T, Dd = 3_000_000, 512, 120
    offs = np.array([-1_200_000, -1_000_000, -612_500, -250_000, 125_000, 400_000, 987_500, 1_212_500], dtype=np.int32)
    kinds = ["pocsag", "flex", "pocsag", "flex", "flex", "pocsag", "flex", "pocsag"]
    codings = iter(["1600/2", "3200/2", "3200/4", "6400/4"])
    lpf = synth.lowpass_taps(T, 9000.0, fs)
    rt_flex = np.round(synth.lowpass_taps(821, 0.45 / 25, 1.0) * 16 * 16384).astype(np.int16)           # 16/25 -> 16 kHz
    rt_poc = np.round(synth.lowpass_taps(2305, 0.45 / 192, 1.0) * 192 * 16384).astype(np.int16)         # 192/125 -> 38.4 kHz
    n = int((16000 * 2 + 4000) * fs / 16000) + 20000
    tt = np.arange(n, dtype=np.float64)
    rng = np.random.default_rng(11)
    acc = 30.0 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    for c, (off, kind) in enumerate(zip(offs, kinds)):
        if kind == "flex":
            lv = flexcases.case_levels(next(codings), 0)[: 16000 * 2 + 4000]
            pos = tt * (16000.0 / fs)
            wave = np.where(pos < len(lv), lv[np.minimum(pos.astype(np.int64), len(lv) - 1)], -1.0)
            dev = 4800.0
        else:
            bits = synth.pocsag_bitstream([(1234560 + c, c & 3, "alpha", f"MIXED RECEIVER CH{c}"), (77 + c, 1, "numeric", f"0{c}-42")])
            wave = synth.nrz_waveform(bits, 1200 if c < 4 else 2400, fs, n, start=5000 + 1111 * c)
            dev = 4500.0
        acc += 1500.0 * np.exp(1j * (2 * np.pi * (float(off) / fs) * tt + synth.fsk_phase(wave, dev, fs)))
    iq = synth.to_cs16(acc)
    del acc, tt
    flex_ch = [c for c, k in enumerate(kinds) if k == "flex"]
    poc_ch = [c for c, k in enumerate(kinds) if k == "pocsag"]

    exp = {}
    for c, off in enumerate(offs):
        _, pcm = oracle.channel(lpf, off, fs, Dd, iq)
        if kinds[c] == "flex":
            exp[c] = oracle.flex(oracle.resample(rt_flex, 16, 25, pcm)[0])
        else:
            exp[c] = [(m[0], m[1], m[2], m[3], m[4], (0,) * 6, m[6]) for m in oracle.pocsag(oracle.resample(rt_poc, 192, 125, pcm)[0])]
    assert all(len(exp[c]) >= 2 for c in poc_ch) and sum(len(exp[c]) for c in flex_ch) >= 20

    batch = 1 << 21
    bank = GpuChan(lpf, offs, fs, Dd, batch, flags=F_ATAN_FMA)
    pf = GpuPager(len(flex_ch), batch // Dd + 8, rt_flex, 16, 25, decoder=DECODER_FLEX, channel_map=flex_ch)
    pp = GpuPager(len(poc_ch), batch // Dd + 8, rt_poc, 192, 125, decoder=DECODER_POCSAG, channel_map=poc_ch)
    got = {c: [] for c in range(len(offs))}
    for pos in range(0, n, batch):
        k = min(batch, n - pos)
        bank.submit(iq[2 * pos: 2 * (pos + k)])
        bank.sync()
        ptr, pitch, cnt = bank.device_pcm()
        pf.feed_device(ptr, pitch, cnt)                     # both banks read the same device PCM, each its own rows
        pp.feed_device(ptr, pitch, cnt)
        for ch, m in pf.poll_full() + pp.poll_full():
            got[ch].append(m)
        bank.discard()
    assert pf.dropped == 0 and pp.dropped == 0
    bank.close(); pf.close(); pp.close()
    for c in range(len(offs)):
        assert got[c] == exp[c], f"channel {c} ({kinds[c]})"
