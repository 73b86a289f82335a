#!/usr/bin/env python
"""Measurement of the decoder side and of the whole chain on the device.

BASELINE configs[2] shape: 256 channels, 1.2 MS/s, D = 25 -> 48 kS/s PCM -> 4/5 resampler -> 38.4 kS/s -> POCSAG-1200
decode; configs[4] shape: 256 channels, 3 MS/s, D = 120 -> 25 kS/s -> 16/25 -> 16 kS/s -> FLEX decode.  Every channel
carries a synthetic FSK signal (the decoders' work depends on the data); the IQ batch goes through the channel bank,
the PCM stays on the device and feeds the pager bank.  Reported: wall time per batch of the two banks (CUDA work is
synchronised around each), the real-time factor (seconds of signal per second), and channel-samples/s of the
decoder kernels alone.  These kernels are one warp per channel and latency bound by construction (H5): the point
of the line is the margin over real time, not a roofline fraction.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tslb200_loader  # noqa: E402

tslb200_loader.load_package()
from tsl_sdr_b200 import synth  # noqa: E402
from tsl_sdr_b200.gpuchan import GpuChan, F_ATAN_FMA  # noqa: E402
from tsl_sdr_b200.gpupager import GpuPager, quantize_taps, DECODER_FLEX, DECODER_POCSAG  # noqa: E402


def run(name, C, fs, T, D, I, Dr, decoder, baud, seconds):
    from scipy.signal import firwin
    n = int(fs * seconds)
    offs = synth.channel_offsets(C, fs)
    lpf = synth.lowpass_taps(T, 9000.0, fs)
    # a few modulated channels + noise everywhere: the decoders of idle channels run the same per-sample state machines
    msgs = [None] * C
    for c in range(0, C, max(1, C // 8)):
        msgs[c] = [(1000 + c, c & 3, "alpha", f"CH{c:04d} TEST MESSAGE {c}")]
    silent = [c for c in range(C) if msgs[c] is None]           # noise only (synthesising 256 carriers on the host is slow)
    iq = synth.synth_pocsag_iq(n, fs, list(offs), msgs, baud=baud, amplitude=1200.0, silent_channels=set(silent))
    rt = firwin(24 * max(I, Dr) + 1, 0.8 / max(I, Dr)) * I
    chan = GpuChan(lpf, offs, fs, D, n, flags=F_ATAN_FMA)
    pager = GpuPager(C, n // D + 8, quantize_taps(rt), I, Dr, decoder=decoder)
    t_chan = t_pager = 0.0
    reps = 3
    nmsg = 0
    for r in range(reps + 1):
        t0 = time.perf_counter()
        chan.submit(iq)
        ptr, pitch, k = chan.device_pcm()
        chan.sync()
        t1 = time.perf_counter()
        pager.feed_device(ptr, pitch, k)
        got = pager.poll_full() if decoder == DECODER_FLEX else pager.dispatch()
        t2 = time.perf_counter()
        chan.discard()
        if r:                       # first pass = warm-up
            t_chan += t1 - t0
            t_pager += t2 - t1
            nmsg += len(got)
    t_chan /= reps
    t_pager /= reps
    print(json.dumps({"shape": name, "channels": C, "seconds_of_signal_per_batch": seconds,
                      "channel_bank_ms_per_batch_incl_h2d": round(t_chan * 1e3, 3),
                      "pager_bank_ms_per_batch_incl_callbacks": round(t_pager * 1e3, 3),
                      "realtime_factor_whole_chain": round(seconds / (t_chan + t_pager), 1),
                      "decoder_channel_samples_per_s": C * k / t_pager, "pcm_rate_hz": fs / D, "messages_per_batch": nmsg / reps,
                      "pager_kernel_launches": int(pager.kernel_launches)}))
    chan.close()
    pager.close()


if __name__ == "__main__":
    run("configs[2]: 256 ch POCSAG-1200, 1.2 MS/s, D=25, resample 4/5", 256, 1_200_000, 127, 25, 4, 5, DECODER_POCSAG, 1200, 2.0)
    run("configs[4] shape: 256 ch FLEX decoder, 3 MS/s, D=120, 512 taps, resample 16/25", 256, 3_000_000, 512, 120, 16, 25,
        DECODER_FLEX, 1600, 1.0)
