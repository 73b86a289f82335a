#!/usr/bin/env python
"""Kernel-level survey of BASELINE.json's other shapes (diagnostics; bench.py stays on configs[1]).

For every shape: device-resident throughput of the fused FIR+FM kernel (CUDA events inside the C ABI), the engine
that was selected, and the achieved fraction of the HBM roofline on algorithmic bytes 4N + 2CK.  Inputs are Gaussian
noise (throughput does not depend on the data); parity for these shapes is covered by tests/test_gpu_parity.py.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tslb200_loader  # noqa: E402

tslb200_loader.load_package()
from tsl_sdr_b200 import synth  # noqa: E402
from tsl_sdr_b200.gpuchan import GpuChan, F_ATAN_FMA  # noqa: E402

def shipped_taps(name):
    """a low-pass prototype as the reference ships it (etc/*.json: 256-tap etc/pocsag_1200khz_fs.json, 512-tap
    etc/flex_25khz_lpf_3mhz.json); regenerated with the same design parameters because /root/reference does not travel"""
    if name == "pocsag_1200khz_fs":
        return synth.lowpass_taps(256, 9000.0, 1_200_000)
    return synth.lowpass_taps(512, 9000.0, 3_000_000)


# (name, C, fs, T, D, cutoff[, dBGain[, shipped tap set]])
SHAPES = [
    ("c1  1 ch x 127 taps, D=100, 2.4 MS/s", 1, 2_400_000, 127, 100, 9000.0),
    ("c2  64 ch x 127 taps, D=100, 2.4 MS/s", 64, 2_400_000, 127, 100, 9000.0),
    ("c3  256 ch x 127 taps, D=25, 1.2 MS/s", 256, 1_200_000, 127, 25, 9000.0),
    ("c3' 256 ch x 256 taps, D=25, 1.2 MS/s", 256, 1_200_000, 256, 25, 9000.0),
    ("c4  128 ch x 255 taps, D=200, 10 MS/s (1024 ch over 8 GPUs)", 128, 10_000_000, 255, 200, 12000.0),
    ("c4' 1024 ch x 255 taps, D=200, 10 MS/s on one GPU", 1024, 10_000_000, 255, 200, 12000.0),
    ("c5  256 ch x 512 taps, D=120, 3 MS/s", 256, 3_000_000, 512, 120, 9000.0),
    ("headline 256 ch x 127 taps, D=100, 2.4 MS/s", 256, 2_400_000, 127, 100, 9000.0),
    # gain coverage: dBGain is applied as 10^(dB/10) to the taps (multifm/receiver.c:220); beyond |tap| = 508 the int8 sum split
    # no longer fits and the engine switches to radix-256 limbs (4 MMAs per K chunk instead of 2, 3 accumulators)
    ("headline + dBGain 4 (sum split, 3 terms)", 256, 2_400_000, 127, 100, 9000.0, 4.0),
    ("headline + dBGain 8 (radix split)", 256, 2_400_000, 127, 100, 9000.0, 8.0),
    ("c2 + dBGain 8 (radix split)", 64, 2_400_000, 127, 100, 9000.0, 8.0),
    ("c3 with the shipped 256-tap etc/pocsag_1200khz_fs.json design, dBGain 4 (etc/pocsag_rtlsdr.json)", 256, 1_200_000, 256, 25, 9000.0, 4.0, "pocsag_1200khz_fs"),
    ("c5 with the shipped 512-tap etc/flex_25khz_lpf_3mhz.json design", 256, 3_000_000, 512, 120, 9000.0, 0.0, "flex_25khz_lpf_3mhz"),
]


def main():
    peak = 6450.3
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    n = 1 << int(os.environ.get("BATCH_LOG2", "24"))
    rng = np.random.default_rng(1)
    iq = np.clip(np.round(rng.normal(0, 3000, 2 * n)), -32768, 32767).astype(np.int16)
    for name, C, fs, T, D, cut, *rest in SHAPES:
        gain_db = rest[0] if rest else 0.0
        lpf = shipped_taps(rest[1]) if len(rest) > 1 else synth.lowpass_taps(T, cut, fs)
        offs = synth.channel_offsets(C, fs)
        gains = None if not gain_db else [10.0 ** (gain_db / 10.0)] * C
        try:
            bank = GpuChan(lpf, offs, fs, D, n, gains=gains, flags=F_ATAN_FMA, engine=int(os.environ.get("ENGINE", "0")))
        except Exception as exc:
            print(json.dumps({"shape": name, "error": str(exc)}))
            continue
        for _ in range(3):
            bank.submit(iq)
            k = bank.pending()
            bank.discard()
        bank.sync()
        bank.timing_read()
        bank.timing_enable(True)
        for _ in range(5):
            bank.submit(iq)
            bank.discard()
        bank.sync()
        ms, cnt = bank.timing_read()
        ms /= max(1, cnt)
        alg = 4.0 * n + 2.0 * C * k
        mma_per_tile = bank.tc_model()[0]
        print(json.dumps({"shape": name, "engine": {1: "imad", 2: "tc"}.get(bank.engine, "?"), "mma_per_tile": mma_per_tile, "kernel_ms": round(ms, 4),
                          "channel_samples_per_s": C * k / (ms * 1e-3), "iq_msps": n / (ms * 1e-3) / 1e6,
                          "alg_GBps": alg / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak,
                          "int16_mac_per_s": 4.0 * T * C * k / (ms * 1e-3)}))
        bank.close()


if __name__ == "__main__":
    main()
