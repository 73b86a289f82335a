#!/usr/bin/env python
"""Wall time of the first (cold: derotators still in their transient) and of a steady-state submit+sync, bench shape."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tslb200_loader
tslb200_loader.load_package()
from tsl_sdr_b200 import synth
from tsl_sdr_b200.gpuchan import GpuChan, F_ATAN_FMA
fs, T, D, C, n = 2400000, 127, 100, 64, 1 << 25
iq = np.clip(np.round(np.random.default_rng(1).normal(0, 3000, 2 * n)), -32768, 32767).astype(np.int16)
import ctypes
b = GpuChan(synth.lowpass_taps(T, 9000.0, fs), synth.channel_offsets(C, fs), fs, D, n, flags=F_ATAN_FMA)
for i in range(4):
    t0 = time.perf_counter(); b.submit(iq); b.sync(); t1 = time.perf_counter(); b.discard()
    print(f"submit {i}: {1e3 * (t1 - t0):.2f} ms (pageable H2D of 134 MB included), launches so far {b.kernel_launches}")
