#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE's own objects (oracle/_ref, compiled unmodified from
/root/reference by oracle/Makefile).  Run here (where /root/reference exists); the .npz files are committed
so that the oracle and the CUDA path can be pinned on machines without the reference tree.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import pyoracle  # noqa: E402
import tslb200_loader  # noqa: E402

tslb200_loader.load_package()
from tsl_sdr_b200 import synth  # noqa: E402


def msgs_to_arrays(msgs):
    meta = np.array([[m[0], m[1], m[2] & 0xffffffff, m[3], m[4]] for m in msgs], dtype=np.int64).reshape(-1, 5)
    text = np.zeros((len(msgs), 520), dtype=np.uint8)
    for i, m in enumerate(msgs):
        text[i, :len(m[6])] = np.frombuffer(m[6], dtype=np.uint8)
    return meta, text


def main():
    pyoracle.build()
    ref = pyoracle.Ref("fma")
    ref_nofma = pyoracle.Ref("nofma")

    # ---- 1. FIR + FM on noise, several shapes (small) ----
    cases = [(127, 100, 2400000, [0, 312500, -1000001]), (127, 25, 1200000, [-320000, 150000]),
             (255, 200, 10000000, [4321000]), (512, 120, 3000000, [-1234567]), (16, 4, 2400000, [25000])]
    out = {}
    for i, (T, D, fs, offs) in enumerate(cases):
        rng = np.random.default_rng(1000 + i)
        n = 4096 * 6
        iq = np.clip(np.round(rng.normal(0, 6000, 2 * n)), -32768, 32767).astype(np.int16)
        lpf = synth.lowpass_taps(T, min(9000.0, fs / 8), fs)
        out[f"c{i}_params"] = np.array([T, D, fs], dtype=np.int64)
        out[f"c{i}_offs"] = np.array(offs, dtype=np.int32)
        out[f"c{i}_lpf"] = lpf
        out[f"c{i}_iq"] = iq
        for j, off in enumerate(offs):
            gain = [1.0, 2.5118864315095806, 0.5][j % 3]
            y, p, st = ref.channel(lpf, off, fs, D, iq, gain=gain, return_state=True)
            _, p_nofma = ref_nofma.channel(lpf, off, fs, D, iq, gain=gain)
            re, im = ref.prepare_taps(lpf, off, fs, gain)
            out[f"c{i}_{j}_gain"] = np.array([gain])
            out[f"c{i}_{j}_y"] = y
            out[f"c{i}_{j}_pcm"] = p
            out[f"c{i}_{j}_pcm_nofma"] = p_nofma
            out[f"c{i}_{j}_taps_re"] = re
            out[f"c{i}_{j}_taps_im"] = im
            out[f"c{i}_{j}_rot"] = np.concatenate([st[0], st[1]])
    np.savez_compressed(os.path.join(HERE, "fir_fm.npz"), **out)

    # ---- 2. fast_atan2f / FM on a grid of (s_im, s_re) ----
    rng = np.random.default_rng(7)
    v = np.concatenate([rng.integers(-2**31, 2**31, 4000), rng.integers(-70000, 70000, 4000),
                        np.array([0, 1, -1, 255, 256, -255, 65535, 2**31 - 1, -2**31, 16384, -16384, 3, 4])])
    s_im = rng.permutation(v)[:8000].astype(np.int64)
    s_re = rng.permutation(v)[:8000].astype(np.int64)
    phi = np.array([ref.fast_atan2f(np.float32(a), np.float32(b)) for a, b in zip(s_im, s_re)], dtype=np.float32)
    phi_nf = np.array([ref_nofma.fast_atan2f(np.float32(a), np.float32(b)) for a, b in zip(s_im, s_re)], dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "atan2.npz"), s_im=s_im, s_re=s_re, phi=phi, phi_nofma=phi_nf)

    # ---- 3. resampler (shipped shapes 4/5, 16/25, 192/125) ----
    out = {}
    rng = np.random.default_rng(9)
    pcm = np.clip(np.round(rng.normal(0, 5000, 1024 * 9)), -32768, 32767).astype(np.int16)
    out["pcm"] = pcm
    for name, (I, Dd, nt) in {"r4_5": (4, 5, 97), "r16_25": (16, 25, 821), "r192_125": (192, 125, 2305), "r3_2": (3, 2, 10)}.items():
        taps = np.round(synth.lowpass_taps(nt, 0.45 * min(1.0 / I, 1.0 / Dd), 1.0) * I * 16384).astype(np.int16)
        out[name + "_taps"] = taps
        out[name + "_out"] = ref.resample(taps, I, Dd, pcm)
    np.savez_compressed(os.path.join(HERE, "resampler.npz"), **out)

    # ---- 4. BCH(31,21) ----
    rng = np.random.default_rng(11)
    words = []
    for _ in range(300):
        cw = synth.pocsag_codeword(int(rng.integers(0, 1 << 21)))
        w = int(f"{cw:032b}"[::-1], 2) & 0x7fffffff           # reference convention: first received bit = LSB
        nerr = int(rng.integers(0, 4))
        for b in rng.choice(31, nerr, replace=False):
            w ^= 1 << int(b)
        words.append(w)
    words += [0, 0x7fffffff, 1, 0x40000000, 0x12345678 & 0x7fffffff]
    res = [ref.bch_decode(w) for w in words]
    np.savez_compressed(os.path.join(HERE, "bch.npz"), words=np.array(words, dtype=np.uint32),
                        rc=np.array([r[0] for r in res], dtype=np.int32), out=np.array([r[1] for r in res], dtype=np.uint32))

    # ---- 5. whole chain: IQ -> FIR/FM -> 4/5 resampler -> POCSAG (SURVEY.md appendix D template) ----
    fs, D, T = 1200000, 25, 127
    offs = [-320000, 150000, 0, 123457]
    msgs = [[(1234567, 3, "alpha", "HELLO B200 TEST 42")], [(2007, 1, "numeric", "0123456789")], None,
            [(1000, 0, "alpha", "CH0003 TEST MESSAGE 1"), (1001, 2, "numeric", "555-0199 [7]")]]
    n = 4096 * 660
    iq = synth.synth_pocsag_iq(n, fs, offs, msgs, baud=1200)
    lpf = synth.lowpass_taps(T, 9000.0, fs)
    rtaps = np.round(synth.lowpass_taps(97, 14000.0, 192000.0) * 4 * 16384).astype(np.int16)
    out = {"params": np.array([T, D, fs, n], dtype=np.int64), "offs": np.array(offs, dtype=np.int32), "lpf": lpf, "rtaps": rtaps}
    for c, off in enumerate(offs):
        _, pcm = ref.channel(lpf, off, fs, D, iq)
        res = ref.resample(rtaps, 4, 5, pcm)
        m = ref.decoder_pocsag(rtaps, 4, 5, pcm)
        assert m == ref.pocsag(res)
        meta, text = msgs_to_arrays(m)
        out[f"ch{c}_pcm_crc"] = np.array([int(np.bitwise_xor.reduce(pcm.astype(np.int64) * np.arange(1, len(pcm) + 1))), len(pcm)])
        out[f"ch{c}_res"] = res
        out[f"ch{c}_meta"] = meta
        out[f"ch{c}_text"] = text
    np.savez_compressed(os.path.join(HERE, "pocsag_chain.npz"), **out)
    # ---- 6. FLEX: all four codings, clean / corrupted / noisy frames straight into pager_flex_on_pcm ----
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import flexcases
    out = {}
    total = 0
    for coding, trial, pcm in flexcases.all_cases():
        m = ref.flex(pcm)
        assert m == ref.flex(pcm, chunk=1000)
        key = f"{coding.replace('/', '_')}_t{trial}"
        out[key + "_crc"] = np.array([int(np.bitwise_xor.reduce(pcm.astype(np.int64) * np.arange(1, len(pcm) + 1))), len(pcm)])
        out[key + "_meta"], out[key + "_text"] = flexcases.msgs_to_arrays(m)
        total += len(m)
    assert total > 300
    np.savez_compressed(os.path.join(HERE, "flex.npz"), **out)
    print("golden fixtures written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
