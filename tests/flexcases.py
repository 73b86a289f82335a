"""Deterministic FLEX test inputs shared by the golden generator, the oracle tests and the GPU parity tests.

Every case is a 16 kHz int16 PCM stream holding two frames of one coding (pager_flex.c:47-96) with a mix of
alphanumeric / numeric / tone / short-instruction messages on short and long addresses; later trials flip bits in
random words (BCH corrects up to two per word, more makes words, vectors or whole phases drop out), add noise,
smoothing and a DC offset.  What matters is not that everything decodes but that every implementation reports
exactly what the reference reports."""
import numpy as np

from tsl_sdr_b200 import flexsynth as fs

BASE_MSGS = [dict(addr=fs.short_address(1234567), kind="alpha", text="HELLO FLEX ON B200"),
             dict(addr=fs.short_address(4242), kind="numeric", digits="0123456789"),
             dict(addr=fs.short_address(777), kind="tone", digits="123"),
             dict(addr=fs.short_address(9000), kind="siv", siv_type=1, data=0x2AA),
             dict(addr=fs.long_address(0x1E0005, 0x1FFFF0), kind="alpha", text="LONG ADDRESS MSG"),
             dict(addr=fs.long_address(0x1E0100, 0x1FFF00), kind="numeric", digits="98765 43210-[]"),
             dict(addr=fs.long_address(0x1E0101, 0x1FFF01), kind="tone", digits="98U", second="12345"),
             dict(addr=fs.short_address(31337), kind="alpha", text="FRAGMENT", seq=1, fragment=True, maildrop=True)]

CODINGS = list(fs.CODINGS)
TRIALS = 6


def case_levels(coding, trial):
    rng = np.random.default_rng(1000 * CODINGS.index(coding) + trial)
    names = fs.CODINGS[coding]["phases"]
    phases = {}
    for i, n in enumerate(names):
        msgs = [dict(addr=fs.short_address(1000 + i), kind="alpha", text=f"PHASE {n} {coding} T{trial}")]
        if i == 0:
            msgs += BASE_MSGS
        else:
            msgs += [dict(addr=fs.short_address(2000 + 10 * i + trial), kind="numeric", digits=f"{trial}{i}55-0199")]
        w = fs.build_phase(msgs, extra_biw=(trial % 2) if i == 0 else 0)
        if trial >= 2:
            w = w.copy()
            for _ in range(3 * (trial - 1)):
                k = int(rng.integers(0, 88))
                for b in rng.integers(0, 32, int(rng.integers(1, 4))):
                    w[k] ^= np.uint64(1 << int(b))
        phases[n] = w
    return np.concatenate([fs.frame_levels(coding, (3 + f) % 16, (17 + f + trial) % 128, phases) for f in range(2)])


def case_pcm(coding, trial):
    lv = case_levels(coding, trial)
    noise = 0.0 if trial < 4 else 700.0 * (trial - 3)
    return fs.pcm_from_levels(lv, amplitude=5000.0 + 500.0 * trial, noise_sigma=noise, smooth=(trial % 3) * 2,
                              dc=(trial % 4) * 300.0, seed=trial)


def all_cases():
    for coding in CODINGS:
        for trial in range(TRIALS):
            yield coding, trial, case_pcm(coding, trial)


def msgs_to_arrays(msgs):
    """oracle-shaped tuples (kind, baud, capcode, function, len, aux, text) -> (meta int64 [n, 11], text uint8 [n, 520])"""
    meta = np.array([[m[0], m[1], m[2], m[3], m[4]] + list(m[5]) for m in msgs], dtype=np.int64).reshape(-1, 11)
    text = np.zeros((len(msgs), 520), dtype=np.uint8)
    for i, m in enumerate(msgs):
        text[i, :len(m[6])] = np.frombuffer(m[6], dtype=np.uint8)
    return meta, text


def arrays_to_msgs(meta, text):
    out = []
    for me, tx in zip(meta, text):
        me = [int(v) for v in me]
        out.append((me[0], me[1], me[2], me[3], me[4], tuple(me[5:11]), bytes(tx[:me[4]])))
    return out
