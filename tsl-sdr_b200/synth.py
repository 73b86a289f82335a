"""Synthetic wide-band IQ generator (test/bench tooling -- not on the product path).

Builds standard POCSAG transmissions (preamble, 0x7CD215D8 sync, BCH(31,21)+parity
codewords, 7-bit alpha payloads), 2-FSK modulates them at +-4.5 kHz, places one carrier
per channel offset and sums them into an interleaved int16 IQ stream -- the cs16 layout
multifm's file source reads (reference: multifm/file_if.c:47-64).  Recipes follow
SURVEY.md Appendix D, which were verified against the reference decoders.
"""
from __future__ import annotations

import numpy as np

POCSAG_SYNC = 0x7CD215D8
POCSAG_IDLE = 0x7A89C197
SEED = 20260925


def bch3121_parity(data21: int) -> int:
    """10 BCH check bits for 21 data bits, g(x) = x^10+x^9+x^8+x^6+x^5+x^3+1 (0x769)."""
    reg = data21 << 10
    for bit in range(30, 9, -1):
        if reg & (1 << bit):
            reg ^= 0x769 << (bit - 10)
    return reg & 0x3FF


def pocsag_codeword(data21: int) -> int:
    cw = (data21 << 11) | (bch3121_parity(data21) << 1)
    return cw | (bin(cw).count("1") & 1)


def pocsag_address_cw(address: int, function: int) -> int:
    return pocsag_codeword(((address >> 3) & 0x3FFFF) << 2 | (function & 3))


def pocsag_alpha_cws(text: str) -> list[int]:
    bits: list[int] = []
    for ch in text.encode("ascii") + b"\x04":
        bits += [(ch >> b) & 1 for b in range(7)]           # LSB first
    while len(bits) % 20:
        bits.append(0)
    out = []
    for i in range(0, len(bits), 20):
        v = 0
        for b in bits[i:i + 20]:
            v = (v << 1) | b
        out.append(pocsag_codeword((1 << 20) | v))
    return out


def pocsag_numeric_cws(digits: str) -> list[int]:
    table = {c: i for i, c in enumerate("0123456789XU -[]")}
    bits: list[int] = []
    for ch in digits:
        v = table[ch]
        bits += [(v >> b) & 1 for b in range(4)]            # LSB first within the BCD nibble
    while len(bits) % 20:
        bits += [0, 0, 1, 1]                                # pad with 0xC (space), LSB first
    out = []
    for i in range(0, len(bits), 20):
        v = 0
        for b in bits[i:i + 20]:
            v = (v << 1) | b
        out.append(pocsag_codeword((1 << 20) | v))
    return out


def pocsag_bitstream(messages, preamble_bits: int = 576) -> np.ndarray:
    """messages: list of (address, function, kind, text) with kind 'alpha'|'numeric'.
    Returns the transmitted bits (uint8), MSB-first words."""
    words: list[int] = []
    slot = 0                                                # position inside the current batch (0..15)

    def emit(cw):
        nonlocal slot
        if slot == 0:
            words.append(POCSAG_SYNC)
        words.append(cw)
        slot = (slot + 1) % 16

    for address, function, kind, text in messages:
        frame = address & 7
        while slot != 2 * frame:
            emit(POCSAG_IDLE)
        emit(pocsag_address_cw(address, function))
        for cw in (pocsag_alpha_cws(text) if kind == "alpha" else pocsag_numeric_cws(text)):
            emit(cw)
        emit(POCSAG_IDLE)
    while slot != 0:
        emit(POCSAG_IDLE)
    bits = [(i + 1) & 1 for i in range(preamble_bits)]      # 1010...
    for w in words:
        bits += [(w >> (31 - b)) & 1 for b in range(32)]
    return np.asarray(bits, dtype=np.uint8)


def nrz_waveform(bits: np.ndarray, baud: float, fs: float, n: int, start: int = 0, smooth: bool = True) -> np.ndarray:
    """+1/-1 waveform sampled at fs; bit 1 -> -1 (lower frequency), bit 0 -> +1
    (reference slicer polarity: pager/pager_pocsag.c:91).  Zero outside the burst."""
    t = (np.arange(n) - start) * (baud / fs)
    idx = np.floor(t).astype(np.int64)
    valid = (idx >= 0) & (idx < len(bits))
    w = np.zeros(n, dtype=np.float64)
    w[valid] = 1.0 - 2.0 * bits[idx[valid]].astype(np.float64)
    if smooth:
        k = max(1, int(fs / baud / 6))
        w = np.convolve(w, np.ones(k) / k, mode="same")
    return w


def fsk_phase(wave: np.ndarray, deviation_hz: float, fs: float) -> np.ndarray:
    return 2.0 * np.pi * deviation_hz / fs * np.cumsum(wave)


def channel_offsets(nr_channels: int, fs: int, span: float = 0.9) -> np.ndarray:
    """Uniform integer-Hz grid over +-span*fs/2."""
    c = np.arange(nr_channels)
    return np.round(-fs / 2 * span + c * (span * fs / nr_channels)).astype(np.int32)


def lowpass_taps(nr_taps: int, cutoff_hz: float, fs: float) -> np.ndarray:
    """Hamming-window low-pass, unity DC gain (what scipy.signal.firwin(T, cutoff, fs=fs) returns)."""
    m = np.arange(nr_taps) - (nr_taps - 1) / 2.0
    h = np.sinc(2.0 * cutoff_hz / fs * m) * np.hamming(nr_taps)
    return h / h.sum()


def to_cs16(x: np.ndarray) -> np.ndarray:
    out = np.empty(2 * len(x), dtype=np.int16)
    out[0::2] = np.clip(np.round(x.real), -32768, 32767).astype(np.int16)
    out[1::2] = np.clip(np.round(x.imag), -32768, 32767).astype(np.int16)
    return out


def synth_pocsag_iq(n: int, fs: int, offsets_hz, messages_per_channel, baud: int = 1200,
                    amplitude: float | None = None, noise_sigma: float = 50.0,
                    deviation_hz: float = 4500.0, seed: int = SEED, start: int = 2000,
                    silent_channels=()) -> np.ndarray:
    """IQ stream of n complex samples: one FM carrier per offset, each 2-FSK keyed with its own
    POCSAG burst.  messages_per_channel[c] is a list for pocsag_bitstream (or None = bare carrier)."""
    rng = np.random.default_rng(seed)
    nch = len(offsets_hz)
    if amplitude is None:
        amplitude = 3000.0 * np.sqrt(2.0) / np.sqrt(max(1, nch))     # total RMS ~3000 LSB
    tt = np.arange(n, dtype=np.float64)
    acc = np.zeros(n, dtype=np.complex128)
    for c, off in enumerate(offsets_hz):
        if c in silent_channels:
            continue
        msgs = messages_per_channel[c] if messages_per_channel is not None else None
        phase = 2.0 * np.pi * (float(off) / fs) * tt + rng.uniform(0, 2 * np.pi)
        if msgs:
            bits = pocsag_bitstream(msgs)
            wave = nrz_waveform(bits, baud, fs, n, start=start + 37 * c)
            phase = phase + fsk_phase(wave, deviation_hz, fs)
        acc += amplitude * np.exp(1j * phase)
    acc += noise_sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return to_cs16(acc)


def synth_noise_tones_iq(n: int, fs: int, offsets_hz, seed: int = SEED, rms: float = 3000.0,
                         noise_sigma: float = 50.0) -> np.ndarray:
    """Cheap throughput-bench input: unmodulated carriers on a subset of the channel grid + noise."""
    rng = np.random.default_rng(seed)
    tt = np.arange(n, dtype=np.float64)
    sel = list(offsets_hz)[:: max(1, len(offsets_hz) // 16)]
    amp = rms * np.sqrt(2.0) / np.sqrt(len(sel))
    acc = np.zeros(n, dtype=np.complex128)
    for off in sel:
        dev = 3000.0 * np.sin(2 * np.pi * rng.uniform(300, 1200) / fs * tt)      # slow FM tone
        acc += amp * np.exp(1j * (2.0 * np.pi * (float(off) / fs) * tt + np.cumsum(dev) * 2 * np.pi / fs))
    acc += noise_sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return to_cs16(acc)
