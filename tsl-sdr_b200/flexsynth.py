"""Synthetic FLEX transmissions (test tooling -- not on the product path).

Builds frames the reference decoder (pager/pager_flex.c) accepts, for all four codings it knows
(1600/2, 3200/2, 3200/4, 6400/4; table at pager_flex.c:47-96): Sync 1 (bit sync, A, B, inverted A, FIW at
1600 bit/s 2-FSK), Sync 2 (comma, C, inverted comma, inverted C at the frame's symbol rate) and one
interleaved block of 88 words per phase.  Word layout, checksums, interleaving and the address / vector /
message formats follow what the reference *parses* (file:line given at each builder); where the reference
deviates from the published FLEX standard the reference wins, because parity is judged against it.

Output is the 16 kHz int16 PCM stream `pager_flex_on_pcm` consumes (10 samples per 1600-baud symbol).
"""
from __future__ import annotations

import numpy as np

SYNC_BS1 = 0xAAAAAAAA           # pager_flex_priv.h:416
SYNC_MAGIC_A = 0x5939           # pager_flex_priv.h:421
SYNC_MAGIC_B = 0x5555           # pager_flex_priv.h:426
SYNC2_C = 0xED84                # pager_flex_priv.h:441

# pager_flex.c:47-96
CODINGS = {
    "1600/2": dict(seq_a=0x78F3, baud=1600, levels=2, sps=10, sync2=4, sym_bits=1, symbols=2816, phases=("A",)),
    "3200/2": dict(seq_a=0x84E7, baud=3200, levels=2, sps=5, sync2=24, sym_bits=1, symbols=5632, phases=("A", "C")),
    "3200/4": dict(seq_a=0x4F97, baud=3200, levels=4, sps=10, sync2=12, sym_bits=2, symbols=2816, phases=("A", "C")),
    "6400/4": dict(seq_a=0x215F, baud=6400, levels=4, sps=5, sync2=32, sym_bits=2, symbols=5632, phases=("A", "B", "C", "D")),
}
NUM_LUT = "0123456789XU -]["    # pager_flex.c:687-704 (note: ']' before '[')


# ---------------------------------------------------------------------------------------------
# words
# ---------------------------------------------------------------------------------------------
def bch_word(data21: int) -> int:
    """32-bit FLEX word: bits 0-20 data, bits 21-30 BCH(31,21) check bits in the reference's convention
    (word bit b <-> polynomial coefficient 30-b, pager/bch_code.c:328-332), bit 31 even parity (ignored,
    pager_flex.c:1121)."""
    data21 &= 0x1FFFFF
    poly = 0
    for b in range(21):
        if (data21 >> b) & 1:
            poly |= 1 << (30 - b)
    rem = poly
    for bit in range(30, 9, -1):
        if rem & (1 << bit):
            rem ^= 0x769 << (bit - 10)
    word = data21
    for j in range(10):
        if (rem >> j) & 1:
            word |= 1 << (30 - j)
    if bin(word).count("1") & 1:
        word |= 1 << 31
    return word


def checksummed(data21: int) -> int:
    """Pick bits 0-3 so that the six-nibble sum of bits 0-20 is 0xf (pager_flex.c:108-119)."""
    data21 &= 0x1FFFF0
    s = 0
    v = data21
    for _ in range(6):
        s += v & 0xF
        v >>= 4
    return data21 | ((0xF - s) & 0xF)


def fiw_word(cycle: int, frame: int, roaming: int = 0, repeat: int = 0, traffic: int = 0) -> int:
    """pager_flex.c:1337-1338 and pager_flex_priv.h:384-388."""
    return bch_word(checksummed((cycle & 0xF) << 4 | (frame & 0x7F) << 8 | (roaming & 1) << 15 | (repeat & 1) << 16 |
                                (traffic & 0xF) << 17))


def biw_word(eob: int, vsw: int, prio: int = 0, carry: int = 0, collapse: int = 0) -> int:
    """Block information word: prio bits 4-7, end-of-block 8-9, vector start word 10-15 (pager_flex.c:1141-1142)."""
    return bch_word(checksummed((prio & 0xF) << 4 | (eob & 3) << 8 | (vsw & 0x3F) << 10 | (carry & 3) << 16 |
                                (collapse & 7) << 18))


def short_address(capcode: int) -> list[int]:
    """pager_flex.c:551-556: capcode = word - 32768 for words in (0x8000, 0x1e0000]."""
    w = capcode + 32768
    assert 0x8000 < w <= 0x1E0000
    return [bch_word(w)]


def long_address(first: int, second: int) -> list[int]:
    """Two-word address; the reference reports 0x1f9001 + ((0x1fffff - second) * 32768 + first - 1) mod 2^32
    (pager_flex.c:557-568).  `first` must not look like a short address."""
    assert not (0x8000 < first <= 0x1E0000 or 0x1F0000 < first < 0x1F7FFF)
    return [bch_word(first), bch_word(second)]


def long_capcode(first: int, second: int) -> int:
    return (0x1F9001 + ((((0x1FFFFF - second) * 32768) & 0xFFFFFFFF) + first - 1)) & 0xFFFFFFFF


def vector_alpha(start: int, length: int) -> int:
    """type 5 in bits 4-6, start word bits 7-13, length bits 14-20 (pager_flex.c:974-1008)."""
    return bch_word(checksummed(5 << 4 | (start & 0x7F) << 7 | (length & 0x7F) << 14))


def vector_numeric(start: int, length: int) -> int:
    """type 3, start word bits 7-13, (length - 1) in bits 14-16 (pager_flex.c:991-997)."""
    return bch_word(checksummed(3 << 4 | (start & 0x7F) << 7 | ((length - 1) & 7) << 14))


def vector_tone(digits: str = "", short_type: int = 0) -> int:
    """type 2; short type in bits 7-8, three digits in bits 9-20 (pager_flex.c:846-855)."""
    v = 2 << 4 | (short_type & 3) << 7
    for i, ch in enumerate(digits[:3]):
        v |= NUM_LUT.index(ch) << (9 + 4 * i)
    return bch_word(checksummed(v))


def vector_siv(siv_type: int, data: int) -> int:
    """type 1; instruction type bits 7-9, data bits 10-20 (pager_flex.c:896-909)."""
    return bch_word(checksummed(1 << 4 | (siv_type & 7) << 7 | (data & 0x7FF) << 10))


def alpha_words(text: str, seq: int = 3, fragment: bool = False, maildrop: bool = False, signature: int = 0x55) -> list[int]:
    """Header word (fragment flag bit 10, sequence bits 11-12, maildrop bit 20) followed by words of three
    7-bit characters; in a first fragment (seq 3) the low 7 bits of the first content word are a signature the
    reference skips; 0x03 ends a word's characters (pager_flex.c:631-669)."""
    head = (1 << 10 if fragment else 0) | (seq & 3) << 11 | (1 << 20 if maildrop else 0)
    chars = [ord(c) & 0x7F for c in text]
    if seq == 3:
        chars = [signature & 0x7F] + chars
    while len(chars) % 3:
        chars.append(0x03)
    out = [bch_word(head)]
    for i in range(0, len(chars), 3):
        out.append(bch_word(chars[i] | chars[i + 1] << 7 | chars[i + 2] << 14))
    return out


def numeric_words(digits: str) -> list[int]:
    """Standard numeric body: the first word carries 2 header bits then 19 digit bits, later words 21 bits, digits
    LSB-first nibbles (pager_flex.c:729-817)."""
    bits = [0, 0]
    for ch in digits:
        v = NUM_LUT.index(ch)
        bits += [(v >> b) & 1 for b in range(4)]
    while len(bits) % 21:
        bits.append(0)
    out = []
    for i in range(0, len(bits), 21):
        w = 0
        for b, bit in enumerate(bits[i:i + 21]):
            w |= bit << b
        out.append(bch_word(w))
    return out


def build_phase(messages, extra_biw: int = 0, idle_fill: int = 0x1FFFFF) -> np.ndarray:
    """One phase of 88 words.  messages: list of dicts
         {"addr": [address words], "kind": "alpha"|"numeric"|"tone"|"siv", ...kind specific...}
       alpha: text, seq, fragment, maildrop;  numeric: digits;  tone: digits (<= 3), second (optional 5 more);
       siv: siv_type, data.
    Layout parsed by the reference (pager_flex.c:1141-1194): word 0 BIW, [1, 1+eob) extra BIWs, then the address
    words, then one vector per address starting at vsw (a long address owns two vector words, the second being
    the message's first word), then message bodies."""
    words = [0] * 88
    eob = extra_biw
    addr_start = 1 + eob
    n_addr_words = sum(len(m["addr"]) for m in messages)
    vsw = addr_start + n_addr_words
    body = vsw + n_addr_words                  # first free word after the vectors
    for i in range(1, 1 + eob):
        words[i] = bch_word(checksummed(1 << 4 | (i & 0x1F) << 7))        # "date" BIW, informational only
    ai, vi = addr_start, vsw
    for m in messages:
        long = len(m["addr"]) == 2
        for w in m["addr"]:
            words[ai] = w
            ai += 1
        kind = m["kind"]
        if kind == "alpha":
            mw = alpha_words(m["text"], m.get("seq", 3), m.get("fragment", False), m.get("maildrop", False))
            if long:                            # the second vector word is the header word; the body follows
                words[vi] = vector_alpha(body, len(mw))
                words[vi + 1] = mw[0]
                mw = mw[1:]
            else:
                words[vi] = vector_alpha(body, len(mw))
        elif kind == "numeric":
            mw = numeric_words(m["digits"])
            if long:
                words[vi] = vector_numeric(body, len(mw))
                words[vi + 1] = mw[0]
                mw = mw[1:]
            else:
                words[vi] = vector_numeric(body, len(mw))
        elif kind == "tone":
            words[vi] = vector_tone(m.get("digits", ""), m.get("short_type", 0))
            if long:
                v = 0
                for i, ch in enumerate(m.get("second", "")[:5]):
                    v |= NUM_LUT.index(ch) << (4 * i)
                words[vi + 1] = bch_word(v)
            mw = []
        elif kind == "siv":
            words[vi] = vector_siv(m["siv_type"], m["data"])
            if long:
                words[vi + 1] = bch_word(0)
            mw = []
        else:
            raise ValueError(kind)
        vi += len(m["addr"])
        assert body + len(mw) <= 88, "phase overflow"
        for w in mw:
            words[body] = w
            body += 1
    words[0] = biw_word(eob, vsw)
    for i in range(body, 88):
        words[i] = bch_word(idle_fill)
    return np.asarray(words, dtype=np.uint64)


# ---------------------------------------------------------------------------------------------
# symbols and waveform
# ---------------------------------------------------------------------------------------------
def _bits_msb(value: int, n: int) -> list[int]:
    return [(value >> (n - 1 - i)) & 1 for i in range(n)]


def frame_levels(coding: str, cycle: int, frame: int, phases: dict, lead_in: int = 40) -> np.ndarray:
    """Per-sample levels (float, full deviation = +-1) of one frame at 16 kHz.
    phases: {"A": words88, ...} for the coding's phases (missing ones are filled with idle words)."""
    cd = CODINGS[coding]
    lv: list[float] = []

    def put_2fsk(bits, sps):                    # bit 1 -> positive sample (pager_flex.c:137)
        for b in bits:
            lv.extend([1.0 if b else -1.0] * sps)

    # Sync 1 at 1600 bit/s: quiet lead-in of zeros, BS1, A, B, inverted A, FIW (LSB first)
    a = (cd["seq_a"] << 16) | SYNC_MAGIC_A
    put_2fsk([0] * lead_in, 10)
    put_2fsk(_bits_msb(SYNC_BS1, 32), 10)
    put_2fsk(_bits_msb(a, 32), 10)
    put_2fsk(_bits_msb(SYNC_MAGIC_B, 16), 10)
    put_2fsk(_bits_msb(~a & 0xFFFFFFFF, 32), 10)
    fiw = fiw_word(cycle, frame)
    put_2fsk([(fiw >> i) & 1 for i in range(32)], 10)

    sps, sym_bits = cd["sps"], cd["sym_bits"]
    # 4-FSK symbol value s = 2*hi + lo -> level (pager_flex.c:149-170): 0 far negative, 1 near negative,
    # 2 far positive, 3 near positive
    lvl4 = {0: -1.0, 1: -1.0 / 3.0, 2: 1.0, 3: 1.0 / 3.0}

    def put_symbols(bits):
        if sym_bits == 1:
            put_2fsk(bits, sps)
        else:
            for i in range(0, len(bits), 2):
                lv.extend([lvl4[2 * bits[i] + bits[i + 1]]] * sps)

    n_comma = cd["sync2"]
    dots = [(i + 1) & 1 for i in range(n_comma)]
    if sym_bits == 2:                           # comma symbols are not sliced by the reference: alternate far levels
        for d in dots:
            lv.extend([1.0 if d else -1.0] * sps)
    else:
        put_2fsk(dots, sps)
    put_symbols(_bits_msb(SYNC2_C, 16))
    if sym_bits == 2:
        for d in dots:
            lv.extend([-1.0 if d else 1.0] * sps)
    else:
        put_2fsk([1 - d for d in dots], sps)
    put_symbols(_bits_msb(~SYNC2_C & 0xFFFF, 16))

    # block: 11 groups of 8 words per phase, bit-interleaved LSB first (pager_flex.c:1201-1222);
    # symbol -> phase routing per pager_flex.c:1242-1285
    idle = build_phase([])
    ph = {name: np.asarray(phases.get(name, idle), dtype=np.uint64) for name in cd["phases"]}

    def phase_bits(words):
        out = []
        for g in range(11):
            for b in range(32):
                for w in range(8):
                    out.append(int(words[8 * g + w] >> b) & 1)
        return out

    pb = {k: phase_bits(v) for k, v in ph.items()}
    names = cd["phases"]
    nsym = cd["symbols"]
    if coding == "1600/2":
        put_2fsk(pb["A"], sps)
    elif coding == "3200/2":
        bits = []
        for i in range(nsym // 2):
            bits += [pb["A"][i], pb["C"][i]]
        put_2fsk(bits, sps)
    elif coding == "3200/4":
        for i in range(nsym):
            lv.extend([lvl4[2 * pb["A"][i] + pb["C"][i]]] * sps)
    else:
        for i in range(nsym // 2):
            lv.extend([lvl4[2 * pb["A"][i] + pb["B"][i]]] * sps)
            lv.extend([lvl4[2 * pb["C"][i] + pb["D"][i]]] * sps)
    assert names
    lv.extend([-1.0] * 40)
    return np.asarray(lv, dtype=np.float64)


def pcm_from_levels(levels: np.ndarray, amplitude: float = 6000.0, noise_sigma: float = 0.0, smooth: int = 0,
                    dc: float = 0.0, seed: int = 20260925, pad: int = 200) -> np.ndarray:
    x = np.concatenate([np.full(pad, -1.0), levels, np.full(pad, -1.0)])
    if smooth > 1:
        x = np.convolve(x, np.ones(smooth) / smooth, mode="same")
    x = amplitude * x + dc
    if noise_sigma > 0:
        x = x + np.random.default_rng(seed).normal(0.0, noise_sigma, len(x))
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)
