"""ctypes bindings for the test oracle -- TEST INFRASTRUCTURE ONLY.

Two libraries:
  * Oracle  -> oracle/liboracle.so : our CPU restatement (oracle/oracle.c)
  * Ref     -> oracle/_ref/libtslref*.so : the reference's own numeric objects,
               compiled unmodified from /root/reference by oracle/Makefile.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

_i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


class Msg(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("baud", C.c_uint32), ("capcode_lo", C.c_uint32),
                ("capcode_hi", C.c_uint32), ("function", C.c_uint32), ("len", C.c_uint32),
                ("aux", C.c_uint32 * 6), ("data", C.c_char * 520)]

    def as_tuple(self):
        raw = C.string_at(C.addressof(self) + Msg.data.offset, min(self.len, 520))
        return (self.kind, self.baud, self.capcode_lo | (self.capcode_hi << 32), self.function,
                self.len, tuple(self.aux), raw)


def build(force: bool = False) -> None:
    """Compile the checker (never the product): liboracle.so always, _ref when /root/reference exists."""
    need = force or not os.path.exists(os.path.join(HERE, "liboracle.so"))
    have_ref_src = os.path.exists("/root/reference/filter/direct_fir.c")
    if have_ref_src and not os.path.exists(os.path.join(HERE, "_ref", "libtslref.so")):
        need = True
    src_m = max(os.path.getmtime(os.path.join(HERE, f)) for f in ("oracle.c", "oracle.h", "ref_driver.c", "Makefile"))
    for lib in ("liboracle.so", os.path.join("_ref", "libtslref.so")):
        p = os.path.join(HERE, lib)
        if os.path.exists(p) and os.path.getmtime(p) < src_m and (have_ref_src or lib == "liboracle.so"):
            need = True
    if need:
        subprocess.run(["make", "-C", HERE, "all"], check=True, stdout=subprocess.DEVNULL)


def _as_i16(a):
    return np.ascontiguousarray(a, dtype=np.int16)


# --------------------------------------------------------------------------------------
class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = self.L = C.CDLL(path)
        L.orc_prepare_taps.argtypes = [_f64p, C.c_size_t, C.c_int32, C.c_uint32, C.c_double, _i16p, _i16p]
        L.orc_derot_incr.argtypes = [C.c_int32, C.c_uint32, C.c_uint, _i16p]
        L.orc_db_to_gain.restype = C.c_double
        L.orc_db_to_gain.argtypes = [C.c_double]
        L.orc_fast_atan2f.restype = C.c_float
        L.orc_fast_atan2f.argtypes = [C.c_float, C.c_float, C.c_int]
        L.orc_fm_pcm.restype = C.c_int16
        L.orc_fm_pcm.argtypes = [C.c_int32, C.c_int32, C.c_int]
        L.orc_atan_table.argtypes = [np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")]
        L.orc_chan_state_init.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_uint]
        L.orc_chan_stream.restype = C.c_size_t
        L.orc_chan_stream.argtypes = [C.c_void_p, C.c_size_t, _i16p, _i16p, C.c_uint, C.c_void_p, C.c_size_t,
                                      C.c_int, C.c_void_p, C.c_void_p]
        L.orc_resamp_init.argtypes = [C.c_void_p, _i16p, C.c_size_t, C.c_uint, C.c_uint]
        L.orc_resamp_free.argtypes = [C.c_void_p]
        L.orc_resamp_stream.restype = C.c_size_t
        L.orc_resamp_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                        C.POINTER(C.c_size_t)]
        L.orc_bch_decode.argtypes = [C.POINTER(C.c_uint32)]
        L.orc_dc_blocker_init.argtypes = [C.c_void_p, C.c_double]
        L.orc_dc_blocker_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_pocsag_new.restype = C.c_void_p
        L.orc_pocsag_new.argtypes = [C.c_size_t]
        L.orc_pocsag_delete.argtypes = [C.c_void_p]
        L.orc_pocsag_on_pcm.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_pocsag_msgs.restype = C.c_size_t
        L.orc_pocsag_msgs.argtypes = [C.c_void_p, C.POINTER(C.POINTER(Msg))]
        L.orc_msg_size.restype = C.c_size_t
        assert L.orc_msg_size() == C.sizeof(Msg)
        L.orc_mm_init.argtypes = [C.c_void_p] + [C.c_float] * 5
        L.orc_mm_process.restype = C.c_size_t
        L.orc_mm_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        L.orc_flex_new.restype = C.c_void_p
        L.orc_flex_new.argtypes = [C.c_size_t]
        L.orc_flex_delete.argtypes = [C.c_void_p]
        L.orc_flex_on_pcm.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_flex_msgs.restype = C.c_size_t
        L.orc_flex_msgs.argtypes = [C.c_void_p, C.POINTER(C.POINTER(Msg))]

    # a1
    def prepare_taps(self, lpf, offset_hz, fs, gain=1.0):
        lpf = np.ascontiguousarray(lpf, dtype=np.float64)
        re = np.zeros(len(lpf), np.int16)
        im = np.zeros(len(lpf), np.int16)
        self.L.orc_prepare_taps(lpf, len(lpf), int(offset_hz), int(fs), float(gain), re, im)
        return re, im

    def derot_incr(self, offset_hz, fs, D):
        out = np.zeros(2, np.int16)
        self.L.orc_derot_incr(int(offset_hz), int(fs), int(D), out)
        return out

    def atan_table(self):
        t = np.zeros(257, np.float32)
        self.L.orc_atan_table(t)
        return t

    # a2-a4 one-shot over a whole stream (state optional for chunked use)
    def new_state(self, offset_hz, fs, D):
        st = (C.c_uint8 * 24)()
        self.L.orc_chan_state_init(st, int(offset_hz), int(fs), int(D))
        return st

    def chan_stream(self, st, c_re, c_im, D, iq, fma=1, want_iq=True):
        iq = _as_i16(iq)
        n = len(iq) // 2
        T = len(c_re)
        K = (n - T) // D + 1 if n >= T else 0
        out_iq = np.zeros(2 * K, np.int16)
        out_pcm = np.zeros(K, np.int16)
        k = self.L.orc_chan_stream(st, T, _as_i16(c_re), _as_i16(c_im), int(D), iq.ctypes.data, n, int(fma),
                                   out_iq.ctypes.data if want_iq else None, out_pcm.ctypes.data)
        assert k == K
        return out_iq, out_pcm

    def channel(self, lpf, offset_hz, fs, D, iq, gain=1.0, fma=1):
        re, im = self.prepare_taps(lpf, offset_hz, fs, gain)
        st = self.new_state(offset_hz, fs, D)
        return self.chan_stream(st, re, im, D, iq, fma)

    # a5
    def resample(self, taps_i16, I, D, pcm):
        r = (C.c_uint8 * 64)()
        taps_i16 = _as_i16(taps_i16)
        assert self.L.orc_resamp_init(r, taps_i16, len(taps_i16), int(I), int(D)) == 0
        pcm = _as_i16(pcm)
        cap = len(pcm) * I // D + 8
        out = np.zeros(cap, np.int16)
        consumed = C.c_size_t(0)
        n = self.L.orc_resamp_stream(r, pcm.ctypes.data, len(pcm), out.ctypes.data, cap, C.byref(consumed))
        self.L.orc_resamp_free(r)
        return out[:n].copy(), consumed.value

    def dc_block(self, pcm, pole=0.9999):
        out = _as_i16(pcm).copy()
        st = (C.c_uint8 * 16)()
        self.L.orc_dc_blocker_init(st, float(pole))
        self.L.orc_dc_blocker_apply(st, out.ctypes.data, len(out))
        return out

    def bch_decode(self, word):
        w = C.c_uint32(word)
        rc = self.L.orc_bch_decode(C.byref(w))
        return rc, w.value

    def pocsag(self, pcm, chunk=0, max_msgs=4096):
        pcm = _as_i16(pcm)
        h = self.L.orc_pocsag_new(max_msgs)
        n = len(pcm)
        chunk = chunk or n
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            self.L.orc_pocsag_on_pcm(h, pcm[s:e].ctypes.data, e - s)
        pm = C.POINTER(Msg)()
        cnt = self.L.orc_pocsag_msgs(h, C.byref(pm))
        msgs = [pm[i].as_tuple() for i in range(cnt)]
        self.L.orc_pocsag_delete(h)
        return msgs


    # f4
    def mm(self, pcm, kw, km, spb, emin, emax, chunk=0, fma=1):
        """Returns (decisions int16, final state (w, m, next_offset, last_sample) as float32)."""
        pcm = _as_i16(pcm)
        st = (C.c_float * 8)()
        self.L.orc_mm_init(st, kw, km, spb, emin, emax)
        n = len(pcm)
        chunk = chunk or n
        out = np.zeros(n + 16, np.int16)
        tot = 0
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            tot += self.L.orc_mm_process(st, pcm[s:e].ctypes.data, e - s, out[tot:].ctypes.data, len(out) - tot, int(fma))
        return out[:tot].copy(), np.array(list(st)[4:8], dtype=np.float32)

    # a8
    def flex(self, pcm, chunk=0, max_msgs=4096):
        pcm = _as_i16(pcm)
        h = self.L.orc_flex_new(max_msgs)
        n = len(pcm)
        chunk = chunk or n
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            self.L.orc_flex_on_pcm(h, pcm[s:e].ctypes.data, e - s)
        pm = C.POINTER(Msg)()
        cnt = self.L.orc_flex_msgs(h, C.byref(pm))
        msgs = [pm[i].as_tuple() for i in range(cnt)]
        self.L.orc_flex_delete(h)
        return msgs


# --------------------------------------------------------------------------------------
def _cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def ref_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libtslref.so"))


class Ref:
    """The reference's own objects. variant: 'fma' (canonical), 'nofma', 'native' (the reference's literal
    -march=native flags; only usable when this CPU has the build host's ISA extensions)."""

    def __init__(self, variant: str = "fma"):
        name = {"fma": "libtslref.so", "nofma": "libtslref_nofma.so", "native": "libtslref_native.so"}[variant]
        path = os.path.join(HERE, "_ref", name)
        if variant == "native":
            isa = os.path.join(HERE, "_ref", "native_isa.txt")
            need = set(open(isa).read().split()) if os.path.exists(isa) else {"missing"}
            alias = {"avx512vpopcntdq": "avx512_vpopcntdq", "avx512vbmi2": "avx512_vbmi2", "avx512vnni": "avx512_vnni",
                     "avx512bitalg": "avx512_bitalg", "avx512bf16": "avx512_bf16", "avx512fp16": "avx512_fp16"}
            have = _cpu_flags()
            if not all(alias.get(x, x) in have for x in need):
                raise OSError("native reference build not runnable on this CPU")
        if not os.path.exists(path):
            raise OSError(f"{path} missing (oracle/_ref is built from /root/reference by oracle/Makefile)")
        L = self.L = C.CDLL(path)
        self.variant = variant
        L.ref_prepare_taps.argtypes = [_f64p, C.c_size_t, C.c_int32, C.c_uint32, C.c_double, _i16p, _i16p]
        L.ref_db_to_gain.restype = C.c_double
        L.ref_db_to_gain.argtypes = [C.c_double]
        L.ref_chan_new.restype = C.c_void_p
        L.ref_chan_new.argtypes = [_f64p, C.c_size_t, C.c_int32, C.c_uint32, C.c_uint, C.c_double]
        L.ref_chan_new_taps.restype = C.c_void_p
        L.ref_chan_new_taps.argtypes = [C.c_size_t, _i16p, _i16p, C.c_uint, C.c_uint32, C.c_int32]
        L.ref_chan_delete.argtypes = [C.c_void_p]
        L.ref_chan_get_state.argtypes = [C.c_void_p, _i16p, _i16p, C.POINTER(C.c_uint32)]
        L.ref_chan_get_taps.argtypes = [C.c_void_p, _i16p, _i16p]
        L.ref_chan_run.restype = C.c_size_t
        L.ref_chan_run.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
        L.ref_chan_flush.restype = C.c_size_t
        L.ref_chan_flush.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.ref_fm_demod.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_fast_atan2f.restype = C.c_float
        L.ref_fast_atan2f.argtypes = [C.c_float, C.c_float]
        L.ref_resamp_new.restype = C.c_void_p
        L.ref_resamp_new.argtypes = [_i16p, C.c_size_t, C.c_uint, C.c_uint, C.c_int, C.c_double]
        L.ref_resamp_delete.argtypes = [C.c_void_p]
        L.ref_resamp_run.restype = C.c_size_t
        L.ref_resamp_run.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ref_pocsag_new.restype = C.c_void_p
        L.ref_pocsag_new.argtypes = [C.c_uint32, C.c_size_t]
        L.ref_pocsag_delete.argtypes = [C.c_void_p]
        L.ref_pocsag_run.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
        L.ref_pocsag_msgs.restype = C.c_size_t
        L.ref_pocsag_msgs.argtypes = [C.c_void_p, C.POINTER(C.POINTER(Msg)), C.POINTER(C.c_size_t)]
        L.ref_decoder_pocsag_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.ref_msg_size.restype = C.c_size_t
        if hasattr(L, "ref_mm_run"):
            L.ref_mm_run.restype = C.c_size_t
            L.ref_mm_run.argtypes = [C.c_float] * 5 + [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_bch_decode.argtypes = [C.POINTER(C.c_uint32)]
        L.ref_flex_new.restype = C.c_void_p
        L.ref_flex_new.argtypes = [C.c_uint32, C.c_size_t]
        L.ref_flex_delete.argtypes = [C.c_void_p]
        L.ref_flex_run.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
        L.ref_flex_msgs.restype = C.c_size_t
        L.ref_flex_msgs.argtypes = [C.c_void_p, C.POINTER(C.POINTER(Msg)), C.POINTER(C.c_size_t)]
        L.ref_bench_multifm.restype = C.c_double
        L.ref_bench_multifm.argtypes = [C.c_size_t, _f64p, C.c_size_t, _i32p, C.c_uint32, C.c_uint, C.c_void_p,
                                        C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t)]
        assert L.ref_msg_size() == C.sizeof(Msg)

    def prepare_taps(self, lpf, offset_hz, fs, gain=1.0):
        lpf = np.ascontiguousarray(lpf, dtype=np.float64)
        re = np.zeros(len(lpf), np.int16)
        im = np.zeros(len(lpf), np.int16)
        self.L.ref_prepare_taps(lpf, len(lpf), int(offset_hz), int(fs), float(gain), re, im)
        return re, im

    def channel(self, lpf, offset_hz, fs, D, iq, gain=1.0, chunks=None, flush=True, return_state=False):
        """Whole-stream run through the real direct_fir + fm_demod, driven like demod_thread_process."""
        lpf = np.ascontiguousarray(lpf, dtype=np.float64)
        iq = _as_i16(iq)
        n = len(iq) // 2
        h = self.L.ref_chan_new(lpf, len(lpf), int(offset_hz), int(fs), int(D), float(gain))
        assert h
        cap = n // D + 16
        out_iq = np.zeros(2 * cap, np.int16)
        out_pcm = np.zeros(cap, np.int16)
        pos = 0
        got = 0
        for c in (chunks or [n]):
            c = min(c, n - pos)
            if c <= 0:
                break
            k = self.L.ref_chan_run(h, iq[2 * pos:].ctypes.data, c, out_iq[2 * got:].ctypes.data,
                                    out_pcm[got:].ctypes.data, cap - got)
            pos += c
            got += k
        if flush:
            got += self.L.ref_chan_flush(h, out_iq[2 * got:].ctypes.data, out_pcm[got:].ctypes.data, cap - got)
        state = None
        if return_state:
            rot = np.zeros(2, np.int16)
            incr = np.zeros(2, np.int16)
            cnt = C.c_uint32(0)
            self.L.ref_chan_get_state(h, rot, incr, C.byref(cnt))
            state = (rot, incr, cnt.value)
        self.L.ref_chan_delete(h)
        if return_state:
            return out_iq[:2 * got].copy(), out_pcm[:got].copy(), state
        return out_iq[:2 * got].copy(), out_pcm[:got].copy()

    def fm_demod(self, y_iq):
        y_iq = _as_i16(y_iq)
        n = len(y_iq) // 2
        out = np.zeros(n, np.int16)
        self.L.ref_fm_demod(y_iq.ctypes.data, n, out.ctypes.data)
        return out

    def fast_atan2f(self, y, x):
        return self.L.ref_fast_atan2f(float(y), float(x))

    def resample(self, taps_i16, I, D, pcm, use_dc=False, pole=0.9999):
        taps_i16 = _as_i16(taps_i16)
        pcm = _as_i16(pcm)
        h = self.L.ref_resamp_new(taps_i16, len(taps_i16), int(I), int(D), int(use_dc), float(pole))
        cap = len(pcm) * I // D + 1024
        out = np.zeros(cap, np.int16)
        n = self.L.ref_resamp_run(h, pcm.ctypes.data, len(pcm), out.ctypes.data, cap)
        self.L.ref_resamp_delete(h)
        return out[:n].copy()

    def bch_decode(self, word):
        w = C.c_uint32(word)
        rc = self.L.ref_bch_decode(C.byref(w))
        return rc, w.value

    def _collect(self, getter, h):
        pm = C.POINTER(Msg)()
        dropped = C.c_size_t(0)
        cnt = getter(h, C.byref(pm), C.byref(dropped))
        assert dropped.value == 0
        return [pm[i].as_tuple() for i in range(cnt)]

    def pocsag(self, pcm, chunk=0, max_msgs=4096):
        pcm = _as_i16(pcm)
        h = self.L.ref_pocsag_new(0, max_msgs)
        self.L.ref_pocsag_run(h, pcm.ctypes.data, len(pcm), chunk)
        msgs = self._collect(self.L.ref_pocsag_msgs, h)
        self.L.ref_pocsag_delete(h)
        return msgs

    def decoder_pocsag(self, taps_i16, I, D, pcm, max_msgs=4096):
        """decoder -m POCSAG: 1024-sample FIFO buffers -> polyphase_fir -> pager_pocsag_on_pcm."""
        taps_i16 = _as_i16(taps_i16)
        pcm = _as_i16(pcm)
        r = self.L.ref_resamp_new(taps_i16, len(taps_i16), int(I), int(D), 0, 0.9999)
        p = self.L.ref_pocsag_new(0, max_msgs)
        self.L.ref_decoder_pocsag_run(r, p, pcm.ctypes.data, len(pcm))
        msgs = self._collect(self.L.ref_pocsag_msgs, p)
        self.L.ref_pocsag_delete(p)
        self.L.ref_resamp_delete(r)
        return msgs

    def mm(self, pcm, kw, km, spb, emin, emax, chunk=0):
        pcm = _as_i16(pcm)
        out = np.zeros(len(pcm) + 16, np.int16)
        st = np.zeros(4, np.float32)
        n = self.L.ref_mm_run(kw, km, spb, emin, emax, pcm.ctypes.data, len(pcm), chunk, out.ctypes.data, len(out), st.ctypes.data)
        return out[:n].copy(), st

    def flex(self, pcm, chunk=0, max_msgs=4096):
        pcm = _as_i16(pcm)
        h = self.L.ref_flex_new(0, max_msgs)
        self.L.ref_flex_run(h, pcm.ctypes.data, len(pcm), chunk)
        msgs = self._collect(self.L.ref_flex_msgs, h)
        self.L.ref_flex_delete(h)
        return msgs

    def bench_multifm(self, lpf, offsets_hz, fs, D, iq, reps=1):
        """Reference threading model (one pthread per channel) timed over in-memory buffers.
        Returns (seconds, total_pcm_outputs)."""
        lpf = np.ascontiguousarray(lpf, dtype=np.float64)
        offs = np.ascontiguousarray(offsets_hz, dtype=np.int32)
        iq = _as_i16(iq)
        total = C.c_size_t(0)
        secs = self.L.ref_bench_multifm(len(offs), lpf, len(lpf), offs, int(fs), int(D), iq.ctypes.data,
                                        len(iq) // 2, int(reps), C.byref(total))
        return secs, total.value
