"""ctypes mirror of include/tslb200_gpupager.h (resampler + POCSAG decode for every channel)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

F_DC_BLOCK, F_KEEP_PCM, F_NO_RESAMPLE, F_INVERT = 0x1, 0x2, 0x4, 0x8
DECODER_POCSAG, DECODER_FLEX = 0, 1
MSG_NUMERIC, MSG_ALPHA, MSG_FLEX_ALNUM, MSG_FLEX_NUM, MSG_FLEX_SIV = 0, 1, 2, 3, 4


class GpuPagerError(RuntimeError):
    def __init__(self, code, where):
        msg = _lib.lib().gpupager_last_error()
        super().__init__(f"{where} failed: {code} ({msg.decode() if msg else ''})")
        self.code = code


def _check(code, where):
    if code != 0:
        raise GpuPagerError(code, where)


def quantize_taps(coeffs):
    coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
    out = np.zeros(len(coeffs), np.int16)
    _check(_lib.lib().gpupager_quantize_taps(coeffs.ctypes.data, len(coeffs), out.ctypes.data), "gpupager_quantize_taps")
    return out


class GpuPager:
    def __init__(self, nr_channels, max_feed_samples, taps_q14=None, interpolate=1, decimate=1, device=0, flags=0,
                 dc_pole=0.9999, decoder=DECODER_POCSAG, channel_map=None):
        L = self._L = _lib.lib()
        cfg = _lib.GpuPagerCfg()
        cfg.struct_size = C.sizeof(_lib.GpuPagerCfg)
        cfg.nr_channels = int(nr_channels)
        cfg.device = int(device)
        cfg.interpolate = int(interpolate)
        cfg.decimate = int(decimate)
        self._taps = None if taps_q14 is None else np.ascontiguousarray(taps_q14, dtype=np.int16)
        cfg.nr_taps = 0 if self._taps is None else len(self._taps)
        cfg.max_feed_samples = int(max_feed_samples)
        cfg.flags = int(flags)
        cfg.dc_pole = float(dc_pole)
        cfg.taps = None if self._taps is None else self._taps.ctypes.data_as(C.POINTER(C.c_int16))
        cfg.decoder = int(decoder)
        self._map = None if channel_map is None else np.ascontiguousarray(channel_map, dtype=np.uint32)
        if self._map is not None:
            assert len(self._map) == int(nr_channels)
            cfg.channel_map = self._map.ctypes.data_as(C.POINTER(C.c_uint32))
        self._h = C.c_void_p()
        _check(L.gpupager_create(C.byref(self._h), C.byref(cfg)), "gpupager_create")
        self.nr_channels = int(nr_channels)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.gpupager_destroy(C.byref(self._h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def feed(self, pcm: np.ndarray):
        """pcm: [nr_channels, n] int16 in host memory."""
        rows = self.nr_channels if self._map is None else int(self._map.max()) + 1
        assert pcm.dtype == np.int16 and pcm.ndim == 2 and pcm.shape[0] >= rows
        pcm = np.ascontiguousarray(pcm)
        _check(self._L.gpupager_feed(self._h, pcm.ctypes.data, pcm.shape[1], pcm.shape[1]), "gpupager_feed")

    def feed_device(self, dev_ptr, pitch, n, stream=0):
        _check(self._L.gpupager_feed_device(self._h, dev_ptr, pitch, n, stream), "gpupager_feed_device")

    def poll(self, cap=4096):
        arr = (_lib.GpuPagerMsg * cap)()
        n = C.c_size_t(0)
        _check(self._L.gpupager_poll(self._h, arr, cap, C.byref(n)), "gpupager_poll")
        out = []
        for i in range(n.value):
            m = arr[i]
            raw = C.string_at(C.addressof(m) + _lib.GpuPagerMsg.text.offset, min(m.len, 512))
            out.append((m.channel, m.kind, m.baud, m.capcode, m.function, m.len, raw))
        return out

    def poll_full(self, cap=4096):
        """Records in the oracle's tuple shape: (channel, (kind, baud, capcode64, function, len, aux, text))."""
        arr = (_lib.GpuPagerMsg * cap)()
        n = C.c_size_t(0)
        _check(self._L.gpupager_poll(self._h, arr, cap, C.byref(n)), "gpupager_poll")
        out = []
        for i in range(n.value):
            m = arr[i]
            raw = C.string_at(C.addressof(m) + _lib.GpuPagerMsg.text.offset, min(m.len, 512))
            out.append((m.channel, (m.kind, m.baud, m.capcode | (m.capcode_hi << 32), m.function, m.len, tuple(m.aux), raw)))
        return out

    def dispatch_flex(self):
        """Fires the three C callbacks of pager/pager_flex.h and returns what they received, as oracle-shaped tuples."""
        got = []

        def on_alnum(user, channel, baud, phase, cycle, frame, cap, fragmented, maildrop, seq, data, length):
            got.append((channel, (2, baud, cap, phase, length, (cycle, frame, fragmented, maildrop, seq, 0), C.string_at(data, length))))
            return 0

        def on_num(user, channel, baud, phase, cycle, frame, cap, data, length):
            got.append((channel, (3, baud, cap, phase, length, (cycle, frame, 0, 0, 0, 0), C.string_at(data, length))))
            return 0

        def on_siv(user, channel, baud, phase, cycle, frame, cap, siv_type, data):
            got.append((channel, (4, baud, cap, phase, 0, (cycle, frame, siv_type, data, 0, 0), b"")))
            return 0
        a, b, c = _lib.ON_FLEX_ALNUM(on_alnum), _lib.ON_FLEX_NUM(on_num), _lib.ON_FLEX_SIV(on_siv)
        n = C.c_size_t(0)
        _check(self._L.gpupager_dispatch_flex(self._h, a, b, c, None, C.byref(n)), "gpupager_dispatch_flex")
        assert n.value == len(got)
        return got

    def dispatch(self):
        """Fires the C callbacks (same argument meaning as pager/pager_pocsag.h) and returns what they received."""
        got = []

        def mk(kind):
            def cb(user, channel, baud, capcode, data, length, function):
                got.append((channel, kind, baud, capcode, function, length, C.string_at(data, length)))
                return 0
            return _lib.ON_MSG(cb)
        on_num, on_alpha = mk(0), mk(1)
        n = C.c_size_t(0)
        _check(self._L.gpupager_dispatch(self._h, on_num, on_alpha, None, C.byref(n)), "gpupager_dispatch")
        assert n.value == len(got)
        return got

    def collect_pcm(self, cap):
        out = np.zeros((self.nr_channels, cap), np.int16)
        n = C.c_size_t(0)
        _check(self._L.gpupager_collect_pcm(self._h, out.ctypes.data, cap, C.byref(n)), "gpupager_collect_pcm")
        return out[:, :n.value]

    @property
    def kernel_launches(self):
        return self._L.gpupager_kernel_launches(self._h)

    @property
    def dropped(self):
        return self._L.gpupager_dropped_msgs(self._h)


class GpuMM:
    """ctypes mirror of gpumm_* (Mueller-Muller timing recovery, pager/mueller_muller.c) for nr_channels streams."""

    def __init__(self, nr_channels, kw, km, samples_per_bit, error_min, error_max, max_feed_samples, fma=True, device=0):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        self.nr_channels = int(nr_channels)
        self.max_feed = int(max_feed_samples)
        _check(self._L.gpumm_create(C.byref(self._h), self.nr_channels, device, kw, km, samples_per_bit, error_min, error_max,
                                    self.max_feed, 1 if fma else 0), "gpumm_create")

    def process(self, pcm: np.ndarray):
        """pcm [nr_channels, n] int16 -> list of per-channel decision arrays"""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        n = pcm.shape[1]
        out = np.zeros((self.nr_channels, max(n, 1)), np.int16)
        nr = np.zeros(self.nr_channels, np.uint32)
        _check(self._L.gpumm_process(self._h, pcm.ctypes.data, n, n, out.ctypes.data, out.shape[1], nr.ctypes.data), "gpumm_process")
        return [out[c, :nr[c]].copy() for c in range(self.nr_channels)]

    def state(self, channel):
        st = np.zeros(4, np.float32)
        _check(self._L.gpumm_get_state(self._h, channel, st.ctypes.data), "gpumm_get_state")
        return st

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.gpumm_destroy(C.byref(self._h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
