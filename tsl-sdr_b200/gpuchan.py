"""ctypes mirror of include/tslb200_gpuchan.h (host-side view of the channel bank)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

F_ATAN_FMA = 0x1
F_KEEP_IQ = 0x2
FMT_CS8, FMT_CU8, FMT_CU8_RTL = 1, 2, 3      # gpuchan_submit_bytes formats (include/tslb200_gpuchan.h)
ENGINE_AUTO, ENGINE_IMAD, ENGINE_TC = 0, 1, 2


class GpuChanError(RuntimeError):
    def __init__(self, code, where):
        msg = _lib.lib().gpuchan_last_error()
        super().__init__(f"{where} failed: {code} ({msg.decode() if msg else ''})")
        self.code = code


def _check(code, where):
    if code != 0:
        raise GpuChanError(code, where)


def prepare_taps(lpf, offset_hz, fs, gain=1.0):
    lpf = np.ascontiguousarray(lpf, dtype=np.float64)
    re = np.zeros(len(lpf), np.int16)
    im = np.zeros(len(lpf), np.int16)
    _check(_lib.lib().gpuchan_prepare_taps(lpf.ctypes.data, len(lpf), int(offset_hz), int(fs), float(gain),
                                           re.ctypes.data, im.ctypes.data), "gpuchan_prepare_taps")
    return re, im


def derot_increment(offset_hz, fs, decimation):
    out = np.zeros(2, np.int16)
    _check(_lib.lib().gpuchan_derot_increment(int(offset_hz), int(fs), int(decimation), out.ctypes.data),
           "gpuchan_derot_increment")
    return out


def db_to_gain(db):
    return _lib.lib().gpuchan_db_to_gain(float(db))


class GpuChan:
    """One bank = all channels of a receiver on one GPU (replaces N demod threads)."""

    def __init__(self, lpf_taps, offsets_hz, sample_rate_hz, decimation, max_batch_samples, device=0,
                 gains=None, flags=F_ATAN_FMA, engine=ENGINE_AUTO):
        L = _lib.lib()
        self._L = L
        self._lpf = np.ascontiguousarray(lpf_taps, dtype=np.float64)
        self._offs = np.ascontiguousarray(offsets_hz, dtype=np.int32)
        self._gains = None if gains is None else np.ascontiguousarray(gains, dtype=np.float64)
        cfg = _lib.GpuChanCfg()
        cfg.struct_size = C.sizeof(_lib.GpuChanCfg)
        cfg.sample_rate_hz = int(sample_rate_hz)
        cfg.decimation = int(decimation)
        cfg.nr_taps = len(self._lpf)
        cfg.nr_channels = len(self._offs)
        cfg.device = int(device)
        cfg.max_batch_samples = int(max_batch_samples)
        cfg.flags = int(flags)
        cfg.engine = int(engine)
        cfg.lpf_taps = self._lpf.ctypes.data_as(C.POINTER(C.c_double))
        cfg.offset_hz = self._offs.ctypes.data_as(C.POINTER(C.c_int32))
        cfg.gain = self._gains.ctypes.data_as(C.POINTER(C.c_double)) if self._gains is not None else None
        self._h = C.c_void_p()
        _check(L.gpuchan_create(C.byref(self._h), C.byref(cfg)), "gpuchan_create")
        self.nr_channels = len(self._offs)
        self.nr_taps = len(self._lpf)
        self.decimation = int(decimation)
        self.flags = int(flags)
        self.max_batch = int(max_batch_samples)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.gpuchan_destroy(C.byref(self._h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data path ----------------------------------------------------------------------
    def submit(self, iq_host: np.ndarray):
        """iq_host: interleaved int16 I,Q in host memory."""
        assert iq_host.dtype == np.int16 and iq_host.flags.c_contiguous
        _check(self._L.gpuchan_submit(self._h, iq_host.ctypes.data, len(iq_host) // 2), "gpuchan_submit")

    def submit_ptr(self, host_ptr: int, n_complex: int):
        _check(self._L.gpuchan_submit(self._h, host_ptr, n_complex), "gpuchan_submit")

    def submit_bytes(self, iq8: np.ndarray, fmt: int):
        """8-bit interleaved I,Q (uint8 view), widened on the device: FMT_CS8 / FMT_CU8 / FMT_CU8_RTL."""
        iq8 = np.ascontiguousarray(iq8).view(np.uint8)
        self._keep8 = iq8
        _check(self._L.gpuchan_submit_bytes(self._h, iq8.ctypes.data, len(iq8) // 2, int(fmt)), "gpuchan_submit_bytes")

    def submit_device(self, dev_ptr: int, n_complex: int, stream: int = 0):
        _check(self._L.gpuchan_submit_device(self._h, dev_ptr, n_complex, stream), "gpuchan_submit_device")

    def stream_wait(self, stream: int):
        _check(self._L.gpuchan_stream_wait(self._h, stream), "gpuchan_stream_wait")

    def sync(self):
        _check(self._L.gpuchan_sync(self._h), "gpuchan_sync")

    def pending(self) -> int:
        n = C.c_size_t(0)
        _check(self._L.gpuchan_pending(self._h, C.byref(n)), "gpuchan_pending")
        return n.value

    def collect(self, out: np.ndarray | None = None) -> np.ndarray:
        """PCM of the last submit as [nr_channels, n] int16."""
        k = self.pending()
        if out is None:
            out = np.zeros((self.nr_channels, max(k, 1)), np.int16)
        n = C.c_size_t(0)
        _check(self._L.gpuchan_collect(self._h, out.ctypes.data, out.shape[1], C.byref(n)), "gpuchan_collect")
        self._last_k = n.value
        return out[:, :n.value]

    def collect_into(self, host_ptr: int, cap_per_channel: int) -> int:
        n = C.c_size_t(0)
        _check(self._L.gpuchan_collect(self._h, host_ptr, cap_per_channel, C.byref(n)), "gpuchan_collect")
        self._last_k = n.value
        return n.value

    def collect_iq(self) -> np.ndarray:
        """Post-FIR IQ of the batch most recently returned by collect()."""
        k = getattr(self, "_last_k", 0)
        out = np.zeros((self.nr_channels, max(k, 1), 2), np.int16)
        n = C.c_size_t(0)
        _check(self._L.gpuchan_collect_iq(self._h, out.ctypes.data, out.shape[1], C.byref(n)), "gpuchan_collect_iq")
        return out[:, :n.value, :]

    def device_pcm(self):
        p, pitch, n = C.c_void_p(), C.c_size_t(0), C.c_size_t(0)
        _check(self._L.gpuchan_device_pcm(self._h, C.byref(p), C.byref(pitch), C.byref(n)), "gpuchan_device_pcm")
        return p.value, pitch.value, n.value

    # -- introspection ---------------------------------------------------------------------
    def taps(self, channel):
        re = np.zeros(self.nr_taps, np.int16)
        im = np.zeros(self.nr_taps, np.int16)
        _check(self._L.gpuchan_get_taps(self._h, channel, re.ctypes.data, im.ctypes.data), "gpuchan_get_taps")
        return re, im

    def rot_state(self, channel):
        rot = np.zeros(2, np.int16)
        incr = np.zeros(2, np.int16)
        k = C.c_uint64(0)
        mu = C.c_uint32(0)
        lam = C.c_uint32(0)
        _check(self._L.gpuchan_get_rot_state(self._h, channel, rot.ctypes.data, incr.ctypes.data, C.byref(k),
                                             C.byref(mu), C.byref(lam)), "gpuchan_get_rot_state")
        return rot, incr, k.value, mu.value, lam.value

    @property
    def engine(self):
        return self._L.gpuchan_engine(self._h)

    @property
    def kernel_launches(self):
        return self._L.gpuchan_kernel_launches(self._h)

    def timing_enable(self, on=True):
        _check(self._L.gpuchan_timing_enable(self._h, 1 if on else 0), "gpuchan_timing_enable")

    def timing_read(self):
        ms = C.c_double(0)
        n = C.c_uint64(0)
        _check(self._L.gpuchan_timing_read(self._h, C.byref(ms), C.byref(n)), "gpuchan_timing_read")
        return ms.value, n.value

    def tc_model(self):
        """(MMAs per tile, N per MMA, outputs per tile, channel groups) of the tensor-core engine; zeros on IMAD."""
        out = (C.c_uint64 * 4)()
        _check(self._L.gpuchan_tc_model(self._h, out), "gpuchan_tc_model")
        return tuple(int(v) for v in out)

    def discard(self):
        _check(self._L.gpuchan_discard(self._h), "gpuchan_discard")

    @property
    def in_flight(self):
        return self._L.gpuchan_in_flight(self._h)


FANOUT_HOST, FANOUT_RELAY = 0, 1


class GpuChanMulti:
    """One process, several GPUs (gpuchan_multi_*): all channels of a configuration sharded over `devices`."""

    def __init__(self, lpf_taps, offsets_hz, sample_rate_hz, decimation, max_batch_samples, devices, gains=None,
                 flags=F_ATAN_FMA, engine=ENGINE_AUTO, fanout=FANOUT_HOST):
        L = self._L = _lib.lib()
        self._lpf = np.ascontiguousarray(lpf_taps, dtype=np.float64)
        self._offs = np.ascontiguousarray(offsets_hz, dtype=np.int32)
        self._gains = None if gains is None else np.ascontiguousarray(gains, dtype=np.float64)
        self._devs = np.ascontiguousarray(devices, dtype=np.int32)
        cfg = _lib.GpuChanCfg()
        cfg.struct_size = C.sizeof(_lib.GpuChanCfg)
        cfg.sample_rate_hz = int(sample_rate_hz)
        cfg.decimation = int(decimation)
        cfg.nr_taps = len(self._lpf)
        cfg.nr_channels = len(self._offs)
        cfg.max_batch_samples = int(max_batch_samples)
        cfg.flags = int(flags)
        cfg.engine = int(engine)
        cfg.lpf_taps = self._lpf.ctypes.data_as(C.POINTER(C.c_double))
        cfg.offset_hz = self._offs.ctypes.data_as(C.POINTER(C.c_int32))
        cfg.gain = self._gains.ctypes.data_as(C.POINTER(C.c_double)) if self._gains is not None else None
        self._h = C.c_void_p()
        _check(L.gpuchan_multi_create(C.byref(self._h), C.byref(cfg), self._devs.ctypes.data, len(self._devs), int(fanout)),
               "gpuchan_multi_create")
        self.nr_channels = len(self._offs)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.gpuchan_multi_destroy(C.byref(self._h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def submit(self, iq_host: np.ndarray):
        assert iq_host.dtype == np.int16 and iq_host.flags.c_contiguous
        self._keep = iq_host
        _check(self._L.gpuchan_multi_submit(self._h, iq_host.ctypes.data, len(iq_host) // 2), "gpuchan_multi_submit")

    def pending(self) -> int:
        n = C.c_size_t(0)
        _check(self._L.gpuchan_multi_pending(self._h, C.byref(n)), "gpuchan_multi_pending")
        return n.value

    def collect(self) -> np.ndarray:
        k = self.pending()
        out = np.zeros((self.nr_channels, max(k, 1)), np.int16)
        n = C.c_size_t(0)
        _check(self._L.gpuchan_multi_collect(self._h, out.ctypes.data, out.shape[1], C.byref(n)), "gpuchan_multi_collect")
        return out[:, :n.value]

    def sync(self):
        _check(self._L.gpuchan_multi_sync(self._h), "gpuchan_multi_sync")

    @property
    def devices(self):
        return self._L.gpuchan_multi_devices(self._h)

    def bank_range(self, i):
        first, cnt = C.c_uint32(0), C.c_uint32(0)
        _check(self._L.gpuchan_multi_bank(self._h, i, None, C.byref(first), C.byref(cnt)), "gpuchan_multi_bank")
        return first.value, cnt.value
