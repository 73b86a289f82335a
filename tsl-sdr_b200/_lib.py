"""Loader for the C-ABI CUDA library.  No fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TSLB200_LIB") or os.path.join(PKG_DIR, "libtslb200.so")     # TSLB200_LIB: diagnostics builds only
_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is deliberately no CPU or PyTorch fallback.")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


class GpuChanCfg(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("sample_rate_hz", C.c_uint32), ("decimation", C.c_uint32),
                ("nr_taps", C.c_uint32), ("nr_channels", C.c_uint32), ("device", C.c_int32),
                ("max_batch_samples", C.c_uint32), ("flags", C.c_uint32), ("engine", C.c_uint32),
                ("reserved", C.c_uint32), ("lpf_taps", C.POINTER(C.c_double)),
                ("offset_hz", C.POINTER(C.c_int32)), ("gain", C.POINTER(C.c_double))]


def _declare(L: C.CDLL) -> None:
    vp, sz = C.c_void_p, C.c_size_t
    L.gpuchan_prepare_taps.argtypes = [vp, sz, C.c_int32, C.c_uint32, C.c_double, vp, vp]
    L.gpuchan_derot_increment.argtypes = [C.c_int32, C.c_uint32, C.c_uint32, vp]
    L.gpuchan_db_to_gain.restype = C.c_double
    L.gpuchan_db_to_gain.argtypes = [C.c_double]
    L.gpuchan_create.argtypes = [C.POINTER(vp), C.POINTER(GpuChanCfg)]
    L.gpuchan_destroy.argtypes = [C.POINTER(vp)]
    L.gpuchan_submit.argtypes = [vp, vp, sz]
    L.gpuchan_submit_device.argtypes = [vp, vp, sz, vp]
    L.gpuchan_submit_bytes.argtypes = [vp, vp, sz, C.c_uint32]
    L.gpuchan_sync.argtypes = [vp]
    L.gpuchan_stream_wait.argtypes = [vp, vp]
    L.gpuchan_pending.argtypes = [vp, C.POINTER(sz)]
    L.gpuchan_collect.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.gpuchan_collect_iq.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.gpuchan_device_pcm.argtypes = [vp, C.POINTER(vp), C.POINTER(sz), C.POINTER(sz)]
    L.gpuchan_get_taps.argtypes = [vp, C.c_uint32, vp, vp]
    L.gpuchan_get_rot_state.argtypes = [vp, C.c_uint32, vp, vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                        C.POINTER(C.c_uint32)]
    L.gpuchan_engine.argtypes = [vp]
    L.gpuchan_kernel_launches.restype = C.c_uint64
    L.gpuchan_kernel_launches.argtypes = [vp]
    L.gpuchan_last_error.restype = C.c_char_p
    L.gpupager_quantize_taps.argtypes = [vp, sz, vp]
    L.gpupager_create.argtypes = [C.POINTER(vp), C.POINTER(GpuPagerCfg)]
    L.gpupager_destroy.argtypes = [C.POINTER(vp)]
    L.gpupager_feed_device.argtypes = [vp, vp, sz, sz, vp]
    L.gpupager_feed.argtypes = [vp, vp, sz, sz]
    L.gpupager_dispatch.argtypes = [vp, ON_MSG, ON_MSG, vp, C.POINTER(sz)]
    L.gpupager_poll.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.gpupager_dispatch_flex.argtypes = [vp, ON_FLEX_ALNUM, ON_FLEX_NUM, ON_FLEX_SIV, vp, C.POINTER(sz)]
    L.gpupager_collect_pcm.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.gpupager_kernel_launches.restype = C.c_uint64
    L.gpupager_kernel_launches.argtypes = [vp]
    L.gpupager_dropped_msgs.restype = C.c_uint64
    L.gpupager_dropped_msgs.argtypes = [vp]
    L.gpupager_last_error.restype = C.c_char_p
    L.gpumm_create.argtypes = [C.POINTER(vp), C.c_uint32, C.c_int32] + [C.c_float] * 5 + [C.c_uint32, C.c_uint32]
    L.gpumm_process.argtypes = [vp, vp, sz, sz, vp, sz, vp]
    L.gpumm_get_state.argtypes = [vp, C.c_uint32, vp]
    L.gpumm_destroy.argtypes = [C.POINTER(vp)]
    L.gpuchan_in_flight.argtypes = [vp]
    L.gpuchan_host_alloc.argtypes = [C.POINTER(vp), sz]
    L.gpuchan_host_free.argtypes = [vp]
    L.gpuchan_tc_selftest.argtypes = [vp, vp, vp, vp] + [C.c_int] * 9 + [vp]
    L.gpuchan_tc_plan_query.argtypes = [C.POINTER(GpuChanCfg), C.c_uint32, vp, vp, sz, vp, sz]
    L.gpuchan_math_selftest.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, vp]
    L.gpuchan_discard.argtypes = [vp]
    L.gpuchan_math_eval.argtypes = [vp, vp, sz, C.c_uint32, vp, vp]
    L.gpuchan_tc_model.argtypes = [vp, vp]
    L.gpuchan_collect_begin.argtypes = [vp, vp, sz]
    L.gpuchan_collect_end.argtypes = [vp, C.POINTER(sz)]
    L.gpuchan_multi_create.argtypes = [C.POINTER(vp), C.POINTER(GpuChanCfg), vp, C.c_uint32, C.c_uint32]
    L.gpuchan_multi_destroy.argtypes = [C.POINTER(vp)]
    L.gpuchan_multi_submit.argtypes = [vp, vp, sz]
    L.gpuchan_multi_pending.argtypes = [vp, C.POINTER(sz)]
    L.gpuchan_multi_collect.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.gpuchan_multi_discard.argtypes = [vp]
    L.gpuchan_multi_sync.argtypes = [vp]
    L.gpuchan_multi_devices.restype = C.c_uint32
    L.gpuchan_multi_devices.argtypes = [vp]
    L.gpuchan_multi_bank.argtypes = [vp, C.c_uint32, C.POINTER(vp), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.gpurelay_flags_bytes.restype = sz
    L.gpurelay_flags_bytes.argtypes = [C.c_uint32, C.c_uint32]
    L.gpurelay_create.argtypes = [C.POINTER(vp), C.POINTER(GpuRelayCfg)]
    L.gpurelay_destroy.argtypes = [C.POINTER(vp)]
    L.gpurelay_slot.argtypes = [vp, C.c_uint32, C.POINTER(vp)]
    L.gpurelay_export.argtypes = [vp, vp]
    L.gpurelay_connect_ipc.argtypes = [vp, vp]
    L.gpurelay_connect_local.argtypes = [vp, vp]
    L.gpurelay_acquire.argtypes = [vp, C.c_uint64, vp]
    L.gpurelay_advance.argtypes = [vp, C.c_uint64, sz, vp, C.POINTER(vp)]
    L.gpurelay_release.argtypes = [vp, C.c_uint64, vp]
    L.gpurelay_last_error.restype = C.c_char_p
    L.gpufm_create.argtypes = [C.POINTER(vp), C.c_int32, C.c_uint32, C.c_uint32]
    L.gpufm_process.argtypes = [vp, vp, sz, vp]
    L.gpufm_destroy.argtypes = [C.POINTER(vp)]
    L.gpuchan_timing_enable.argtypes = [vp, C.c_int]
    L.gpuchan_timing_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]


class GpuRelayCfg(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("rank", C.c_uint32), ("world", C.c_uint32), ("nr_slots", C.c_uint32),
                ("device", C.c_int32), ("reserved", C.c_uint32), ("slot_bytes", C.c_uint64), ("shm_name", C.c_char_p),
                ("flags_host", C.c_void_p)]


class GpuPagerCfg(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("nr_channels", C.c_uint32), ("device", C.c_int32),
                ("interpolate", C.c_uint32), ("decimate", C.c_uint32), ("nr_taps", C.c_uint32),
                ("max_feed_samples", C.c_uint32), ("flags", C.c_uint32), ("dc_pole", C.c_double),
                ("taps", C.POINTER(C.c_int16)), ("decoder", C.c_uint32), ("reserved", C.c_uint32),
                ("channel_map", C.POINTER(C.c_uint32))]


class GpuPagerMsg(C.Structure):
    _fields_ = [("channel", C.c_uint32), ("kind", C.c_uint32), ("baud", C.c_uint32), ("capcode", C.c_uint32),
                ("function", C.c_uint32), ("len", C.c_uint32), ("capcode_hi", C.c_uint32), ("aux", C.c_uint32 * 6),
                ("text", C.c_char * 512)]


ON_FLEX_ALNUM = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint16, C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint64, C.c_int,
                            C.c_int, C.c_uint8, C.POINTER(C.c_char), C.c_size_t)
ON_FLEX_NUM = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint16, C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint64,
                          C.POINTER(C.c_char), C.c_size_t)
ON_FLEX_SIV = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint16, C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint64, C.c_uint8,
                          C.c_uint32)
ON_MSG = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint16, C.c_uint32, C.POINTER(C.c_char), C.c_size_t, C.c_uint8)


# every symbol include/tslb200_gpuchan.h declares (checked by the CPU test-suite)
EXPORTS = ["gpuchan_prepare_taps", "gpuchan_derot_increment", "gpuchan_db_to_gain", "gpuchan_create",
           "gpuchan_destroy", "gpuchan_submit", "gpuchan_submit_device", "gpuchan_submit_bytes", "gpuchan_sync", "gpuchan_pending",
           "gpuchan_collect", "gpuchan_collect_iq", "gpuchan_device_pcm", "gpuchan_get_taps",
           "gpuchan_get_rot_state", "gpuchan_engine", "gpuchan_kernel_launches", "gpuchan_last_error", "gpuchan_timing_enable", "gpuchan_in_flight", "gpuchan_discard", "gpuchan_host_alloc", "gpuchan_host_free", "gpuchan_tc_selftest", "gpuchan_math_selftest", "gpuchan_tc_plan_query", "gpuchan_stream_wait",
           "gpuchan_timing_read", "gpuchan_tc_model", "gpuchan_math_eval", "gpuchan_collect_begin", "gpuchan_collect_end",
           "gpuchan_multi_create", "gpuchan_multi_destroy", "gpuchan_multi_submit", "gpuchan_multi_pending", "gpuchan_multi_collect",
           "gpuchan_multi_discard", "gpuchan_multi_sync", "gpuchan_multi_devices", "gpuchan_multi_bank",
           "gpurelay_flags_bytes", "gpurelay_create", "gpurelay_destroy", "gpurelay_slot", "gpurelay_export", "gpurelay_connect_ipc",
           "gpurelay_connect_local", "gpurelay_acquire", "gpurelay_advance", "gpurelay_release", "gpurelay_last_error", "gpufm_create", "gpufm_process", "gpufm_destroy",
           "gpupager_quantize_taps", "gpupager_create", "gpupager_destroy", "gpupager_feed_device", "gpupager_feed",
           "gpupager_dispatch", "gpupager_dispatch_flex", "gpupager_poll", "gpupager_collect_pcm", "gpupager_kernel_launches",
           "gpupager_dropped_msgs", "gpupager_last_error",
           "gpumm_create", "gpumm_process", "gpumm_get_state", "gpumm_destroy"]
