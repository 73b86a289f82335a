"""tsl-sdr_b200 -- B200-native multifm channel bank (file_if -> FIR/mix/decimate -> FM -> pager path).

The directory name carries a hyphen (it mirrors the upstream project name), so import it through
``load_package()`` in the repository root's ``tslb200_loader.py``; it registers this package as
``tsl_sdr_b200``.  The product is the C-ABI shared library ``libtslb200.so`` (CUDA, sm_100a);
this package is only the thin ctypes mirror used by tests and bench.py.
"""
from . import _lib          # noqa: F401
from .gpuchan import GpuChan, prepare_taps, derot_increment, db_to_gain  # noqa: F401
from .gpupager import GpuPager, quantize_taps  # noqa: F401
