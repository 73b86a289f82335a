"""ctypes mirror of include/tslb200_gpurelay.h for one-process-per-GPU launches (torchrun): the IQ batch enters the ingest
GPU (rank 0) and travels down the chain rank 0 -> 1 -> ... -> N-1 over NVLink on the copy engines.  torch.distributed is
used once, at start-up, to hand every rank the CUDA IPC handles of its predecessor's slots; the data path itself has no
collective and no host synchronisation (stream-ordered counters in a shared-memory segment)."""
from __future__ import annotations

import ctypes as C

from . import _lib


class RelayError(RuntimeError):
    def __init__(self, code, where):
        msg = _lib.lib().gpurelay_last_error()
        super().__init__(f"{where} failed: {code} ({msg.decode() if msg else ''})")
        self.code = code


def _check(code, where):
    if code != 0:
        raise RelayError(code, where)


def chain_parent(rank: int) -> int | None:
    """Predecessor of `rank` in the relay chain (None for the ingest rank)."""
    return None if rank == 0 else rank - 1


class Relay:
    HANDLE = 64

    def __init__(self, dist, rank, world, device, slot_bytes, nr_slots, tag="0"):
        import torch
        L = self._L = _lib.lib()
        self.rank, self.world, self.nr_slots, self.slot_bytes = rank, world, nr_slots, slot_bytes
        self._name = f"/tslb200_relay_{tag}".encode()
        cfg = _lib.GpuRelayCfg()
        cfg.struct_size = C.sizeof(_lib.GpuRelayCfg)
        cfg.rank, cfg.world, cfg.nr_slots, cfg.device, cfg.slot_bytes = rank, world, nr_slots, device, slot_bytes
        cfg.shm_name = self._name
        self._h = C.c_void_p()
        if rank == 0:                                   # rank 0 creates the counter segment, the others open it afterwards
            _check(L.gpurelay_create(C.byref(self._h), C.byref(cfg)), "gpurelay_create")
        dist.barrier()
        if rank != 0:
            _check(L.gpurelay_create(C.byref(self._h), C.byref(cfg)), "gpurelay_create")
        # every rank publishes the IPC handles of its slots; rank r opens those of rank r - 1
        mine = (C.c_uint8 * (self.HANDLE * nr_slots))()
        _check(L.gpurelay_export(self._h, mine), "gpurelay_export")
        xdev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")    # gloo moves CPU tensors
        t = torch.tensor(list(mine), dtype=torch.uint8, device=xdev)
        allh = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allh, t)
        parent = chain_parent(rank)
        if parent is not None:
            raw = bytes(allh[parent].cpu().tolist())
            buf = (C.c_uint8 * len(raw)).from_buffer_copy(raw)
            _check(L.gpurelay_connect_ipc(self._h, buf), "gpurelay_connect_ipc")
        dist.barrier()
        self._aux = torch.cuda.Stream(device=torch.device("cuda", device))

    def slot_ptr(self, i) -> int:
        p = C.c_void_p()
        _check(self._L.gpurelay_slot(self._h, i, C.byref(p)), "gpurelay_slot")
        return p.value

    def slot_tensor(self, torch, i):
        """The slot as an int16 torch tensor (no copy; the relay owns the memory)."""
        n = self.slot_bytes // 2

        class _Arr:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i2", "data": (self.slot_ptr(i), False), "version": 3}
        return torch.as_tensor(_Arr(), device=torch.device("cuda", torch.cuda.current_device()))

    def acquire(self, seq, producer_stream):
        """Ingest rank: make producer_stream wait until slot seq % nr_slots may be overwritten."""
        _check(self._L.gpurelay_acquire(self._h, seq, producer_stream), "gpurelay_acquire")

    def advance(self, seq, nbytes, bank=None, producer_stream=0) -> int:
        ready = C.c_void_p()
        _check(self._L.gpurelay_advance(self._h, seq, nbytes, producer_stream, C.byref(ready)), "gpurelay_advance")
        return ready.value

    def consumed(self, seq, bank):
        """The bank has been handed batch seq: mark the point after which this GPU no longer reads the slot."""
        bank.stream_wait(self._aux.cuda_stream)
        _check(self._L.gpurelay_release(self._h, seq, self._aux.cuda_stream), "gpurelay_release")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.gpurelay_destroy(C.byref(self._h))
