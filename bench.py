#!/usr/bin/env python
"""bench.py -- throughput of the hot path (wide-band IQ -> per-channel FIR/mix/decimate -> derotate -> FM PCM).

A "step" is SUBMITS passes of the channel bank, each over one batch of 2^25 synthetic int16 IQ samples.
Default workload = the shape BASELINE.json's north_star quotes its target on: 256 channels x 127 taps, decimate by
100, 2.4 MS/s, per B200 (--config c2|c3|c4|c5 select BASELINE.json's other shapes; c4 is the strong-scaling sweep).
  value : channel-samples/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the C ABI with HOST buffers: pinned H2D of every batch + D2H of all PCM inside the
          timed region.  N > 1: the batch crosses PCIe ONCE (pinned host -> ingest GPU), rides the relay chain, and every
          GPU returns its channels' PCM over its own PCIe link (--e2e-fanout host: every GPU instead copies the batch
          itself from one shared pinned host segment; measured NUMA-bound on the 8-GPU boxes).
  N > 1 : channels shard across ranks; the only data-path exchange is the fan-out of the IQ batch from the ingest GPU
          (rank 0), inside the timed region: a pipelined relay chain over NVLink driven by the copy engines
          (include/tslb200_gpurelay.h), or torch.distributed.broadcast (NCCL) with --fanout nccl.
  --impl reference : the reference's own CPU code (oracle/_ref, one pthread per channel) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402

METRIC = "channel_samples_per_s"
UNIT = "channel-samples/s"
DTYPE = "int32 (int16 x int16 -> int32 MAC) + f32/f64 atan2"

# BASELINE.json configs -> shapes.  C = channels per GPU (weak) or in total (strong).
CONFIGS = {
    "headline": dict(C=256, fs=2_400_000, T=127, D=100, cutoff=9000.0, scaling="weak",
                     what="north_star target shape: 256 channels x 127 taps, decimate by 100, 2.4 MS/s cs16 IQ, per B200"),
    "c2": dict(C=64, fs=2_400_000, T=127, D=100, cutoff=9000.0, scaling="weak",
               what="BASELINE configs[1]: 64-channel channeliser + FM discriminator, 2.4 MS/s, 127 taps, per B200"),
    "c3": dict(C=256, fs=1_200_000, T=127, D=25, cutoff=9000.0, scaling="weak",
               what="BASELINE configs[2] front end: 256 channels, 1.2 MS/s (etc/pocsag_rtlsdr.json shape), 127 taps, decimate by 25"),
    "c4": dict(C=1024, fs=10_000_000, T=255, D=200, cutoff=12000.0, scaling="strong",
               what="BASELINE configs[3]: 1024-channel FM-only sweep, 10 MS/s, 255 taps, decimate by 200, channels split over the GPUs"),
    "c5": dict(C=256, fs=3_000_000, T=512, D=120, cutoff=9000.0, scaling="weak",
               what="BASELINE configs[4] front end: 256 channels, 3 MS/s, 512 taps (etc/flex_25khz_lpf_3mhz.json shape), decimate by 120"),
}


def shape(args, n_gpus):
    cfg = dict(CONFIGS[args.config])
    if cfg["scaling"] == "strong":
        assert cfg["C"] % n_gpus == 0
        cfg["c_gpu"], cfg["c_total"] = cfg["C"] // n_gpus, cfg["C"]
    else:
        cfg["c_gpu"], cfg["c_total"] = cfg["C"], cfg["C"] * n_gpus
    return cfg


def config_dict(args, cfg, n_gpus):
    """Identical in both arms (same keys, same values): what the workload is, not how an arm runs it."""
    n = 1 << args.batch_log2
    return {"workload": f"{args.config}: {cfg['what']}", "channels": cfg["c_total"], "channels_per_gpu": cfg["c_gpu"],
            "taps": cfg["T"], "decimation": cfg["D"], "fs": cfg["fs"], "batch_complex_samples": n,
            "submits_per_step": args.submits,
            "l2_policy": f"inputs larger than L2: alternating batches of {4 * n >> 20} MiB",
            "parallelism": (f"channels sharded {cfg['c_gpu']}/GPU over {n_gpus} GPUs, IQ batch fanned out from the ingest GPU"
                            if n_gpus > 1 else "1 GPU")}


def channel_plan(cfg):
    from tsl_sdr_b200 import synth
    offs = synth.channel_offsets(cfg["c_total"], cfg["fs"])
    lpf = synth.lowpass_taps(cfg["T"], cfg["cutoff"], cfg["fs"])
    return lpf, offs


def measured_traffic(config, engine, n):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed ncu capture of
    this very workload (profiles/r02_traffic.json); None when the capture does not match what is being run."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if t.get("config") == config and t.get("engine") == engine and int(t.get("batch_complex_samples", 0)) == n:
            return float(t["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def i8_peak():
    """Measured tcgen05 kind::i8 rate of this part (tools/tc_microbench.cu, profiles/r02_i8_peak.json), int8 MAC/s."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_i8_peak.json")))
        return float(t["int8_mac_per_s"]), t["source"]
    except Exception:
        return 148 * 8192 * 1.965e9, "nominal 8192 int8 MAC/clk/SM x 148 SMs x 1965 MHz (no measurement found)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.index), "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def wait_first_sample(self, timeout=5.0):
        """nvidia-smi needs about a second to start: do not begin the timed region before it is sampling"""
        t0 = time.time()
        while self.p is not None and time.time() - t0 < timeout:
            self.f.flush()
            if os.path.getsize(self.f.name) > 0:
                return
            time.sleep(0.05)

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        sm, mx, reasons = [], [], set()
        for line in open(self.f.name):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def gen_batches(torch, dev, n, nr, fs, seed):
    """Synthetic wide-band IQ on the device: a few FM carriers across the band + Gaussian noise, int16 pairs."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = []
    t = torch.arange(n, device=dev, dtype=torch.float64)
    for b in range(nr):
        acc_re = torch.randn(n, device=dev, generator=g, dtype=torch.float32) * 50.0
        acc_im = torch.randn(n, device=dev, generator=g, dtype=torch.float32) * 50.0
        for i, frac in enumerate((-0.421875, -0.140625, 0.0140625, 0.2109375, 0.4078125)):
            off = frac * fs
            ph = (2.0 * np.pi * off / fs) * t + 0.7 * i + 3.0 * torch.sin(2 * np.pi * (400.0 + 150 * i + 31 * b) / fs * t)
            ph = torch.remainder(ph, 2 * np.pi).to(torch.float32)
            acc_re += 1900.0 * torch.cos(ph)
            acc_im += 1900.0 * torch.sin(ph)
        iq = torch.stack((acc_re, acc_im), dim=1).round().clamp(-32768, 32767).to(torch.int16).contiguous()
        out.append(iq.view(-1))
    del t
    return out


def load_ref():
    import pyoracle
    try:
        return pyoracle.Ref("native"), "-O3 -march=native (the reference's Release flags)"
    except OSError:
        return pyoracle.Ref("fma"), "-O3 -march=x86-64-v3"


def run_reference(args):
    """The reference's CPU implementation of the same path (oracle/_ref) on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import pyoracle
    import tslb200_loader
    tslb200_loader.load_package()
    from tsl_sdr_b200 import synth
    n_gpus = args.gpus
    cfg = shape(args, n_gpus)
    lpf, offs = channel_plan(cfg)
    nr_ch = len(offs)
    fs, D = cfg["fs"], cfg["D"]
    n = 1 << 21
    iq = synth.synth_noise_tones_iq(n, fs, offs[:: max(1, nr_ch // 16)], seed=1)
    kind = "reference"
    try:
        ref, flavour = load_ref()
    except OSError:
        ref = None
    ncpu = os.cpu_count() or 1
    if ref is not None:
        secs, outs = ref.bench_multifm(lpf, offs, fs, D, iq, reps=1)          # calibrate
        reps = max(1, min(64, int(1.0 / max(secs, 1e-3))))
        vals = []
        for _ in range(args.warmup):
            ref.bench_multifm(lpf, offs, fs, D, iq, reps=reps)
        t_tot = 0.0
        for _ in range(args.steps):
            secs, outs = ref.bench_multifm(lpf, offs, fs, D, iq, reps=reps)
            vals.append(outs / secs)
            t_tot += secs
        value = statistics.median(vals)
        ms = 1e3 * t_tot / args.steps
        sample = (f"bounded sample of the workload: {nr_ch} channel pthreads x {reps} passes over {n} in-memory complex samples "
                  f"per step (4096-sample sample_bufs, no file/FIFO I/O); {flavour}")
        cores = min(nr_ch, ncpu)
    else:
        kind = "port"
        orc = pyoracle.Oracle()
        t0 = time.perf_counter()
        _, p = orc.channel(lpf, offs[0], fs, D, iq)
        secs = time.perf_counter() - t0
        value = len(p) / secs
        ms = secs * 1e3
        sample = f"oracle port, 1 thread, 1 channel over {n} samples"
        cores = 1
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": config_dict(args, cfg, n_gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                             "host_cpus": ncpu, "iq_msps": value * D / nr_ch / 1e6},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def bind_to_gpu_numa_node(torch, index):
    """Run this rank (and therefore first-touch its pinned host buffers) on the NUMA node its GPU hangs off, as a
    multi-socket receiver would place its per-GPU staging: without it the DMA of seven of eight GPUs crosses the socket
    interconnect.  Returns the node or None when the platform does not say."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


class SharedPinned:
    """One copy of the IQ batches in host memory for all ranks of the box (the single ingest buffer of a receiver
    process): a /dev/shm segment every rank maps and page-locks, so each GPU DMAs it over its own PCIe link."""

    def __init__(self, torch, dist, rank, nbytes, tag):
        self.torch = torch
        self.path = f"/dev/shm/tslb200_bench_{tag}.bin"
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(nbytes)
        if dist is not None:
            dist.barrier()
        self.arr = np.memmap(self.path, dtype=np.int16, mode="r+", shape=(nbytes // 2,))
        self.ptr = self.arr.ctypes.data
        rc = torch.cuda.cudart().cudaHostRegister(self.ptr, nbytes, 0)
        if int(rc) != 0:
            raise SystemExit(f"cudaHostRegister failed: {rc}")
        self.rank, self.dist = rank, dist

    def close(self):
        self.torch.cuda.cudart().cudaHostUnregister(self.ptr)
        del self.arr
        if self.dist is not None:
            self.dist.barrier()
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="headline", choices=sorted(CONFIGS))
    ap.add_argument("--batch-log2", type=int, default=25, help="complex samples per submit = 2^this (134 MB > L2 at 25)")
    ap.add_argument("--submits", type=int, default=16, help="submits per step (keeps the timed region above 50 ms: 10 steps x 16 x 0.3 ms)")
    ap.add_argument("--engine", type=int, default=0)
    ap.add_argument("--fanout", default="relay", choices=["relay", "nccl"], help="N > 1: how the IQ batch reaches the other GPUs")
    ap.add_argument("--e2e-fanout", default="relay", choices=["relay", "host"], help="N > 1: how the host batch reaches the GPUs in the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import tslb200_loader
    tslb200_loader.load_package()
    from tsl_sdr_b200 import shard
    from tsl_sdr_b200.gpuchan import GpuChan, F_ATAN_FMA

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(torch, local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    cfg = shape(args, world)
    fs, T, D, c_gpu = cfg["fs"], cfg["T"], cfg["D"], cfg["c_gpu"]
    n = 1 << args.batch_log2
    S = args.submits
    lpf, offs_all = channel_plan(cfg)
    ch_lo, ch_hi = shard.shard_range(len(offs_all), world, rank)      # contiguous channel range of this rank
    offs = offs_all[ch_lo:ch_hi]
    assert ch_hi - ch_lo == c_gpu
    bank = GpuChan(lpf, offs, fs, D, n, device=local_rank, flags=F_ATAN_FMA, engine=args.engine)

    NB = 2
    stream = torch.cuda.Stream(device=dev)      # a real (non-NULL) stream: events and kernels share it
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0
    relay = None
    if world > 1 and args.fanout == "relay":
        from tsl_sdr_b200.relay import Relay
        relay = Relay(dist, rank, world, local_rank, 4 * n, NB + 1, tag=os.environ.get("MASTER_PORT", "0"))
        batches = [relay.slot_tensor(torch, i) for i in range(relay.nr_slots)]
        if rank == 0:
            for i, b in enumerate(gen_batches(torch, dev, n, relay.nr_slots, fs, seed=20260925)):
                batches[i].copy_(b)
    elif rank == 0:
        batches = gen_batches(torch, dev, n, NB, fs, seed=20260925)
    else:
        batches = [torch.empty(2 * n, dtype=torch.int16, device=dev) for _ in range(NB)]
    NBUF = len(batches)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # NCCL fan-out: one tiny stream per input buffer, ordered after the bank's kernels that read that buffer; the producer
    # of the buffer's next content waits for it -- so the broadcast of batch i+1 overlaps the kernel of batch i, but batch
    # i+2 never overwrites what the kernel of batch i still reads
    readers = [torch.cuda.Stream(device=dev) for _ in range(NBUF)]
    seq = [0]

    def submit_resident():
        i = seq[0]
        seq[0] += 1
        slot = i % NBUF
        buf = batches[slot]
        if relay is not None:
            # rank 0: the batch is resident; others: pull it from the previous GPU of the chain (copy engine, NVLink).
            # Returns the stream after whose current position the slot holds batch i on this GPU.
            rs = relay.advance(i, 4 * n, bank)
            bank.submit_device(buf.data_ptr(), n, rs)
            relay.consumed(i, bank)
        elif dist is not None:
            stream.wait_stream(readers[slot])
            shard.broadcast_iq(dist, buf, src=0)
            bank.submit_device(buf.data_ptr(), n, sptr)
            bank.stream_wait(readers[slot].cuda_stream)
        else:
            bank.submit_device(buf.data_ptr(), n, sptr)
        k = bank.pending()
        bank.discard()                                  # results stay on the device in this leg
        return k

    # host buffers of the end-to-end leg, allocated up front so that the two timed legs run back to back (the clock
    # sampler spans both; an idle gap between them would show up as low clocks)
    e2e_relay = relay is not None and args.e2e_fanout == "relay"
    host_in = pin_in = None
    if e2e_relay:
        if rank == 0:                                   # the one host copy of the stream lives with the ingest rank
            pin_in = [batches[b].cpu().pin_memory() for b in range(NB)]
        h2d_stream = torch.cuda.Stream(device=dev)
    else:
        host_in = SharedPinned(torch, dist, rank, NB * 4 * n, os.environ.get("MASTER_PORT", str(os.getpid())))
        if rank == 0:
            for b in range(NB):
                host_in.arr[b * 2 * n:(b + 1) * 2 * n] = batches[b].cpu().numpy()
    pin_out = torch.empty((c_gpu, n // D + 16), dtype=torch.int16).pin_memory()
    barrier()

    # ---------------- device-resident throughput ("value") ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup * S):
        submit_resident()
    if rank == 0:
        sampler.wait_first_sample()
        for _ in range(S):                              # the GPU idled while nvidia-smi started: one more warm-up step
            submit_resident()
    else:
        for _ in range(S):
            submit_resident()
    barrier()
    bank.timing_read()
    bank.timing_enable(True)
    launches0 = bank.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    k_out = 0
    for _ in range(args.steps * S):
        k_out += submit_resident()
    bank.stream_wait(sptr)                              # the bank works on its own streams: order ours after it
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    kern_ms, kern_n = bank.timing_read()
    bank.timing_enable(False)
    launches = bank.kernel_launches - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    tot_out = torch.tensor([float(k_out * c_gpu)], dtype=torch.float64, device=dev)
    kmax = torch.tensor([kern_ms / max(1, kern_n)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_out, op=dist.ReduceOp.SUM)
        dist.all_reduce(kmax, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = float(tot_out.item()) / (ms_total * 1e-3)

    # ---------------- end to end through the C ABI with host buffers ----------------
    k_per_submit = k_out // (args.steps * S)
    eseq = [0]

    def e2e_submit():
        i = eseq[0]
        eseq[0] += 1
        if not e2e_relay:
            bank.submit_ptr(host_in.ptr + (i % NB) * 4 * n, n)          # pinned H2D inside the C ABI call, this GPU's own link
            return
        q = seq[0]                                                      # the relay's batch numbering continues
        seq[0] += 1
        slot = batches[q % NBUF]
        producer = 0
        if rank == 0:                                                   # host -> ingest GPU: the only PCIe crossing of the input
            relay.acquire(q, h2d_stream.cuda_stream)
            with torch.cuda.stream(h2d_stream):
                slot.copy_(pin_in[i % NB], non_blocking=True)
            producer = h2d_stream.cuda_stream
        rs = relay.advance(q, 4 * n, producer_stream=producer)
        bank.submit_device(slot.data_ptr(), n, rs)
        relay.consumed(q, bank)

    def e2e_collect():
        return bank.collect_into(pin_out.data_ptr(), pin_out.shape[1])  # D2H of every channel's PCM, blocking

    # two batches in flight: the H2D of batch i+1 overlaps the kernels of batch i and the D2H of batch i-1
    e2e_submit()
    for _ in range(2):
        e2e_submit()
        e2e_collect()
    barrier()
    t0 = time.perf_counter()
    e_out = 0
    for _ in range(args.steps * S):
        e2e_submit()
        e_out += e2e_collect()
    barrier()
    e_secs = time.perf_counter() - t0
    te = torch.tensor([e_secs], dtype=torch.float64, device=dev)
    eo = torch.tensor([float(e_out * c_gpu)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(eo, op=dist.ReduceOp.SUM)
    e2e_value = float(eo.item()) / float(te.item())
    clocks = sampler.stop() if rank == 0 else None          # sampled every 20 ms across both timed legs
    e2e_collect()                                                       # drain the batch still in flight
    checksum = int(pin_out[:, :k_per_submit].to(torch.int64).sum().item())

    if rank == 0:
        peak, peak_src = peaks()
        engine = {1: "imad", 2: "tc"}.get(bank.engine, "?")
        alg_bytes = 4.0 * n + 2.0 * c_gpu * k_per_submit                # 4N + 2*C*K per launch
        kern_avg_ms = kern_ms / max(1, kern_n)
        achieved = alg_bytes / (kern_avg_ms * 1e-3) / 1e9
        macs16 = 4.0 * T * c_gpu * k_per_submit
        mma_per_tile, mma_n, tile_out, groups = bank.tc_model()
        i8_issued = float(mma_per_tile) * 128 * mma_n * 32 * groups * (-(-k_per_submit // tile_out) if tile_out else 0)
        i8pk, i8src = i8_peak()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic",
            "config": config_dict(args, cfg, world),
            "engine": engine,
            "fanout": (args.fanout if world > 1 else None),
            "host_numa_node_rank0": numa_node,
            "iq_msps": value * D / cfg["c_total"] / 1e6,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * n * S * (1 if (world == 1 or e2e_relay) else world),
                    "d2h_bytes_per_step": 2 * c_gpu * k_per_submit * S * world, "pcm_checksum": checksum,
                    "note": ("pinned host batch -> ingest GPU (one PCIe crossing) -> NVLink relay chain -> every GPU's PCM back over its own PCIe link"
                             if e2e_relay else "every GPU copies the batch over its own PCIe link from one shared pinned host segment") if world > 1 else
                            "pinned host batch -> gpuchan_submit -> gpuchan_collect into pinned host PCM"},
            "roofline": {"bound": "hbm", "kernel": "tc_fir_fm_kernel (fused mix+FIR+decimate+derotate+FM)" if engine == "tc" else "fir_fm_imad_kernel",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src,
                         "traffic": measured_traffic(args.config, engine, n) if world == 1 else None,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms_per_launch": kern_avg_ms, "kernel_ms_per_launch_max_over_ranks": float(kmax.item()),
                         "kernel_share_of_step": kern_ms / ms_total,
                         "int16_mac_per_s": macs16 / (kern_avg_ms * 1e-3),
                         "int8_mac_issued_per_s": i8_issued / (kern_avg_ms * 1e-3),
                         "mac_frac": i8_issued / (kern_avg_ms * 1e-3) / i8pk, "mac_peak_int8_per_s": i8pk, "mac_peak_source": i8src,
                         "note": "the path is bound by the exact derotate/atan2/PCM epilogue -- FMA-pipe cycles, issue slots and the register "
                                 "file (DESIGN.md 5.2) -- not by HBM or by the tensor pipe: both fractions are reported as required"},
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                ref, flavour = load_ref()
                iq = batches[0][: 2 * (1 << 21)].cpu().numpy()
                secs, outs = ref.bench_multifm(lpf, offs_all, fs, D, iq, reps=1)
                reps = max(1, min(64, int(10.0 / max(secs, 1e-3))))
                secs, outs = ref.bench_multifm(lpf, offs_all, fs, D, iq, reps=reps)
                line["cpu_baseline"] = {"value": outs / secs, "unit": UNIT, "cores": min(len(offs_all), os.cpu_count() or 1),
                                        "kind": "reference", "host_cpus": os.cpu_count(),
                                        "sample": f"reference objects ({flavour}), {len(offs_all)} channel pthreads x {reps} "
                                                  f"passes over the first {1 << 21} samples of batch 0 ({secs:.2f} s wall)"}
            except Exception as exc:       # the oracle always exists; _ref may not
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(exc)}
        print(json.dumps(line))
    bank.close()
    if host_in is not None:
        host_in.close()
    if relay is not None:
        relay.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
