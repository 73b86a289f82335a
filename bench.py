#!/usr/bin/env python
"""bench.py -- throughput of the hot path (wide-band IQ -> per-channel FIR/mix/decimate -> FM PCM).

A "step" is one pass of the channel bank over one batch of synthetic int16 IQ.
Workload (BASELINE.json configs[1]): 64 channels/GPU, 2.4 MS/s shape, 127-tap LPF, decimate-by-100.
  value : channel-samples/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the C ABI with HOST buffers (pinned H2D of the batch + D2H of all PCM per step)
  N > 1 : channels shard across ranks (64 per GPU, weak scaling); rank 0 owns the IQ batch and broadcasts
          it over NCCL inside the timed region; no other collective exists on this path.
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

FS = 2_400_000
T = 127
D = 100
C_PER_GPU = 64
CUTOFF_HZ = 9000.0
METRIC = "channel_samples_per_s"
UNIT = "channel-samples/s"


def workload_name(n_gpus):
    return (f"{C_PER_GPU * n_gpus}-channel FIR(127 taps, complex band-pass)+decimate-by-100+derotate+FM discriminator, "
            f"2.4 MS/s cs16 IQ shape, {C_PER_GPU} channels per B200")


def channel_plan(n_gpus):
    from tsl_sdr_b200 import synth
    offs = synth.channel_offsets(C_PER_GPU * n_gpus, FS)
    lpf = synth.lowpass_taps(T, CUTOFF_HZ, FS)
    return lpf, offs


def measured_traffic(engine, n):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed ncu capture of
    this very workload (profiles/r01_traffic.json); None when the capture does not match what is being run."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if t.get("engine") == engine and int(t.get("batch_complex_samples", 0)) == n:
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


def gen_batches(torch, dev, n, nr, seed):
    """Synthetic wide-band IQ on the device: a few FM carriers on the channel grid + Gaussian noise, int16 pairs."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = []
    t = torch.arange(n, device=dev, dtype=torch.float64)
    for b in range(nr):
        acc_re = torch.randn(n, device=dev, generator=g, dtype=torch.float32) * 50.0
        acc_im = torch.randn(n, device=dev, generator=g, dtype=torch.float32) * 50.0
        for i, off in enumerate((-1_012_500, -337_500, 33_750, 506_250, 978_750)):
            ph = (2.0 * np.pi * off / FS) * t + 0.7 * i + 3.0 * torch.sin(2 * np.pi * (400.0 + 150 * i + 31 * b) / FS * t)
            ph = torch.remainder(ph, 2 * np.pi).to(torch.float32)
            acc_re += 1900.0 * torch.cos(ph)
            acc_im += 1900.0 * torch.sin(ph)
        iq = torch.stack((acc_re, acc_im), dim=1).round().clamp(-32768, 32767).to(torch.int16).contiguous()
        out.append(iq.view(-1))
    del t
    return out


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
    lpf, offs = channel_plan(n_gpus)
    nr_ch = len(offs)
    n = 1 << 21
    iq = synth.synth_noise_tones_iq(n, FS, offs[:: max(1, nr_ch // 16)], seed=1)
    kind = "reference"
    try:
        try:
            ref = pyoracle.Ref("native")
            flavour = "-O3 -march=native (the reference's Release flags)"
        except OSError:
            ref = pyoracle.Ref("fma")
            flavour = "-O3 -march=x86-64-v3"
    except OSError:
        ref = None
    ncpu = os.cpu_count() or 1
    if ref is not None:
        secs, outs = ref.bench_multifm(lpf, offs, FS, D, iq, reps=1)          # calibrate
        reps = max(1, min(64, int(1.0 / max(secs, 1e-3))))
        vals = []
        for _ in range(args.warmup):
            ref.bench_multifm(lpf, offs, FS, D, iq, reps=reps)
        t_tot = 0.0
        for _ in range(args.steps):
            secs, outs = ref.bench_multifm(lpf, offs, FS, D, iq, reps=reps)
            vals.append(outs / secs)
            t_tot += secs
        value = statistics.median(vals)
        ms = 1e3 * t_tot / args.steps
        sample = (f"{nr_ch} channel pthreads x {reps} passes over {n} in-memory complex samples per step "
                  f"(4096-sample sample_bufs, no file/FIFO I/O); {flavour}")
        cores = min(nr_ch, ncpu)
    else:
        kind = "port"
        orc = pyoracle.Oracle()
        t0 = time.perf_counter()
        _, p = orc.channel(lpf, offs[0], FS, D, iq)
        secs = time.perf_counter() - t0
        value = len(p) / secs
        ms = secs * 1e3
        sample = f"oracle port, 1 thread, 1 channel over {n} samples"
        cores = 1
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32 (int16 x int16 -> int32 MAC) + f32/f64 atan2", "data": "synthetic",
            "config": {"workload": workload_name(n_gpus), "channels": nr_ch, "taps": T, "decimation": D, "fs": FS},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                             "host_cpus": ncpu, "iq_msps": value * D / nr_ch / 1e6},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-log2", type=int, default=25, help="complex samples per step = 2^this (134 MB > L2 at 25)")
    ap.add_argument("--engine", type=int, default=0)
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
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    n = 1 << args.batch_log2
    lpf, offs_all = channel_plan(world)
    ch_lo, ch_hi = shard.shard_range(len(offs_all), world, rank)      # contiguous channel range of this rank
    offs = offs_all[ch_lo:ch_hi]
    assert ch_hi - ch_lo == C_PER_GPU
    bank = GpuChan(lpf, offs, FS, D, n, device=local_rank, flags=F_ATAN_FMA, engine=args.engine)

    NB = 2
    if rank == 0:
        batches = gen_batches(torch, dev, n, NB, seed=20260925)
    else:
        batches = [torch.empty(2 * n, dtype=torch.int16, device=dev) for _ in range(NB)]
    stream = torch.cuda.Stream(device=dev)      # a real (non-NULL) stream: events and kernels share it
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # one tiny stream per input buffer: it is ordered after the bank's kernels that read that buffer, and the producer
    # of the buffer's next content (NCCL broadcast / H2D copy on `stream`) waits for it -- so the broadcast of batch
    # i+1 overlaps the kernel of batch i, but batch i+2 never overwrites what the kernel of batch i still reads
    readers = [torch.cuda.Stream(device=dev) for _ in range(NB)]

    def step_resident(i):
        buf = batches[i % NB]
        if dist is not None:
            stream.wait_stream(readers[i % NB])
            shard.broadcast_iq(dist, buf, src=0)
        bank.submit_device(buf.data_ptr(), n, sptr)
        if dist is not None:
            bank.stream_wait(readers[i % NB].cuda_stream)
        k = bank.pending()
        bank.discard()                                  # results stay on the device in this leg
        return k

    # host buffers of the end-to-end leg, allocated up front so that the two timed legs run back to back (the clock
    # sampler spans both; an idle gap between them would show up as low clocks)
    pin_in = [torch.empty(2 * n, dtype=torch.int16).pin_memory() for _ in range(NB)]
    if rank == 0:
        for a, b in zip(pin_in, batches):
            a.copy_(b)
    pin_out = torch.empty((C_PER_GPU, n // D + 16), dtype=torch.int16).pin_memory()
    torch.cuda.synchronize()

    # ---------------- device-resident throughput ("value") ----------------
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    bank.timing_read()
    bank.timing_enable(True)
    launches0 = bank.kernel_launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    k_out = 0
    for i in range(args.steps):
        k_out += step_resident(args.warmup + i)
    bank.stream_wait(sptr)                              # the bank works on its own streams: order ours after it
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    kern_ms, kern_n = bank.timing_read()
    bank.timing_enable(False)
    launches = bank.kernel_launches - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    tot_out = torch.tensor([float(k_out * C_PER_GPU)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_out, op=dist.ReduceOp.SUM)
    ms_total = float(t.item())
    value = float(tot_out.item()) / (ms_total * 1e-3)

    # ---------------- end to end through the C ABI with host buffers ----------------
    k_per_step = k_out // args.steps

    def e2e_submit(i):
        if dist is None:
            bank.submit_ptr(pin_in[i % NB].data_ptr(), n)              # pinned H2D inside the C ABI call
        else:
            buf = batches[i % NB]
            stream.wait_stream(readers[i % NB])
            if rank == 0:
                buf.copy_(pin_in[i % NB], non_blocking=True)
            shard.broadcast_iq(dist, buf, src=0)
            bank.submit_device(buf.data_ptr(), n, sptr)
            bank.stream_wait(readers[i % NB].cuda_stream)

    def e2e_collect():
        return bank.collect_into(pin_out.data_ptr(), pin_out.shape[1])  # D2H of every channel's PCM, blocking

    # two batches in flight: the H2D of batch i+1 overlaps the kernels of batch i and the D2H of batch i-1
    e2e_submit(0)
    for i in range(2):
        e2e_submit(i + 1)
        e2e_collect()
    barrier()
    t0 = time.perf_counter()
    e_out = 0
    for i in range(args.steps):
        e2e_submit(i + 3)
        e_out += e2e_collect()
    barrier()
    e_secs = time.perf_counter() - t0
    te = torch.tensor([e_secs], dtype=torch.float64, device=dev)
    eo = torch.tensor([float(e_out * C_PER_GPU)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(eo, op=dist.ReduceOp.SUM)
    e2e_value = float(eo.item()) / float(te.item())
    clocks = sampler.stop() if rank == 0 else None          # sampled every 20 ms across both timed legs
    e2e_collect()                                                       # drain the batch still in flight
    checksum = int(pin_out[:, :k_per_step].to(torch.int64).sum().item())

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = 4.0 * n + 2.0 * C_PER_GPU * k_per_step              # 4N + 2*C*K per launch
        kern_avg_ms = kern_ms / max(1, kern_n)
        achieved = alg_bytes / (kern_avg_ms * 1e-3) / 1e9
        macs = 4.0 * T * C_PER_GPU * k_per_step
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32 (int16 x int16 -> int32 MAC) + f32/f64 atan2", "data": "synthetic",
            "config": {"workload": workload_name(world), "channels": C_PER_GPU * world, "taps": T, "decimation": D,
                       "fs": FS, "batch_complex_samples": n, "engine": {1: "imad", 2: "tc"}.get(bank.engine, "?"),
                       "l2_policy": f"inputs larger than L2: {NB} alternating batches of {4 * n >> 20} MiB",
                       "parallelism": f"channels sharded {C_PER_GPU}/GPU, NCCL broadcast of IQ" if world > 1 else "1 GPU"},
            "iq_msps": value * D / C_PER_GPU / world / 1e6 * 1.0,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * n,
                    "d2h_bytes_per_step": 2 * C_PER_GPU * k_per_step * world, "pcm_checksum": checksum},
            "roofline": {"bound": "hbm", "kernel": "fir_fm kernel (fused mix+FIR+decimate+derotate+FM)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src,
                         "traffic": measured_traffic({1: "imad", 2: "tc"}.get(bank.engine, "?"), n) if world == 1 else None,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms_per_launch": kern_avg_ms, "kernel_share_of_step": kern_ms / ms_total,
                         "int16_mac_per_s": macs / (kern_avg_ms * 1e-3),
                         "note": "path is integer-compute bound at 64 ch x 127 taps (SURVEY.md 8d): HBM fraction is "
                                 "reported as required; see DESIGN.md for the MAC/s roofline"},
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                import pyoracle
                try:
                    ref = pyoracle.Ref("native"); flavour = "-O3 -march=native"
                except OSError:
                    ref = pyoracle.Ref("fma"); flavour = "-O3 -march=x86-64-v3"
                from tsl_sdr_b200 import synth
                iq = batches[0][: 2 * (1 << 21)].cpu().numpy()
                secs, outs = ref.bench_multifm(lpf, offs_all, FS, D, iq, reps=1)
                reps = max(1, min(64, int(1.5 / max(secs, 1e-3))))
                secs, outs = ref.bench_multifm(lpf, offs_all, FS, D, iq, reps=reps)
                line["cpu_baseline"] = {"value": outs / secs, "unit": UNIT, "cores": min(len(offs_all), os.cpu_count() or 1),
                                        "kind": "reference", "host_cpus": os.cpu_count(),
                                        "sample": f"reference objects ({flavour}), {len(offs_all)} channel pthreads x {reps} "
                                                  f"passes over the first {1 << 21} samples of batch 0 ({secs:.2f} s wall)"}
            except Exception as exc:       # the oracle always exists; _ref may not
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(exc)}
        print(json.dumps(line))
    bank.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
