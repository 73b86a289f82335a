"""CPU, world_size 2 over gloo: the N > 1 plumbing of the path -- channel sharding and the IQ broadcast -- with the
oracle standing in for the CUDA bank (tests may use the oracle; the product never does).  Each rank receives the IQ
batch by broadcast from rank 0, computes its own contiguous channel range, and the union must equal the
single-process result channel for channel."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, C, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import pyoracle
    import tslb200_loader
    tslb200_loader.load_package()
    from tsl_sdr_b200 import shard, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fs, T, D = 2400000, 127, 100
    offs = synth.channel_offsets(C, fs)
    lpf = synth.lowpass_taps(T, 9000.0, fs)
    orc = pyoracle.Oracle()
    lo, hi = shard.shard_range(C, world, rank)
    states = [orc.new_state(offs[c], fs, D) for c in range(lo, hi)]
    taps = [orc.prepare_taps(lpf, offs[c], fs) for c in range(lo, hi)]
    outs = [[] for _ in range(lo, hi)]
    total = 0
    carry = np.zeros(0, np.int16)
    for b in range(3):                                  # three batches: stream state carries across broadcasts
        buf = torch.zeros(2 * n, dtype=torch.int16)
        if rank == 0:
            rng = np.random.default_rng(100 + b)
            buf.copy_(torch.from_numpy(np.clip(np.round(rng.normal(0, 3000, 2 * n)), -32768, 32767).astype(np.int16)))
        shard.broadcast_iq(dist, buf, src=0)
        window = np.concatenate([carry, buf.numpy()])
        k = 0
        for i in range(hi - lo):
            _, p = orc.chan_stream(states[i], taps[i][0], taps[i][1], D, window, want_iq=False)
            outs[i].append(p)
            k = len(p)
        carry = window[2 * k * D:]
        total += k * (hi - lo)
    whole = shard.gather_counts(dist, torch, total)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), lo=lo, hi=hi, whole=whole,
             pcm=np.stack([np.concatenate(o) for o in outs]) if hi > lo else np.zeros((0, 0), np.int16))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("C", [5, 8])
def test_two_ranks_shard_channels_and_broadcast_iq(oracle, tmp_path, C):
    import torch.multiprocessing as mp
    from tsl_sdr_b200 import shard, synth
    world, n = 2, 30000
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, C, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    fs, T, D = 2400000, 127, 100
    offs = synth.channel_offsets(C, fs)
    lpf = synth.lowpass_taps(T, 9000.0, fs)
    iq = np.concatenate([np.clip(np.round(np.random.default_rng(100 + b).normal(0, 3000, 2 * n)), -32768, 32767).astype(np.int16)
                         for b in range(3)])
    got = {}
    covered = []
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        lo, hi = int(z["lo"]), int(z["hi"])
        assert (lo, hi) == shard.shard_range(C, world, r)
        covered += list(range(lo, hi))
        for i, c in enumerate(range(lo, hi)):
            got[c] = z["pcm"][i]
        whole = int(z["whole"])
    assert covered == list(range(C))
    K = (3 * n - T) // D + 1
    assert whole == K * C
    for c in range(C):
        _, exp = oracle.channel(lpf, offs[c], fs, D, iq)
        assert len(exp) == K and np.array_equal(got[c], exp)


def test_shard_range_properties():
    from tsl_sdr_b200 import shard
    for C in (1, 7, 64, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(C, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == C
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
