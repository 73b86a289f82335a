"""Worker of tests/test_gpu_multi.py::test_relay_chain_across_processes (run under torch.distributed.run, one process per
rank; with fewer GPUs than ranks several ranks share device 0 -- CUDA IPC and the stream-ordered counters work the same).
Rank 0 owns the IQ stream: every batch is copied from host memory into its relay slot and published; the other ranks pull
it down the chain.  Every rank feeds its own channel range from its own slot and checks the PCM against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    import torch.distributed as dist
    import pyoracle
    import tslb200_loader
    tslb200_loader.load_package()
    from tsl_sdr_b200 import shard, synth
    from tsl_sdr_b200.gpuchan import GpuChan
    from tsl_sdr_b200.relay import Relay

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    ndev = int(os.environ.get("RELAY_TEST_DEVICES", "1"))
    device = rank % max(1, ndev)
    torch.cuda.set_device(device)
    dist.init_process_group("gloo")
    fs, T, D, C = 2_400_000, 127, 100, 6
    offs = synth.channel_offsets(C, fs)
    lpf = synth.lowpass_taps(T, 9000.0, fs)
    n, chunk = 400_000, 41_000
    iq = synth.synth_noise_tones_iq(n, fs, offs)                    # same seed on every rank: only rank 0 uses the samples
    lo, hi = shard.shard_range(C, world, rank)
    bank = GpuChan(lpf, offs[lo:hi], fs, D, chunk, device=device)
    relay = Relay(dist, rank, world, device, 4 * chunk, 3, tag=os.environ.get("MASTER_PORT", "0"))
    st = torch.cuda.Stream()
    got = []
    seq = 0
    for s in range(0, n, chunk):
        k = min(chunk, n - s)
        slot = relay.slot_tensor(torch, seq % 3)
        producer = 0
        if rank == 0:
            src = torch.from_numpy(iq[2 * s: 2 * (s + k)].copy()).pin_memory()
            relay.acquire(seq, st.cuda_stream)                      # the slot's previous content has been consumed everywhere
            with torch.cuda.stream(st):
                slot[: 2 * k].copy_(src, non_blocking=True)
            producer = st.cuda_stream
        ready = relay.advance(seq, 4 * k, producer_stream=producer)
        bank.submit_device(slot.data_ptr(), k, ready)
        relay.consumed(seq, bank)
        got.append(bank.collect().copy())
        if rank == 0:
            st.synchronize()
        seq += 1
    got = np.concatenate(got, axis=1)
    orc = pyoracle.Oracle()
    for i, c in enumerate(range(lo, hi)):
        exp = orc.channel(lpf, offs[c], fs, D, iq)[1]
        assert np.array_equal(got[i], exp), f"rank {rank} channel {c}"
    dist.barrier()
    bank.close()
    relay.close()
    open(os.path.join(sys.argv[1], f"rank{rank}.ok"), "w").write("ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
