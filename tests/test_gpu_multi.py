"""GPU: the multi-GPU forms of the path through the C ABI.

* gpuchan_multi_* (one process, N devices): channels sharded over the devices, every batch delivered to all of them,
  PCM back in configuration order -- against the oracle, for both fan-outs (one PCIe copy per device / NVLink relay chain).
  On a one-GPU box the same code runs with the device listed twice (two banks + a two-hop chain on one device), which
  exercises everything but the NVLink transfer itself.
* gpurelay_* across processes (one process per GPU, the torchrun launch of bench.py): tests/relay_worker.py under
  torch.distributed.run with two ranks; every rank checks its own channels against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tsl_sdr_b200 import synth
from tsl_sdr_b200.gpuchan import GpuChanMulti, FANOUT_HOST, FANOUT_RELAY

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def device_lists():
    import torch
    n = torch.cuda.device_count()
    out = [[0, 0], [0, 0, 0]]
    if n >= 2:
        out.append([0, 1])
    if n >= 4:
        out.append([0, 1, 2, 3])
    return out


@pytest.mark.parametrize("fanout", [FANOUT_HOST, FANOUT_RELAY])
def test_multi_device_bank_matches_oracle(oracle, fanout):
    fs, T, D, C = 2_400_000, 127, 100, 7                    # 7 channels: uneven shards
    offs = synth.channel_offsets(C, fs)
    lpf = synth.lowpass_taps(T, 9000.0, fs)
    n = 300_000
    iq = synth.synth_noise_tones_iq(n, fs, offs)
    exp = [oracle.channel(lpf, offs[c], fs, D, iq)[1] for c in range(C)]
    chunk = 37_000                                          # 9 submits: the relay's 3 slots wrap three times
    for devs in device_lists():
        bank = GpuChanMulti(lpf, offs, fs, D, chunk, devs, fanout=fanout)
        assert bank.devices == len(devs)
        covered = []
        for i in range(bank.devices):
            first, cnt = bank.bank_range(i)
            covered += list(range(first, first + cnt))
        assert covered == list(range(C))
        got = []
        pinned = []
        for s in range(0, n, chunk):
            part = np.ascontiguousarray(iq[2 * s: 2 * min(n, s + chunk)])
            pinned.append(part)                             # keep host buffers alive until collected
            bank.submit(part)
            got.append(bank.collect().copy())
        bank.close()
        got = np.concatenate(got, axis=1)
        for c in range(C):
            assert np.array_equal(got[c], exp[c]), (devs, fanout, c)


def test_relay_chain_across_processes(tmp_path):
    import torch
    ndev = torch.cuda.device_count()
    world = 2 if ndev < 4 else 4
    env = dict(os.environ, RELAY_TEST_DEVICES=str(ndev))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "tests", "relay_worker.py"),
                        str(tmp_path)], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    for rank in range(world):
        assert (tmp_path / f"rank{rank}.ok").exists(), r.stderr[-2000:]
