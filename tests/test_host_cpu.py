"""CPU: the host C side (tsl-sdr_b200/host/: JSON config reader, receiver set-up) without a GPU.

The multifm JSON schema is the reference's (multifm/multifm.c:103-156, multifm/receiver.c:133-229):
merged files, required keys with the reference's error tags (NO-SAMPLE-RATE, NO-CENTER-FREQ, NO-DECIMATION,
BAD-DECIMATION-FACTOR, BAD-FILTER-TAPS, INSUFF-FILTER-TAPS, MISSING-CHANNELS, CANT-OPEN-FIFO), the case-sensitive `dBGain` key (the shipped
etc/pocsag_rtlsdr.json spells it `dbGain`, which the reference silently ignores -- so do we), and on a box without an
sm_100 device the binary fails loudly at bank creation instead of computing anything on the CPU."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tsl-sdr_b200", "b200-multifm")


def run(*files):
    p = subprocess.run([BIN, *files], capture_output=True, text=True, timeout=60)
    return p.returncode, p.stdout + p.stderr


@pytest.fixture()
def cfg(tmp_path):
    (tmp_path / "iq.bin").write_bytes(os.urandom(65536))
    base = {"device": {"type": "file", "filename": str(tmp_path / "iq.bin"), "fileFormat": "cs16"},
            "sampleRateHz": 1200000, "centerFreqHz": 152000000, "decimationFactor": 25, "nrSampBufs": 16,
            "channels": [{"outFifo": str(tmp_path / "ch0.pcm"), "chanCenterFreq": 151680000, "dbGain": 4.0},
                         {"outFifo": str(tmp_path / "ch1.pcm"), "chanCenterFreq": 152150000, "dBGain": 3.0}]}
    taps = {"lpfTaps": [1.0 / 32] * 32}

    def write(name, obj):
        f = tmp_path / name
        f.write_text(obj if isinstance(obj, str) else json.dumps(obj))
        return str(f)
    return base, taps, write


@pytest.mark.skipif(not os.path.exists(BIN), reason="host binary not built")
def test_usage_and_malformed(cfg):
    base, taps, write = cfg
    assert run()[0] == 2
    rc, out = run(write("bad.json", '{"device": '))
    assert rc == 1 and "MALFORMED-CONFIG" in out
    rc, out = run(write("arr.json", "[1, 2]"))
    assert rc == 1 and "MALFORMED-CONFIG" in out
    rc, out = run(write("nodev.json", {"sampleRateHz": 1}))
    assert rc == 1 and "MISSING-DEVICE" in out
    rc, out = run(write("rtl.json", {"device": {"type": "rtlsdr"}}))
    assert rc == 1 and "UNSUPPORTED-DEVICE" in out


@pytest.mark.skipif(not os.path.exists(BIN), reason="host binary not built")
def test_required_keys_have_the_reference_error_names(cfg):
    base, taps, write = cfg
    # multifm/receiver.c:139-184
    for key, tag in (("sampleRateHz", "NO-SAMPLE-RATE"), ("centerFreqHz", "NO-CENTER-FREQ"),
                     ("decimationFactor", "NO-DECIMATION"), ("channels", "MISSING-CHANNELS")):
        c = {k: v for k, v in base.items() if k != key}
        rc, out = run(write("c.json", c), write("t.json", taps))
        assert rc == 1 and tag in out, out
    rc, out = run(write("c.json", dict(base, decimationFactor=0)), write("t.json", taps))
    assert rc == 1 and "BAD-DECIMATION-FACTOR" in out, out
    rc, out = run(write("c.json", base))                                # no lpfTaps in any merged file
    assert rc == 1 and "BAD-FILTER-TAPS" in out, out
    rc, out = run(write("c.json", base), write("t.json", {"lpfTaps": [1.0]}))
    assert rc == 1 and "INSUFF-FILTER-TAPS" in out, out
    rc, out = run(write("c.json", dict(base, nrSampBufs=0)), write("t.json", taps))
    assert rc == 1 and "BAD-SAMP-BUFS" in out, out
    rc, out = run(write("c.json", dict(base, sampleRateHz=1200000.5)), write("t.json", taps))   # integers only
    assert rc == 1 and "NO-SAMPLE-RATE" in out, out


@pytest.mark.skipif(not os.path.exists(BIN), reason="host binary not built")
def test_json_reader_rejects_truncated_and_deep_input(cfg):
    base, taps, write = cfg
    rc, out = run(write("bs.json", '{"device": "abc\\'))              # text ends in a lone backslash
    assert rc == 1 and "MALFORMED-CONFIG" in out and "unterminated string" in out
    rc, out = run(write("deep.json", "[" * 100000))
    assert rc == 1 and "MALFORMED-CONFIG" in out and "nesting too deep" in out


@pytest.mark.skipif(not os.path.exists(BIN), reason="host binary not built")
def test_merged_files_gain_key_and_no_cpu_fallback(cfg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the end-to-end replay is covered by tests/test_gpu_host_binary.py")
    base, taps, write = cfg
    rc, out = run(write("c.json", base), write("t.json", taps))        # taps come from the second file (multifm.c:105-111)
    assert "[1]: 151.68000 MHz Gain: 0.000000 dB" in out               # `dbGain` typo ignored like the reference does
    assert "[2]: 152.15000 MHz Gain: 3.000000 dB" in out
    assert rc == 1 and "no CPU fallback" in out                        # bank creation refuses: nothing runs on the CPU
    assert not os.path.exists(base["channels"][0]["outFifo"]) or os.path.getsize(base["channels"][0]["outFifo"]) == 0
