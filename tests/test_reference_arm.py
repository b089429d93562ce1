"""The CPU arm of bench.py (`--impl reference`, `cpu_baseline`): oracle/ref_runner.py must import and run the byte-compiled
reference from oracle/_ref at toy sizes.  (The timed sizes are bench.py's business; this guards the import chain and the
three entry points the driver's reference arm goes through.)"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "oracle", "_ref")),
                                reason="oracle/_ref not staged (run __graft_entry__.build() where /root/reference exists)")


def test_reference_train_step_runs():
    from oracle import ref_runner
    r = ref_runner.time_train(steps=1, warmup=0, B=2, T=40, threads=2)
    assert r["kind"] == "reference" and r["cores"] == 2 and r["value"] > 0
    assert "unmodified reference" in r["sample"]


def test_reference_infer_runs():
    from oracle import ref_runner
    r = ref_runner.time_infer(steps=1, warmup=0, T=16, L=12, threads=2)
    assert r["kind"] == "reference" and r["value"] > 0 and r["rtf"] > 0


def test_reference_mel_runs():
    from oracle import ref_runner
    r = ref_runner.time_mel(n_utt=2, threads=2)
    assert r["kind"] == "reference" and r["value"] > 0


def test_bench_cli_reference_arm_prints_one_json_line():
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "mel",
                          "--steps", "1", "--warmup", "0", "--utterances", "3"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and d["unit"] == d["e2e"]["unit"]
