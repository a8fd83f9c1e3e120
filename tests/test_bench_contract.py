"""bench.py contract (CPU part): the reference arm runs without a GPU, prints exactly ONE line on
stdout and that line is the JSON object the driver parses."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--sets", "6", "--kmers", "100000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0
    for key in ("metric", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # the arm says what it measured, and measures both of the reference's modes on all the sets
    assert "reference_sample" in d["config"] and d["config"]["n_sets"] == 6
    assert d["phases"]["pair_rate"] > 0 and d["phases"]["t_decode_wave_s"] > 0
    if d["cpu_baseline"]["kind"] == "reference":
        assert d["sampled"]["n_sets"] == 6 and d["sampled"]["value"] > 0


def test_our_arm_fails_loudly_without_a_gpu():
    """no CPU fallback: without a CUDA device the product arm must not print a bench line"""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "0", "--kmers", "100000"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
