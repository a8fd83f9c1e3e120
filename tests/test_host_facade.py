"""Host logic without a device: the parts of the C++ facade test that make no device call (tests/cpp/facade_test.cc
--host-only): Kmer (reference test/kmer.cc), ParallelDisjointSet under 8 threads against a serial union-find
(test/parallel_disjoint_set.cc), and the streamvbyte-0124 length codec, whose bytes for a fixed vector must equal the
oracle's restatement (oracle/kmsc_oracle.c)."""
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "kmer-sets-compression_b200" / "host"


def test_facade_host_only(oracle):
    r = subprocess.run(["make", "-s", "-C", str(HOST), "bin/facade_test"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(HOST / "bin" / "facade_test"), "--host-only"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-800:] + r.stderr[-400:]
    line = next(l for l in r.stdout.split("\n") if l.startswith("svb0124"))
    got = bytes(int(x, 16) for x in line.split()[1:])
    fixed = np.array([0 if i % 5 == 0 else ((i * 2654435761) & 0xFFFFFFFF) >> (i % 4 * 8) for i in range(37)], np.uint32)
    assert got == oracle.svb_encode(fixed).tobytes()
