"""Host logic without a device: the streaming FASTA readers of the C++ facade (host/kmsc/io.h; reference
lib/core/kmer_counter.h:163-166 whole-record contract, lib/core/io.h popen contract) through tests/cpp/io_test.cc."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "kmer-sets-compression_b200" / "host"


def test_record_chunk_readers(tmp_path):
    r = subprocess.run(["make", "-s", "-C", str(HOST), "bin/io_test"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(HOST / "bin" / "io_test")], capture_output=True, text=True, timeout=600, env={"TMPDIR": str(tmp_path), "PATH": "/usr/bin:/bin"})
    assert r.returncode == 0 and "io_test: ok" in r.stdout, r.stdout + r.stderr
