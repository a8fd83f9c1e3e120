"""Times kmerset-build (FASTA -> counts -> cutoff -> KmerSet -> SPSS -> file) of the C++ facade
on a random genome of G bases given as one long read; prints the tool's own log."""
import subprocess, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200"))
import synth
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 23
tmp = Path(tempfile.mkdtemp(prefix="kmsc_spss_"))
fa = tmp / "g.fa"
fa.write_bytes(b">g\n" + synth.to_ascii(synth.random_genome(G)) + b"\n")
t = time.time()
r = subprocess.run([str(ROOT / "kmer-sets-compression_b200/host/bin/kmerset-build"), f"--k={K}", "--check", f"--out={tmp/'out.txt'}", str(fa)],
                   capture_output=True, text=True)
print(f"kmerset-build K={K} on {G} bases: {time.time()-t:.2f} s, rc={r.returncode}")
print(r.stderr[-800:])
lines = (tmp / "out.txt").read_text().split("\n")
print("SPSS strings:", len(lines) - 1, "total chars:", sum(len(x) for x in lines))
