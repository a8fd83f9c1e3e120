"""Re-run the CLI end-to-end scenario of tests/test_gpu_facade.py with the binaries of a
given directory (e.g. an ASan build) and print everything they say."""
import os, subprocess, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200"))
import synth
bins = Path(sys.argv[1]); K = int(sys.argv[2]) if len(sys.argv) > 2 else 15
tmp = Path(tempfile.mkdtemp(prefix="kmsc_cli_"))
seqs = synth.phylogeny_sequences(5, 6000, p=0.02, seed=K)
env = dict(os.environ, ASAN_OPTIONS="protect_shadow_gap=0:detect_leaks=0:abort_on_error=0")
files = []
for i, s in enumerate(seqs):
    txt = synth.to_ascii(s).decode()
    lines = []
    for rep in range(3):
        for a in range(0, len(txt) - 200, 150):
            lines += [f">r{rep}_{a}", txt[a:a + 230]]
    fa = tmp / f"in{i}.fa"; fa.write_text("\n".join(lines) + "\n")
    out = tmp / f"set{i}.txt"
    r = subprocess.run([str(bins / "kmerset-build"), f"--k={K}", "--cutoff=2", f"--out={out}", str(fa)], capture_output=True, text=True, env=env)
    print("build", i, r.returncode, r.stderr[-600:])
    files.append(str(out))
r = subprocess.run([str(bins / "kmerset-multiple-compress"), f"--k={K}", f"--out={tmp / 'dump'}", "--seed=3"] + files, capture_output=True, text=True, env=env)
print("compress rc", r.returncode); print(r.stderr[-6000:])
