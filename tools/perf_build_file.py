"""f3: kmerset-build on a FASTA FILE (reads, k = 31, cutoff 4) through the chunked device counter, with the file
reader overlapped with the device (default) and in turn (KMSC_IO_OVERLAP=0). The file is written to /dev/shm
(or $TMPDIR) from bench.py's C4 generator; BYTES (default 2e9) of 2-line FASTA."""
import os, subprocess, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
nbytes = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000_000
dev = torch.device("cuda", 0)
recs, rec = bench.gen_fasta_torch(nbytes, 150, 20_000_000, dev)
tmpdir = tempfile.mkdtemp(prefix="kmsc_f3_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
path = os.path.join(tmpdir, "reads.fa")
recs.flatten().cpu().numpy().tofile(path)
n_reads = recs.shape[0]
del recs
torch.cuda.empty_cache()
exe = ROOT / "kmer-sets-compression_b200" / "host" / "bin" / "kmerset-build"
print(f"{os.path.getsize(path) / 1e9:.2f} GB, {n_reads} reads x 150 bp at {path}")
for ov in ("2", "2", "0", "2"):   # 2 = page-locked reader (the default), 1 = pageable overlapped, 0 = in turn
    t = time.time()
    r = subprocess.run([str(exe), "--k=31", "--cutoff=4", "--chunk_mb=256", path], capture_output=True, text=True,
                       env=dict(os.environ, KMSC_IO_OVERLAP=ov, KMSC_TIMING="1"))
    dt = time.time() - t
    size = [l for l in r.stderr.split("\n") if "kmer_set.Size" in l or "Hash" in l]
    print(f"KMSC_IO_OVERLAP={ov}: rc={r.returncode} wall {dt:.2f} s = {n_reads * 150 / dt / 1e9:.2f} Gbases/s  {size}")
    print("   ", [l for l in r.stderr.split("\n") if "timing" in l])
os.remove(path); os.rmdir(tmpdir)
