"""C1 through kmerset-multiple-compress with KMSC_TIMING=1: seconds per phase of the greedy driver."""
import os, subprocess, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
tmp = tempfile.mkdtemp(prefix="kmsc_c1t_")
files = bench.c1_files(tmp)
exe = ROOT / "kmer-sets-compression_b200" / "host" / "bin" / "kmerset-multiple-compress"
for iters in (int(a) for a in (sys.argv[1:] or ["3"])):
    t = time.time()
    r = subprocess.run([str(exe), "--k=15", "--seed=77", f"--max_iterations={iters}"] + files, capture_output=True, text=True,
                       env=dict(os.environ, KMSC_TIMING="1"))
    print(f"max_iterations={iters}: rc={r.returncode} wall {time.time()-t:.2f} s")
    print("\n".join(l for l in r.stderr.split("\n") if "timing" in l or "merges =" in l))
