"""Times the host-side ingest of a text CSR on this machine's cores: OpenMP text parse against
the binary sidecar (MATFAC_CSR_CACHE).  ML-20M-shaped by default (138 k users, 20 M ratings, ~200 MB of text)."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
MF = os.path.join(ROOT, "matfac_b200", "mf")
n_users, n_items, nnz = (int(x) for x in os.environ.get("SHAPE", "138493,26744,20000263").split(","))
rng = np.random.default_rng(0)
deg = np.maximum(1, (rng.pareto(1.2, n_users) + 1) * nnz / n_users / 6).astype(np.int64)
deg = np.minimum(deg * nnz // deg.sum(), n_items // 2)
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "m.csr")
    t0 = time.time()
    with open(path, "w") as f:
        for u in range(n_users):
            items = np.sort(rng.choice(n_items, size=int(deg[u]), replace=False))
            vals = rng.integers(1, 11, size=items.size) / 2
            f.write(" ".join(f"{i} {v:g}" for i, v in zip(items, vals)) + "\n")
    print(f"text file: {os.path.getsize(path) / 1e6:.0f} MB, {int(deg.sum())} ratings (written in {time.time() - t0:.0f} s)", flush=True)
    tiny = os.path.join(d, "tiny.csr")
    with open(tiny, "w") as f:
        f.write("0 1\n" * n_users)
    def run(cache, threads):
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        env.pop("MATFAC_CSR_CACHE", None)
        if cache:
            env["MATFAC_CSR_CACHE"] = cache
        out = os.path.join(d, "dump"); os.makedirs(out, exist_ok=True)
        t = time.time()
        p = subprocess.run([MF, "--trainmat", path, "--valmat", tiny, "--testmat", tiny, "--prefix", os.path.join(out, "x"), "--facdim", "2",
                            "--dry_run", "1"], env=env, cwd=out, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        assert p.returncode == 0, p.stdout.decode()[-2000:]
        return time.time() - t
    nc = os.cpu_count()
    print(f"text parse, 1 thread          : {run(None, 1):6.2f} s")
    print(f"text parse, {nc} threads         : {run(None, nc):6.2f} s")
    print(f"text parse + sidecar write    : {run('1', nc):6.2f} s  (sidecar {os.path.getsize(path + '.bin') / 1e6:.0f} MB)")
    print(f"sidecar read                  : {run('r', nc):6.2f} s   (whole `mf --dry_run`: CSC index, invalid sets, factor init included)")
