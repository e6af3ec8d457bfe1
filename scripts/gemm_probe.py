"""Probe of the tensor-core pre-expansion (hh_gemm.cu) on a B200: accuracy against an fp64 product and speed, for the
CTA-group / chunk variants.  Every configuration runs in its own process under a timeout so that a wrong barrier
protocol cannot hang the box.

    python scripts/gemm_probe.py                 # the sweep
    python scripts/gemm_probe.py --one N DENS MAXC   # one configuration in this process (env selects the variant)
"""
import json
import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def random_links(n, density, maxc, seed=1, weights=False):
    import numpy as np
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    m = int(n * n * density / 2)
    i = rng.integers(0, n, m)
    j = rng.integers(0, n, m)
    ok = i != j
    i, j = i[ok], j[ok]
    if weights:
        v = rng.random(len(i)).astype(np.float32) * maxc + 0.01
    else:
        # mostly small counts, a few large ones (like neighbouring contigs)
        v = np.minimum(rng.geometric(0.4, len(i)), maxc).astype(np.float32)
        big = rng.random(len(i)) < 0.01
        v[big] = rng.integers(1, maxc + 1, int(big.sum())).astype(np.float32)
    a = sp.coo_matrix((v, (i, j)), shape=(n, n)).tocsr()
    a.sum_duplicates()
    a = sp.triu(a, 1)
    a = a + a.T
    if not weights:
        a.data = np.minimum(a.data, maxc)
    a = sp.csc_matrix(a + sp.identity(n, dtype=np.float32, format="csc"), dtype=np.float32)
    a.sort_indices()
    return a


def one(n, density, maxc, weights, check):
    import numpy as np
    from haphic_b200._lib import Context
    from haphic_b200.links import LinkMatrix
    from haphic_b200.mcl import Mcl
    link = random_links(n, density, maxc, weights=weights)
    out = {"n": n, "density": density, "maxc": maxc, "weights": weights, "cg": os.environ.get("HH_GEMM_CG", "2"),
           "chunk": os.environ.get("HH_GEMM_CHUNK", "dflt")}
    with Context(0) as ctx:
        mat = LinkMatrix.from_csc(ctx, link)
        t0 = time.time()
        mc = Mcl(mat, preexp="dense")
        out["wall_s"] = round(time.time() - t0, 3)
        out.update({k: (round(v, 3) if isinstance(v, float) else v) for k, v in mc.preexp.items()})
        if mc.preexp["gemm_ms"] > 0:
            out["tflops"] = round(mc.preexp["flops"] / mc.preexp["gemm_ms"] / 1e9, 1)
        if check:
            m1 = mc.m1().astype(np.float64)
            d = link.toarray().astype(np.float64)
            s = d.sum(axis=0)
            m0 = (d / s).astype(np.float32).astype(np.float64)       # the fp32 matrix the reference multiplies
            exact = m0 @ m0
            nz = exact != 0
            rel = (m1[nz] - exact[nz]) / exact[nz]
            out["max_rel"] = float(np.abs(rel).max())
            out["mean_rel"] = float(rel.mean())
            out["rms_rel"] = float(np.sqrt((rel ** 2).mean()))
            out["pattern_equal"] = bool(np.array_equal(m1 != 0, nz))
            out["sym_tiles_ok"] = bool(np.isfinite(m1).all())
            # the fp32 ascending-k product the sparse engine (and SciPy) computes, as the yardstick
            mcs = Mcl(mat, preexp="sparse")
            ms = mcs.m1().astype(np.float64)
            rs = (ms[nz] - exact[nz]) / exact[nz]
            out["sparse_max_rel"] = float(np.abs(rs).max())
            out["sparse_mean_rel"] = float(rs.mean())
            out["sparse_ms"] = round(mcs.preexp["total_ms"], 3)
            mcs.close()
        mc.close()
        mat.close()
    print("PROBE " + json.dumps(out), flush=True)


def sweep():
    runs = []
    for cg in ("1", "2"):
        for chunk in ("0", "1", "2", "4"):
            runs.append((dict(HH_GEMM_CG=cg, HH_GEMM_CHUNK=chunk), ["1000", "0.3", "200", "0", "1"]))
    for cg in ("1", "2"):
        runs.append((dict(HH_GEMM_CG=cg), ["777", "0.5", "5000", "0", "1"]))        # two A planes, ragged edge
        runs.append((dict(HH_GEMM_CG=cg), ["300", "0.5", "3", "1", "1"]))           # weights: three A planes, n < 2 tiles
        runs.append((dict(HH_GEMM_CG=cg), ["3000", "0.2", "200", "0", "1"]))
    for cg in ("1", "2"):
        for chunk in ("0", "1", "2", "4"):
            runs.append((dict(HH_GEMM_CG=cg, HH_GEMM_CHUNK=chunk), ["10240", "0.186", "200", "0", "0"]))
    runs.append((dict(HH_GEMM_CG="2", HH_GEMM_CHUNK="2"), ["10240", "0.186", "5000", "0", "0"]))
    for env, args in runs:
        e = dict(os.environ)
        e.update(env)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"] + args, env=e, timeout=240,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            tail = [l for l in r.stdout.splitlines() if l.startswith("PROBE ")]
            print(env, args, "rc", r.returncode, tail[-1] if tail else r.stdout[-800:], flush=True)
        except subprocess.TimeoutExpired:
            print(env, args, "TIMEOUT (hang)", flush=True)
            break        # a hung kernel may have wedged the GPU: stop


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        one(int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), sys.argv[5] == "1", sys.argv[6] == "1")
    else:
        sweep()
