#!/usr/bin/env python3
"""A/B harness for the MCL column kernel: builds the C3 link matrix once, then times pre-expansion and
the first iterations under different HH_MCL_* settings (read by hh_mcl_create).  Also checks that every
variant produces bit-identical iterates (the accumulation order does not depend on W / FLAT)."""
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench as B
from haphic_b200._lib import Context
from haphic_b200.links import LinkTable
from haphic_b200.mcl import Mcl


def main():
    a = B.parse_args()
    dev = torch.device("cuda", 0)
    asm, rank, in_nx, rec, _ = B.make_inputs(a, dev)
    ctx = Context(0)
    n = asm.n
    keep = np.ones(n, np.uint8)
    tab = LinkTable(ctx, asm.lengths, rank, in_nx, 500000, capacity_hint=int(0.45 * rec.shape[0]))
    tab.add(rec, asynchronous=True)
    tab.finish()
    index, _ = tab.linked_index(keep)
    mat = tab.to_matrix(keep, np.nonzero(index < 0)[0].astype(np.int32))
    tab.close()
    del rec
    configs = [dict(WINDOW=1)]
    sel = os.environ.get("TUNE_CONFIGS")
    if sel:
        configs = [configs[int(k)] for k in sel.split(",")]
    ref_hash = None
    for cfg in configs:
        for k, v in cfg.items():
            os.environ["HH_MCL_" + k] = str(v)
        t0 = time.perf_counter()
        mc = Mcl(mat)
        out = {"cfg": cfg, "preexp_ms": round(mc.preexp_ms, 1), "norm_ms": round(mc.normalize_ms, 2)}
        for r, iters in ((1.5, 200), (2.0, 200)):
            st = mc.run(r, iters, 1e-4)
            out["r{}".format(r)] = [round(float(x), 2) for x in st["iter_ms"]][:12]
            out["tot{}".format(r)] = round(float(st["iter_ms"].sum()), 1)
            out["rounds{}".format(r)] = st["rounds"]
            out["nnz{}".format(r)] = st["iter_nnz"].tolist()[-3:]
        fin = mc.result()
        h = hashlib.sha1(fin.indptr.tobytes() + fin.indices.tobytes() + fin.data.tobytes()).hexdigest()[:12]
        out["hash"] = h
        out["identical"] = (ref_hash is None) or (h == ref_hash)
        ref_hash = ref_hash or h
        out["wall_s"] = round(time.perf_counter() - t0, 2)
        print(json.dumps(out), flush=True)
        mc.close()
    mat.close()
    ctx.close()


if __name__ == "__main__":
    main()
