"""One pass of the hot path at C3 for ncu captures: partitioned link counting, dense pre-expansion, and the first two
iterations of two inflations (dense iteration 0, block-diagonal GEMM iteration 1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from haphic_b200 import synth
from haphic_b200._lib import Context
from haphic_b200.links import LinkTable, name_rank
from haphic_b200.mcl import Mcl

pairs = int(os.environ.get("PAIRS", "200000000"))
asm = synth.make_assembly(24, 50000, 20000, seed=12345)
rank = name_rank(asm.names)
in_nx = np.ones(asm.n, np.uint8)
rec = synth.make_pairs_range(asm, 0, pairs, seed=12346, device="cuda")
ctx = Context(0)
tab = LinkTable(ctx, asm.lengths, rank, in_nx, 500000, capacity_hint=int(0.45 * pairs))
tab.add(rec, asynchronous=True)
info = tab.finish()
keep = np.ones(asm.n, np.uint8)
index, _ = tab.linked_index(keep)
mat = tab.to_matrix(keep, np.nonzero(index < 0)[0].astype(np.int32))
mc = Mcl(mat, preexp=os.environ.get("PREEXP", "dense"))
for r in (2.0, 1.5):
    st = mc.run(r, 2, 1e-4)
    print(r, st["iter_ms"], flush=True)
print(mc.preexp)
