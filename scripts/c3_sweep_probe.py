import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from haphic_b200 import synth
from haphic_b200._lib import Context
from haphic_b200.links import LinkTable, name_rank
from haphic_b200.mcl import Mcl, inflation_values
asm = synth.make_assembly(24, 50000, 20000, seed=12345)
rank = name_rank(asm.names); in_nx = np.ones(asm.n, np.uint8)
rec = synth.make_pairs_range(asm, 0, 200_000_000, seed=12346, device="cuda")
ctx = Context(0)
tab = LinkTable(ctx, asm.lengths, rank, in_nx, 500000, capacity_hint=int(0.45*200e6))
tab.add(rec, asynchronous=True); tab.finish()
keep = np.ones(asm.n, np.uint8)
index,_ = tab.linked_index(keep)
mat = tab.to_matrix(keep, np.nonzero(index<0)[0].astype(np.int32))
for chunk in (os.environ.get("CHUNKS","2").split(",")):
    os.environ["HH_GEMM_CHUNK"]=chunk
    mc = Mcl(mat, preexp="dense")
    print("chunk", chunk, {k:(round(v,2) if isinstance(v,float) else v) for k,v in mc.preexp.items()}, flush=True)
    if chunk != os.environ.get("CHUNKS","2").split(",")[-1]: mc.close()
for r in inflation_values(1.1, 3.0, 0.1)[::int(os.environ.get("STRIDE","3"))]:
    st = mc.run(float(r), 200, 1e-4)
    print("r", r, "rounds", st["rounds"], "ms", round(float(st["iter_ms"].sum()),1), [round(float(x),1) for x in st["iter_ms"][:8]], st["iter_nnz"][:6].tolist(), flush=True)
