"""Multi-GPU parity check (run under torchrun, one process per GPU): the inflation-parallel sharded sweep of
haphic_b200.dist must give, for every inflation, the bytes a single GPU owning every column gives.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py

Every rank builds the same link matrix (C2 shape by default: 5k contigs / 20M pairs), runs the sharded sweep with its column
block, and the owner of each inflation compares the result with `Mcl.run` on a whole-matrix engine of its own device."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from haphic_b200 import dist as hdist
from haphic_b200 import synth
from haphic_b200._lib import Context
from haphic_b200.links import LinkTable, name_rank
from haphic_b200.mcl import Mcl


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    contigs = int(os.environ.get("CONTIGS", "5000"))
    pairs = int(os.environ.get("PAIRS", "20000000"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    asm = synth.make_assembly(12, contigs, 20000, seed=4321)
    rec = synth.make_pairs_range(asm, 0, pairs, seed=4322, device="cuda")
    ctx = Context(local)
    tab = LinkTable(ctx, asm.lengths, name_rank(asm.names), np.ones(asm.n, np.uint8), 500000, capacity_hint=int(0.6 * pairs))
    tab.add(rec)
    tab.finish()
    keep = np.ones(asm.n, np.uint8)
    index, _ = tab.linked_index(keep)
    mat = tab.to_matrix(keep, np.nonzero(index < 0)[0].astype(np.int32))
    inflations = [1.3, 1.5, 2.0, 2.4, 3.0]
    blocks = hdist.column_blocks(mat.n, world)
    shard = Mcl(mat, col_lo=blocks[rank][0], col_hi=blocks[rank][1])
    whole = Mcl(mat)
    report = []

    def check(k, r, eng):
        got = eng.result()
        st = whole.run(r, 200, 1e-4)
        want = whole.result()
        same = (np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
                and np.array_equal(got.data, want.data))
        report.append({"inflation": r, "rank": rank, "bit_equal": bool(same), "single_rounds": st["rounds"]})

    stats = hdist.sharded_mcl_sweep(shard, inflations, 200, 1e-4, blocks, on_result=check)
    for rep in report:
        rep["sharded_rounds"] = stats[inflations.index(rep["inflation"])]["rounds"]
    allrep = [None] * world
    dist.all_gather_object(allrep, report)
    if rank == 0:
        flat = sorted([x for r in allrep for x in r], key=lambda x: x["inflation"])
        ok = all(x["bit_equal"] and x["sharded_rounds"] == x["single_rounds"] for x in flat) and len(flat) == len(inflations)
        print("DISTCHECK " + json.dumps({"world": world, "n": mat.n, "engine": shard.preexp["mode"], "ok": ok, "inflations": flat}), flush=True)
    shard.close()
    whole.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
