"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: column-block planning, the
variable-length all-gather, the link-table merge protocol and the sharded MCL loop.  The compute
engines are CPU stand-ins built on the oracle; the exchange code is the product's
(haphic_b200/dist.py) unchanged."""

import os
import socket
import traceback

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import planted_blocks


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleShard:
    """begin/step/pack/unpack/commit over a column block, computed with the CPU oracle."""

    def __init__(self, link, lo, hi):
        from oracle import haphic_oracle as orc
        self.orc = orc
        self.n = link.shape[0]
        self.lo, self.hi = lo, hi
        self.own = (lo, hi)
        self.set_block_calls = 0
        m0 = orc.col_normalize_l1(link)
        self.m1_block = sp.csc_matrix(orc.expand(m0, 2)[:, lo:hi])
        self.cur = None

    def begin(self, inflation, pruning):
        self.r, self.p = inflation, pruning
        self.cur = None
        self.blocks = {}
        self.lo, self.hi = self.own

    def set_block(self, lo, hi):
        assert self.cur is not None
        self.lo, self.hi = lo, hi
        self.set_block_calls += 1

    def step(self, it):
        orc = self.orc
        src = self.m1_block if it == 0 else sp.csc_matrix(self.cur @ self.cur[:, self.lo:self.hi], dtype=np.float32)
        blk = orc.prune(orc.inflate(src, self.r), self.p)
        delta = orc.convergence_delta(blk, self.cur[:, self.lo:self.hi]) if it > 0 else 0.0
        self.blocks = {(self.lo, self.hi): blk}
        return int(blk.nnz), 0, float(delta)

    def pack(self, nnz):
        blk = self.blocks[(self.lo, self.hi)]
        blk.sort_indices()
        return (torch.from_numpy(np.diff(blk.indptr).astype(np.int32)), torch.from_numpy(blk.indices.astype(np.int32)),
                torch.from_numpy(blk.data.astype(np.float32)))

    def unpack(self, lo, hi, ln, idx, val):
        indptr = np.concatenate([[0], np.cumsum(ln.numpy())])
        self.blocks[(lo, hi)] = sp.csc_matrix((val.numpy(), idx.numpy(), indptr), shape=(self.n, hi - lo))

    def commit(self):
        keys = sorted(self.blocks)
        assert keys[0][0] == 0 and keys[-1][1] == self.n
        self.cur = sp.csc_matrix(sp.hstack([self.blocks[k] for k in keys]), dtype=np.float32)


class FakeTable:
    """finish/export/merge with the link-table semantics (adds + first-seen minima)."""

    def __init__(self, entries, totals, n_rec, n_used):
        self.e = {(int(r[0]), int(r[1])): r[2:].astype(np.int64).copy() for r in entries}
        self.tot = totals.astype(np.int64).copy()
        self.n_rec, self.n_used = n_rec, n_used

    def finish(self):
        pass

    def export(self):
        arr = np.array([[k[0], k[1]] + v.tolist() for k, v in self.e.items()], dtype=np.int64).reshape(-1, 9)
        return torch.from_numpy(arr), torch.from_numpy(self.tot.copy()), self.n_rec, self.n_used

    def merge(self, ent, tot, n_rec, n_used):
        for r in ent.numpy():
            k = (int(r[0]), int(r[1]))
            v = r[2:]
            if k in self.e:
                cur = self.e[k]
                cur[0] += v[0]; cur[1] += v[1]; cur[2] = min(cur[2], v[2]); cur[3] = min(cur[3], v[3]); cur[4:] += v[4:]
            else:
                self.e[k] = v.astype(np.int64).copy()
        self.tot += tot.numpy()
        self.n_rec += n_rec
        self.n_used += n_used


class FakeRoutedTable:
    """route / add_routed / finish_partition / export / adopt with the library's semantics, on the CPU oracle."""

    def __init__(self, lengths, rk, nx, flank):
        self.args = (lengths, rk, nx, flank)
        self.rec = []
        self.pos = []
        self.whole = None

    @staticmethod
    def owner(a, b, world):
        lo, hi = min(a, b), max(a, b)
        return (lo * 1000003 + hi) % world

    def route(self, rec, stream_lo, world):
        r = rec.numpy()
        n = len(self.args[0])
        ok = (r[:, 0] != r[:, 2]) & (r[:, 0] >= 0) & (r[:, 0] < n) & (r[:, 2] >= 0) & (r[:, 2] < n)
        dest = np.array([self.owner(int(a), int(b), world) if o else -1 for a, b, o in zip(r[:, 0], r[:, 2], ok)])
        pos = np.arange(len(r)) + stream_lo
        order = np.concatenate([np.nonzero(dest == d)[0] for d in range(world)]).astype(np.int64)
        counts = [int((dest == d).sum()) for d in range(world)]
        return torch.from_numpy(r[order]), torch.from_numpy(pos[order].astype(np.int32)), counts

    def add_routed(self, rec, pos):
        self.rec.append(rec.numpy())
        self.pos.append(pos.numpy())

    def finish_partition(self):
        from oracle import haphic_oracle as orc
        rec = np.concatenate(self.rec) if self.rec else np.zeros((0, 4), np.int32)
        pos = np.concatenate(self.pos) if self.pos else np.zeros(0, np.int32)
        o = np.argsort(pos, kind="stable")         # the oracle counts in stream order
        rec, pos = rec[o], pos[o]
        r = orc.count_links_numpy(rec, *self.args)
        ent = np.zeros((len(r["full_vals"]), 9), np.int64)
        ent[:, 0:2] = r["full_keys"]
        ent[:, 2] = r["full_vals"]
        ent[:, 4] = pos[r["full_first"]]
        ent[:, 5] = 0xFFFFFFFF
        fl = {tuple(k): (v, f) for k, v, f in zip(r["flank_keys"].tolist(), r["flank_vals"].tolist(), r["flank_first"].tolist())}
        for e, k in enumerate(r["full_keys"].tolist()):
            if tuple(k) in fl:
                ent[e, 3] = fl[tuple(k)][0]
                ent[e, 5] = pos[fl[tuple(k)][1]]
        self.part = (ent, r["ctg_link_total"].astype(np.int64), r["n_used"])

        class Info:
            n_used = r["n_used"]
        return Info()

    def export(self):
        return torch.from_numpy(self.part[0]), torch.from_numpy(self.part[1].copy()), 0, self.part[2]

    def adopt(self, entries, tot, n_rec, n_used, stream_end):
        self.whole = (entries.numpy().copy(), tot.numpy().copy(), n_rec, n_used, stream_end)
        return self.whole


def _worker(rank, world, port, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from haphic_b200 import dist as hdist
        from oracle import haphic_oracle as orc
        # 1) variable-length all-gather
        t = torch.arange(3 + 4 * rank, dtype=torch.int32) + 100 * rank
        got = hdist.allgather_varlen(t)
        assert [len(g) for g in got] == [3, 7]
        assert got[1][0].item() == 100 and got[0][2].item() == 2
        empty = hdist.allgather_varlen(torch.zeros(0 if rank == 0 else 2, dtype=torch.float32))
        assert [len(g) for g in empty] == [0, 2]
        # 2) link-table merge: two shards of one stream -> both ranks hold the single-stream table
        rng = np.random.default_rng(5)
        pairs = np.stack([rng.integers(0, 30, 4000), rng.integers(0, 1000, 4000), rng.integers(0, 30, 4000),
                          rng.integers(0, 1000, 4000)], 1)
        lengths = np.full(30, 1000)
        rk = np.arange(30)
        nx = np.ones(30, np.uint8)

        def table_of(sub, off):
            r = orc.count_links_numpy(sub, lengths, rk, nx, 100)
            ent = np.zeros((len(r["full_vals"]), 9), np.int64)
            ent[:, 0:2] = r["full_keys"]
            ent[:, 2] = r["full_vals"]
            ent[:, 4] = r["full_first"] + off
            ent[:, 5] = 0xFFFFFFFF
            fl = {tuple(k): (v, f) for k, v, f in zip(r["flank_keys"].tolist(), r["flank_vals"].tolist(), r["flank_first"].tolist())}
            for e, k in enumerate(r["full_keys"].tolist()):
                if tuple(k) in fl:
                    ent[e, 3] = fl[tuple(k)][0]
                    ent[e, 5] = fl[tuple(k)][1] + off
            return FakeTable(ent, r["ctg_link_total"], len(sub), r["n_used"])

        half = 2000
        mine = table_of(pairs[:half], 0) if rank == 0 else table_of(pairs[half:], half)
        hdist.merge_link_tables(mine)
        whole = table_of(pairs, 0)
        assert set(mine.e) == set(whole.e)
        for k in whole.e:
            assert np.array_equal(mine.e[k][:4], whole.e[k][:4]), k
        assert np.array_equal(mine.tot, whole.tot) and mine.n_rec == 4000 and mine.n_used == whole.n_used
        # 2b) routed counting: all-to-all of records, disjoint partitions, adopted union == single-stream table
        rt = FakeRoutedTable(lengths, rk, nx, 100)
        shard = pairs[:half] if rank == 0 else pairs[half:]
        ent, tot, n_rec, n_used, s_end = hdist.routed_link_build(rt, torch.from_numpy(shard.astype(np.int32)), 0 if rank == 0 else half)
        assert (n_rec, n_used, s_end) == (4000, whole.n_used, 4000) and np.array_equal(tot, whole.tot)
        got = {(int(r[0]), int(r[1])): r[2:] for r in ent}
        assert set(got) == set(whole.e) and len(ent) == len(whole.e)
        for k in whole.e:
            assert np.array_equal(got[k][:4], whole.e[k][:4]), k
        # 3) sharded MCL == single MCL (rounds, convergence, final matrix), uneven blocks included
        link, _ = planted_blocks(6, 30, seed=3, noise=0.5)
        n = link.shape[0]
        m1 = orc.expand(orc.col_normalize_l1(link), 2)
        for blocks in (hdist.column_blocks(n, world), [(0, 50), (50, n)]):
            eng = OracleShard(link, *blocks[rank])
            st = hdist.sharded_mcl_run(eng, 2.0, 100, 1e-4, blocks)
            fin, rounds, conv = orc.mcl(m1, 2, 2.0, 100, 1e-4)
            assert (st["rounds"], st["converged"]) == (rounds, conv)
            assert abs(eng.cur - fin).max() < 1e-7
            assert st["iter_nnz"][-1] == fin.nnz
            assert eng.set_block_calls == 1 and (eng.lo, eng.hi) == (0, n)      # the tail ran replicated, without exchange
        # 4) the inflation-parallel sweep: iteration 0 of every inflation sharded, then inflation k on rank k mod world alone
        infl = [1.6, 2.0, 2.6]
        want = [orc.mcl(m1, 2, r, 100, 1e-4) for r in infl]
        for blocks in (hdist.column_blocks(n, world), [(0, 50), (50, n)]):
            eng = OracleShard(link, *blocks[rank])
            seen = []
            stats = hdist.sharded_mcl_sweep(eng, infl, 100, 1e-4, blocks,
                                            on_result=lambda k, r, e: seen.append((k, float(abs(e.cur - want[k][0]).max()))))
            assert [(st["rounds"], st["converged"], st["owner"]) for st in stats] == [(w[1], w[2], k % world) for k, w in enumerate(want)]
            assert [st["iter_nnz"][-1] for st in stats] == [w[0].nnz for w in want]
            assert [k for k, _ in seen] == [k for k in range(len(infl)) if k % world == rank]
            assert all(d < 1e-7 for _, d in seen)
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:
        q.put((rank, traceback.format_exc()))


def test_column_blocks():
    from haphic_b200 import dist as hdist
    assert hdist.column_blocks(10, 3) == [(0, 3), (3, 6), (6, 10)]
    b = hdist.balanced_column_blocks([1, 1, 1, 1, 10, 1, 1, 1, 1, 1], 2)
    assert b[0][0] == 0 and b[-1][1] == 10 and b[0][1] == b[1][0]
    assert hdist.balanced_column_blocks([], 2) == [(0, 0), (0, 0)]


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", "rank {}: {}".format(rank, msg)
