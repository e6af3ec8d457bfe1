"""GPU parity: Markov clustering (hh_matrix_* / hh_mcl_*) against the golden fixtures of the
reference and against the CPU oracle, through the C ABI.

Tolerances: the first normalisation M0 is bit-exact for integer link counts; every later float is
compared at 1e-6 relative where the sparsity patterns agree (the reference's own SpGEMM, Intel MKL,
has an unspecified accumulation order, so bit-equality of products is not defined); iteration
counts, convergence flags and final cluster assignments must be identical."""

import numpy as np
import torch
import pytest
import scipy.sparse as sp

from tests.util import canon, csc_from, load_golden, planted_blocks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from haphic_b200._lib import Context
    c = Context(0)
    yield c
    c.close()


def labels(clusters, n):
    lab = np.full(n, -1, np.int64)
    for c in clusters:
        lab[list(c)] = min(c)
    return lab


def compare_sparse(got, ref, rtol, what, max_pattern_diff=0, floor=0.0):
    """Same pattern (up to `max_pattern_diff` borderline entries below `floor`) and close values."""
    got, ref = canon(got), canon(ref)
    if (got.nnz == ref.nnz and np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)):
        assert np.allclose(got.data, ref.data, rtol=rtol, atol=0), what
        return
    d = abs(got - ref)
    only = (got != 0).astype(np.int8) - (ref != 0).astype(np.int8)
    only.eliminate_zeros()
    assert only.nnz <= max_pattern_diff, "{}: {} pattern differences".format(what, only.nnz)
    assert d.max() <= floor, "{}: max abs difference {}".format(what, d.max())


@pytest.mark.parametrize("tag", ["links_a", "block200", "block600", "block200_e3", "block200_e4"])
def test_mcl_matches_reference_golden(ctx, tag):
    from haphic_b200.links import LinkMatrix
    from haphic_b200.mcl import Mcl, interpret_result
    g = load_golden("mcl_{}.npz".format(tag))
    n = len(g["link_indptr"]) - 1
    link = csc_from(g, "link", n)
    mat = LinkMatrix.from_csc(ctx, link)
    back = mat.to_scipy()
    assert np.array_equal(back.indptr, link.indptr) and np.array_equal(back.indices, link.indices)
    assert np.array_equal(back.data, link.data)
    mc = Mcl(mat, expansion=int(g["expansion"]))
    m0 = mc.m0()
    ref0 = csc_from(g, "m0", n)
    assert np.array_equal(m0.indptr, ref0.indptr) and np.array_equal(m0.indices, ref0.indices)
    assert np.array_equal(m0.data, ref0.data), "first column normalisation must be bit-exact on integer counts"
    m1 = mc.m1()
    assert m1.shape == (n, n)
    assert np.allclose(m1, g["m1_dense"], rtol=2e-6, atol=1e-12)
    assert np.array_equal(m1 != 0, g["m1_dense"] != 0)
    pruning = float(g["pruning"])
    for r in g["inflations"].tolist():
        key = "r{}".format(str(r).replace(".", "p"))
        k = 1
        while key + "_iter{}_indptr".format(k) in g.files:
            st = mc.run(r, max_iter=k, pruning=pruning)
            assert st["rounds"] == k
            # an fp32 rounding difference eps in one iterate is amplified ~r times by the next inflation:
            # eps_k ~ eps * (1 + r + ... + r^(k-1)), eps = 2e-6 (sum of ~100 fp32 products, fma vs mul+add)
            rtol = 2e-6 * sum(r ** t for t in range(k))
            compare_sparse(mc.result(), csc_from(g, key + "_iter{}".format(k), n), rtol, (tag, key, k),
                           max_pattern_diff=2, floor=2 * pruning)
            k += 1
        st = mc.run(r, max_iter=200, pruning=pruning)
        assert st["rounds"] == int(g[key + "_niter"]), (key, st["rounds"], int(g[key + "_niter"]))
        assert st["converged"] == bool(g[key + "_converged"])
        fin = mc.result()
        clusters = interpret_result(fin)
        assert (clusters is not None) == bool(g[key + "_clusters_valid"])
        if clusters is not None:
            assert np.array_equal(labels(clusters, n), g[key + "_labels"])
        compare_sparse(fin, csc_from(g, key + "_final", n), 1e-5, (tag, key, "final"))
        # column-stochastic result
        sums = np.asarray(fin.sum(axis=0)).ravel()
        assert np.allclose(sums[sums != 0], 1.0, atol=1e-5)
    mc.close()
    mat.close()


@pytest.mark.parametrize("n_blocks,block,noise", [(150, 100, 0.5), (310, 100, 0.3)])
def test_mcl_matches_oracle_larger_geometry(ctx, n_blocks, block, noise):
    """n = 15,000 (16 row blocks per column) and n = 31,000 (32 row blocks): same iteration count,
    same clusters as the CPU oracle."""
    from haphic_b200.links import LinkMatrix
    from haphic_b200.mcl import Mcl, interpret_result
    from oracle import haphic_oracle as orc
    link, truth = planted_blocks(n_blocks, block, seed=n_blocks, noise=noise)
    n = link.shape[0]
    mat = LinkMatrix.from_csc(ctx, link)
    mc = Mcl(mat)
    m0 = orc.col_normalize_l1(link)
    got0 = mc.m0()
    assert np.array_equal(got0.indices, m0.indices) and np.array_equal(got0.data, m0.data)
    m1 = orc.expand(m0, 2)
    for r in (1.6, 2.0):
        ofin, rounds, conv = orc.mcl(m1, 2, r, 200, 1e-4)
        st = mc.run(r, 200, 1e-4)
        assert (st["rounds"], st["converged"]) == (rounds, conv)
        a = interpret_result(mc.result())
        b = orc.interpret_result(ofin)
        assert (a is None) == (b is None)
        if a is not None:
            assert np.array_equal(labels(a, n), labels(b, n))
    mc.close()
    mat.close()


def test_mcl_global_accumulator_path(ctx):
    """n > 57,600 does not fit the shared-memory accumulator: the global-memory variant must give
    the same clusters (planted, recoverable) and a column-stochastic result."""
    from haphic_b200.links import LinkMatrix
    from haphic_b200.mcl import Mcl, interpret_result
    link, truth = planted_blocks(1200, 50, seed=5, noise=0.0)
    n = link.shape[0]
    assert n == 60000
    mat = LinkMatrix.from_csc(ctx, link)
    mc = Mcl(mat)
    st = mc.run(2.0, 200, 1e-4)
    fin = mc.result()
    clusters = interpret_result(fin)
    assert clusters is not None and st["converged"]
    lab = labels(clusters, n)
    # every cluster lies inside one planted block
    for c in clusters:
        assert len(set(truth[list(c)].tolist())) == 1
    sums = np.asarray(fin.sum(axis=0)).ravel()
    assert np.allclose(sums, 1.0, atol=1e-5)
    mc.close()
    mat.close()


def test_mcl_step_interface_two_column_shards_equal_single(ctx):
    """Two column shards stepped side by side and exchanged through pack/unpack (what two ranks do
    with an all-gather) produce bit-identical iterates to the single-shard run."""
    from haphic_b200.links import LinkMatrix
    from haphic_b200.mcl import Mcl
    link, _ = planted_blocks(40, 60, seed=2, noise=0.4)
    n = link.shape[0]
    mat = LinkMatrix.from_csc(ctx, link)
    whole = Mcl(mat)
    st = whole.run(1.8, 200, 1e-4)
    want = canon(whole.result())
    cut = n // 3
    s0, s1 = Mcl(mat, col_lo=0, col_hi=cut), Mcl(mat, col_lo=cut, col_hi=n)
    m1 = np.concatenate([s0.m1(), s1.m1()], axis=1)
    assert np.array_equal(m1, whole.m1())
    s0.begin(1.8, 1e-4)
    s1.begin(1.8, 1e-4)
    rounds = 0
    replicated = False
    for it in range(200):
        n0, p0, d0 = s0.step(it)
        n1, p1, d1 = s1.step(it)
        if replicated:
            assert (n0, d0) == (n1, d1) == (int(st["iter_nnz"][it]), float(st["iter_delta"][it]))
        else:
            b0, b1 = s0.pack_flat(n0, 0), s1.pack(n1)          # both packings
            f = b0
            s1.unpack(0, cut, f[:cut], f[cut:cut + n0], f[cut + n0:cut + 2 * n0].view(torch.float32))
            s0.unpack(cut, n, *b1)
        s0.commit()
        s1.commit()
        if not replicated and it >= 3:       # what dist.sharded_mcl_run does for a tiny iterate: stop exchanging
            s0.set_block(0, n)
            s1.set_block(0, n)
            replicated = True
        rounds = it + 1
        if it > 1 and max(d0, d1) <= 1e-8:
            break
    assert replicated and rounds == st["rounds"]

    for s in (s0, s1):
        got = canon(s.result())
        assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
        assert np.array_equal(got.data, want.data)
    # the next mcl() call starts from the owned blocks again
    s0.begin(2.5, 1e-4)
    assert (s0.col_lo, s0.col_hi) == (0, cut)
    for o in (whole, s0, s1):
        o.close()
    mat.close()


def test_mcl_rejects_unsupported(ctx):
    from haphic_b200._lib import HHError
    from haphic_b200.links import LinkMatrix
    from haphic_b200.mcl import Mcl
    link, _ = planted_blocks(3, 10, seed=1)
    mat = LinkMatrix.from_csc(ctx, link)
    for bad in (1, 0, 9):                 # mkl_matrix_power needs k >= 2; the library stops at 8
        with pytest.raises(HHError):
            Mcl(mat, expansion=bad)
    mat.close()


@pytest.mark.parametrize("n,density,topN", [(400, 0.05, 10), (300, 0.004, 10), (64, 0.5, 5)])
def test_rank_sums_match_host(ctx, n, density, topN):
    """filter_fragments' rank-sum statistic (864-892): GPU == the dense host computation, including ties
    (integer link counts), fragments with fewer than topN linked neighbours and empty rows."""
    from itertools import combinations
    from haphic_b200.links import LinkMatrix
    rng = np.random.default_rng(n)
    a = np.triu((rng.random((n, n)) < density) * rng.integers(1, 6, size=(n, n)), 1).astype(np.float32)
    a = a + a.T
    a[7, :] = 0
    a[:, 7] = 0                      # an unlinked fragment
    mat = LinkMatrix.from_csc(ctx, sp.csc_matrix(a))
    got = mat.rank_sums(topN)
    mat.close()
    order = np.argsort(-a, axis=1, kind="stable")
    rank_of = np.empty((n, n), np.int64)
    rank_of[np.arange(n)[:, None], order] = np.arange(n)[None, :]
    want = np.zeros(n, np.int64)
    for x in range(n):
        top = order[x, :topN].tolist()
        want[x] = sum(min(rank_of[p, q], rank_of[q, p]) for p, q in combinations(top, 2))
    assert np.array_equal(got, want)


def info_nnz(tab):
    return tab.info.nnz_flank if tab.info is not None else -1


def test_global_accumulator_column_blocks_match_oracle(ctx):
    """n = 60,000 > 57,600 rows: the column accumulator no longer fits shared memory and every column kernel (normalise,
    expansion, dense iteration 0) runs on the L2-resident global accumulator.  On a realistic matrix (synthetic Hi-C stream,
    not planted blocks) three column blocks are compared with the CPU oracle: counts and M0 bit-exact, the block of
    M1 = M0 . M0 within 2e-6 with an identical pattern, and the first pruned iterate of the block (inflate, normalise,
    prune, normalise: column-local, 2030-2042)."""
    import torch
    from haphic_b200 import synth
    from haphic_b200.links import LinkTable, name_rank
    from haphic_b200.mcl import Mcl
    from oracle import haphic_oracle as orc
    asm = synth.make_assembly(24, 60000, 20000, seed=77)
    rank = name_rank(asm.names)
    in_nx = np.ones(asm.n, np.uint8)
    rec = synth.make_pairs_range(asm, 0, 8_000_000, seed=78, device="cuda")
    tab = LinkTable(ctx, asm.lengths, rank, in_nx, 500000)
    tab.add(rec)
    tab.finish()
    keep = np.ones(asm.n, np.uint8)
    index, n_linked = tab.linked_index(keep)
    tail = np.nonzero(index < 0)[0].astype(np.int32)
    mat = tab.to_matrix(keep, tail)
    n = mat.n
    assert n == 60000
    ref = orc.count_links_c(rec.cpu().numpy(), asm.lengths, rank, in_nx, 500000)
    link, oindex = orc.dict_to_matrix(ref["flank_keys"], ref["flank_vals"], keep, tail_order=tail.tolist())
    got = mat.to_scipy()
    assert got.nnz == link.nnz and np.array_equal(got.indptr, link.indptr), (got.nnz, link.nnz, int(info_nnz(tab)))
    assert np.array_equal(got.indices, link.indices), int((got.indices != link.indices).sum())
    assert np.array_equal(got.data, link.data), int((got.data != link.data).sum())
    m0 = orc.col_normalize_l1(link)
    first = True
    for lo, hi in [(0, 96), (30000, 30096), (59904, 60000)]:
        mc = Mcl(mat, col_lo=lo, col_hi=hi, preexp="sparse")
        assert not mc.preexp["mode"] == "dense"
        if first:
            g0 = mc.m0()
            assert np.array_equal(g0.indices, m0.indices) and np.array_equal(g0.data, m0.data)
            first = False
        want = sp.csc_matrix(m0 @ m0[:, lo:hi], dtype=np.float32)
        m1 = mc.m1()
        wd = want.toarray()
        assert np.array_equal(m1 != 0, wd != 0)
        nz = wd != 0
        assert (np.abs(m1[nz].astype(np.float64) - wd[nz]) / wd[nz]).max() <= 2e-6
        # iteration 0 of inflation 2.0 on the block
        mc.begin(2.0, 1e-4)
        nnz, _, _ = mc.step(0)
        ln, idx, val = mc.pack(nnz)
        ln, idx, val = ln.cpu().numpy(), idx.cpu().numpy(), val.cpu().numpy()
        gotp = sp.csc_matrix((val, idx, np.concatenate([[0], np.cumsum(ln)])), shape=(n, hi - lo))
        wantp = orc.prune(orc.inflate(want, 2.0), 1e-4)
        compare_sparse(gotp, wantp, 4e-6, ("block", lo), max_pattern_diff=2, floor=2e-4)
        mc.close()
    mat.close()
    tab.close()
