"""GPU parity at BASELINE.json configs[1] size (C2: 10k contigs / 50M synthetic pairs, full MCL to convergence), against
the CPU oracle where it can follow in minutes:

  * link counting: every counter of the 50M-record stream bit-exact against oracle.count_links_numpy
    (full / flank / HT dicts in insertion order, per-contig totals);
  * dict_to_matrix and the first normalisation M0: bit-exact against the oracle;
  * pre-expansion M1: both engines (tensor-core GEMM and Gustavson) within 2e-6 of the exact fp64 product of M0, identical
    pattern;
  * Markov clustering: clusters and convergence flags equal to the oracle's mcl() for inflations 2.0 / 3.0 (pre-expansion of
    the oracle = the reference's dense mode, an fp32 BLAS product), and for the reference's default 20-inflation sweep
    (1.1 .. 3.0) identical between the sequential-fp32 engine and the tensor-core engines of this library (the oracle needs
    hours for the low inflations).  Iteration counts agree within one round: every engine is within 1e-6 of the exact products
    but rounds differently, and the reference's stopping rule max(|M - L| - 1e-5 |L|) <= 1e-8 sits at fp32 resolution -- the
    same freedom the reference's own MKL and NumPy modes have against each other.
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_CONTIGS, N_CHR, MEAN_LEN, N_PAIRS, SEED = 10000, 16, 30000, 50_000_000, 12345


@pytest.fixture(scope="module")
def c2():
    from haphic_b200 import synth
    from haphic_b200._lib import Context
    from haphic_b200.links import LinkTable, name_rank
    ctx = Context(0)
    asm = synth.make_assembly(N_CHR, N_CONTIGS, MEAN_LEN, seed=SEED)
    rank = name_rank(asm.names)
    in_nx = np.ones(asm.n, np.uint8)
    rec = synth.make_pairs_range(asm, 0, N_PAIRS, seed=SEED + 1, device="cuda")
    tab = LinkTable(ctx, asm.lengths, rank, in_nx, 500000, capacity_hint=int(0.6 * min(N_PAIRS, asm.n * (asm.n - 1) // 2)))
    tab.add(rec, asynchronous=True)
    info = tab.finish()
    keep = np.ones(asm.n, np.uint8)
    index, n_linked = tab.linked_index(keep)
    tail = np.nonzero(index < 0)[0].astype(np.int32)
    mat = tab.to_matrix(keep, tail)
    d = dict(ctx=ctx, asm=asm, rank=rank, in_nx=in_nx, rec=rec, tab=tab, info=info, keep=keep, index=index, tail=tail, mat=mat)
    yield d
    mat.close()
    tab.close()
    ctx.close()


@pytest.fixture(scope="module")
def c2_oracle(c2):
    from oracle import haphic_oracle as orc
    # the C restatement of the counting loop (oracle/haphic_oracle.c, pinned by the reference's golden fixtures in
    # tests/test_oracle_golden.py): seconds for 50M records where the numpy version needs minutes
    ref = orc.count_links_c(c2["rec"].cpu().numpy(), c2["asm"].lengths, c2["rank"], c2["in_nx"], 500000)
    link, oindex = orc.dict_to_matrix(ref["flank_keys"], ref["flank_vals"], c2["keep"], tail_order=c2["tail"].tolist())
    m0 = orc.col_normalize_l1(link)
    return dict(ref=ref, link=link, index=oindex, m0=m0)


def test_c2_link_counters_bit_exact(c2, c2_oracle):
    ref, tab, n = c2_oracle["ref"], c2["tab"], c2["asm"].n
    assert c2["info"].n_used == ref["n_used"] and c2["info"].nnz_full == len(ref["full_vals"])
    got = tab.fetch()
    assert np.array_equal(np.stack([got["key_i"], got["key_j"]], 1), ref["full_keys"])       # dict insertion order
    assert np.array_equal(got["full"].astype(np.int64), ref["full_vals"])
    assert np.array_equal(got["first_full"].astype(np.int64), ref["full_first"])
    sel = np.nonzero(got["flank"] > 0)[0]
    sel = sel[np.argsort(got["first_flank"][sel], kind="stable")]
    assert np.array_equal(np.stack([got["key_i"][sel], got["key_j"][sel]], 1), ref["flank_keys"])
    assert np.array_equal(got["flank"][sel].astype(np.int64), ref["flank_vals"])
    assert np.array_equal(got["ht"].astype(np.int64), ref["ht"])          # HT_link_dict: HH / HT / TH / TT of every pair
    assert np.array_equal(tab.fetch_ctg(), ref["ctg_link_total"])


def test_c2_matrix_and_first_normalisation_bit_exact(c2, c2_oracle):
    from haphic_b200.mcl import Mcl
    assert np.array_equal(np.where(c2["index"] >= 0, c2["index"], c2_oracle["index"]), c2_oracle["index"])
    got = c2["mat"].to_scipy()
    link = c2_oracle["link"]
    assert np.array_equal(got.indptr, link.indptr) and np.array_equal(got.indices, link.indices)
    assert np.array_equal(got.data, link.data)
    mc = Mcl(c2["mat"], preexp="sparse", col_lo=0, col_hi=64)       # M0 is built whole whatever the owned block is
    m0 = mc.m0()
    mc.close()
    ref0 = c2_oracle["m0"]
    assert np.array_equal(m0.indices, ref0.indices) and np.array_equal(m0.data, ref0.data)


def test_c2_preexpansion_both_engines_vs_exact_product(c2, c2_oracle):
    from haphic_b200.mcl import Mcl
    m0 = c2_oracle["m0"]
    d = m0.toarray().astype(np.float64)
    exact = d @ d
    nz = exact != 0
    worst = {}
    for engine in ("dense", "sparse"):
        mc = Mcl(c2["mat"], preexp=engine)
        assert mc.preexp["mode"] == engine
        m1 = mc.m1()
        mc.close()
        assert np.array_equal(m1 != 0, nz), engine
        rel = np.abs(m1.astype(np.float64)[nz] - exact[nz]) / exact[nz]
        worst[engine] = float(rel.max())
        del m1, rel
    assert worst["dense"] <= 2e-6 and worst["sparse"] <= 4e-6, worst


def test_c2_mcl_matches_oracle_and_engines_agree_on_the_default_sweep(c2, c2_oracle):
    from haphic_b200.mcl import Mcl, inflation_values, interpret_result
    from oracle import haphic_oracle as orc
    n = c2["asm"].n

    def labels(clusters):
        lab = np.full(n, -1, np.int64)
        for c in clusters:
            lab[list(c)] = min(c)
        return lab

    import os
    os.environ["HH_MCL_BLOCKGEMM"] = "0"
    sparse = Mcl(c2["mat"], preexp="sparse")           # sequential fp32 everywhere: the oracle's arithmetic
    os.environ["HH_MCL_BLOCKGEMM"] = "1"
    dense = Mcl(c2["mat"], preexp="dense")             # tensor cores for the pre-expansion and the dense component blocks
    import scipy.sparse as sp
    d0 = c2_oracle["m0"].toarray()
    m1 = sp.csc_matrix(d0 @ d0)                         # numpy.linalg.matrix_power(M0, 2) of the reference's dense mode (2149)
    del d0
    try:
        for r in (2.0, 3.0):
            ofin, rounds, conv = orc.mcl(m1, 2, r, 200, 1e-4)
            want = labels(orc.interpret_result(ofin))
            for eng in (sparse, dense):
                st = eng.run(r, 200, 1e-4)
                assert st["converged"] == conv and abs(st["rounds"] - rounds) <= 1, (r, eng.preexp["mode"], st["rounds"], rounds)
                assert np.array_equal(labels(interpret_result(eng.result())), want), (r, eng.preexp["mode"])
        del m1
        # the reference's default sweep: 20 inflations 1.1 .. 3.0 (HapHiC_cluster.py:2139-2155, 2699-2705)
        for r in inflation_values(1.1, 3.0, 0.1):
            os.environ["HH_MCL_BLOCKGEMM"] = "1"
            a = dense.run(float(r), 200, 1e-4)
            ca = interpret_result(dense.result())
            os.environ["HH_MCL_BLOCKGEMM"] = "0"
            b = sparse.run(float(r), 200, 1e-4)
            cb = interpret_result(sparse.result())
            assert a["converged"] == b["converged"] and abs(a["rounds"] - b["rounds"]) <= 1, (str(r), a["rounds"], b["rounds"])
            assert (ca is None) == (cb is None)
            if ca is not None:
                assert np.array_equal(labels(ca), labels(cb)), str(r)
    finally:
        os.environ.pop("HH_MCL_BLOCKGEMM", None)
    dense.close()
    sparse.close()
