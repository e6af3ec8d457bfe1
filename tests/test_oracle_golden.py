"""Pin the CPU oracle against the golden fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""

import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import haphic_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def name_rank_of(names):
    order = sorted(range(len(names)), key=lambda i: names[i])
    rank = np.empty(len(names), dtype=np.int64)
    rank[order] = np.arange(len(names))
    return rank


def csc_from(g, prefix, n):
    return sp.csc_matrix((g[prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]), shape=(n, n))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_link_counts_match_reference(tag):
    g = load("links_{}.npz".format(tag))
    names = g["names"].tolist()
    rank = name_rank_of(names)
    flank_bp = int(g["flank_kb"]) * 1000
    res = orc.count_links_numpy(g["pairs"], g["lengths"], rank, g["in_nx"], flank_bp)
    for k in ("full_keys", "full_vals", "flank_keys", "flank_vals", "HT_keys", "HT_vals", "ctg_link_ids",
              "ctg_link_vals", "clm_keys", "clm_offsets", "clm_vals"):
        assert np.array_equal(res[k], g[k]), k
    assert orc.clm_text(res["clm_keys"], res["clm_offsets"], res["clm_vals"], names) == str(g["clm_text"])
    # the per-pair loop restatement agrees with the vectorised one (first 5000 records)
    sub = g["pairs"][:5000]
    full, flank_d, HT, clm, ctg = orc.count_links_loop(sub, g["lengths"], rank, g["in_nx"], flank_bp)
    r2 = orc.count_links_numpy(sub, g["lengths"], rank, g["in_nx"], flank_bp)
    assert np.array_equal(np.array(list(full.keys()), dtype=np.int32).reshape(-1, 2), r2["full_keys"])
    assert list(full.values()) == r2["full_vals"].tolist()
    assert np.array_equal(np.array(list(flank_d.keys()), dtype=np.int32).reshape(-1, 2), r2["flank_keys"])
    assert list(flank_d.values()) == r2["flank_vals"].tolist()
    assert np.array_equal(np.array(list(HT.keys()), dtype=np.int32).reshape(-1, 4), r2["HT_keys"])
    assert list(HT.values()) == r2["HT_vals"].tolist()
    assert list(ctg.keys()) == r2["ctg_link_ids"].tolist() and list(ctg.values()) == r2["ctg_link_vals"].tolist()
    assert sum((v for v in clm.values()), []) == r2["clm_vals"].tolist()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_c_oracle_link_counts_match_reference(tag):
    """oracle/haphic_oracle.c (the single-core C port timed as a CPU baseline) against the reference's dicts."""
    g = load("links_{}.npz".format(tag))
    names = g["names"].tolist()
    rank = name_rank_of(names)
    res = orc.count_links_c(g["pairs"], g["lengths"], rank, g["in_nx"], int(g["flank_kb"]) * 1000)
    for k in ("full_keys", "full_vals", "flank_keys", "flank_vals"):
        assert np.array_equal(res[k], g[k]), k
    want = {tuple(k): int(v) for k, v in zip(g["HT_keys"].tolist(), g["HT_vals"].tolist())}
    got = {}
    for (i, j), row in zip(res["full_keys"].tolist(), res["ht"].tolist()):
        for c, v in enumerate(row):
            if v:
                got[(i, c >> 1, j, c & 1)] = v
    assert got == want
    tot = np.zeros(len(names), np.int64)
    tot[g["ctg_link_ids"]] = g["ctg_link_vals"]
    assert np.array_equal(res["ctg_link_total"], tot)
    ref = orc.count_links_numpy(g["pairs"], g["lengths"], rank, g["in_nx"], int(g["flank_kb"]) * 1000)
    assert res["n_used"] == ref["n_used"] and np.array_equal(res["full_first"], ref["full_first"])
    assert np.array_equal(res["flank_first"], ref["flank_first"])


def test_normalize_by_nlinks_matches_reference():
    g = load("links_b.npz")
    rank = name_rank_of(g["names"].tolist())
    res = orc.count_links_numpy(g["pairs"], g["lengths"], rank, g["in_nx"], int(g["flank_kb"]) * 1000)
    v = orc.normalize_by_nlinks(res["flank_keys"], res["flank_vals"], res["ctg_link_total"])
    assert np.array_equal(v, g["flank_norm_vals"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_dict_to_matrix_matches_reference(tag):
    g = load("links_{}.npz".format(tag))
    vals = g["flank_norm_vals"] if "flank_norm_vals" in g.files else g["flank_vals"]
    # the reference appends kept-but-unlinked fragments in set order: take that order from the fixture
    idx_ref = g["matrix_index"]
    tail = np.argsort(np.where(idx_ref >= 0, idx_ref, np.iinfo(np.int32).max), kind="stable").tolist()
    m, index = orc.dict_to_matrix(g["flank_keys"], vals, g["filtered"], tail_order=tail)
    assert np.array_equal(index, idx_ref)
    n = m.shape[0]
    ref = csc_from(g, "link", n)
    assert np.array_equal(m.indptr, ref.indptr) and np.array_equal(m.indices, ref.indices)
    assert np.array_equal(m.data, ref.data)


@pytest.mark.parametrize("tag", ["links_a", "block200", "block600"])
def test_mcl_matches_reference(tag):
    g = load("mcl_{}.npz".format(tag))
    n = len(g["link_indptr"]) - 1
    link = csc_from(g, "link", n)
    m0 = orc.col_normalize_l1(link)
    ref0 = csc_from(g, "m0", n)
    assert np.array_equal(m0.indices, ref0.indices) and np.array_equal(m0.data, ref0.data)
    e = int(g["expansion"])
    m1 = orc.expand(m0, e)
    assert np.allclose(np.asarray(m1.todense()), g["m1_dense"], rtol=1e-6, atol=0)
    for r in g["inflations"].tolist():
        key = "r{}".format(str(r).replace(".", "p"))
        trace = []
        fin, rounds, conv = orc.mcl(m1, e, r, 200, float(g["pruning"]), trace=trace)
        assert rounds == int(g[key + "_niter"]) and conv == bool(g[key + "_converged"])
        assert rounds == int(g[key + "_niter_dense"])
        k = 1
        while key + "_iter{}_indptr".format(k) in g.files:
            refk = csc_from(g, key + "_iter{}".format(k), n)
            refk.eliminate_zeros()
            mk = trace[k - 1].copy()
            mk.eliminate_zeros()
            assert np.array_equal(mk.indptr, refk.indptr) and np.array_equal(mk.indices, refk.indices), (key, k)
            assert np.allclose(mk.data, refk.data, rtol=1e-6, atol=0), (key, k)
            k += 1
        reff = csc_from(g, key + "_final", n)
        reff.eliminate_zeros()
        f2 = fin.copy()
        f2.eliminate_zeros()
        assert np.array_equal(f2.indices, reff.indices) and np.allclose(f2.data, reff.data, rtol=1e-6, atol=0)
        clusters = orc.interpret_result(fin)
        assert (clusters is not None) == bool(g[key + "_clusters_valid"])
        if clusters is not None:
            assert np.array_equal(orc.labels_from_clusters(clusters, n), g[key + "_labels"])


def test_inflation_sweep_names():
    sw = orc.inflation_sweep(1.1, 3.0, 0.1)
    assert len(sw) == 20 and str(sw[0]) == "1.1" and str(sw[-1]) == "3.0" and str(sw[9]) == "2.0"
