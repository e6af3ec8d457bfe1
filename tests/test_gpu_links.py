"""GPU parity: link counting (hh_links_*) against the golden fixtures of the reference and
against the CPU oracle, through the C ABI.  Integer results must be bit-exact."""

import numpy as np
import pytest

from tests.util import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from haphic_b200._lib import Context
    c = Context(0)
    yield c
    c.close()


def rank_of(names):
    from haphic_b200.links import name_rank
    return name_rank(list(names))


def ht_dict(keys_i, keys_j, ht):
    d = {}
    for e in range(len(keys_i)):
        for c in range(4):
            v = int(ht[e, c])
            if v:
                d[(int(keys_i[e]), c >> 1, int(keys_j[e]), c & 1)] = v
    return d


def check_against(ref, got, tot):
    assert np.array_equal(np.stack([got["key_i"], got["key_j"]], 1), ref["full_keys"])
    assert np.array_equal(got["full"].astype(np.int64), ref["full_vals"])
    sel = np.nonzero(got["flank"] > 0)[0]
    sel = sel[np.argsort(got["first_flank"][sel], kind="stable")]
    assert np.array_equal(np.stack([got["key_i"][sel], got["key_j"][sel]], 1), ref["flank_keys"])
    assert np.array_equal(got["flank"][sel].astype(np.int64), ref["flank_vals"])
    want = {tuple(k): int(v) for k, v in zip(ref["HT_keys"].tolist(), ref["HT_vals"].tolist())}
    assert ht_dict(got["key_i"], got["key_j"], got["ht"]) == want
    want_tot = np.zeros(len(tot), np.int64)
    want_tot[ref["ctg_link_ids"]] = ref["ctg_link_vals"]
    assert np.array_equal(tot, want_tot)


@pytest.fixture(params=["direct", "partitioned"])
def counting_mode(request, monkeypatch):
    """both counting engines: one big hash table, and partition-then-aggregate (forced on here; by default it takes over
    for streams of 16M records and more)"""
    monkeypatch.setenv("HH_LINKS_PARTITION", "1" if request.param == "partitioned" else "0")
    if request.param == "partitioned":
        monkeypatch.setenv("HH_LINKS_NPART_LOG", "5")
    return request.param


@pytest.mark.parametrize("tag", ["a", "b"])
def test_links_match_reference_golden(ctx, tag, counting_mode):
    from haphic_b200.links import LinkTable
    g = load_golden("links_{}.npz".format(tag))
    tab = LinkTable(ctx, g["lengths"], rank_of(g["names"].tolist()), g["in_nx"], int(g["flank_kb"]) * 1000)
    tab.add(g["pairs"])
    info = tab.finish()
    assert info.n_records == len(g["pairs"])
    assert info.nnz_full == len(g["full_vals"]) and info.nnz_flank == len(g["flank_vals"])
    check_against(g, tab.fetch(), tab.fetch_ctg())
    tab.close()


def synth_case(nchr, n_contigs, mean_len, n_pairs, seed, sort=False):
    from haphic_b200 import synth
    asm = synth.make_assembly(nchr, n_contigs, mean_len, seed=seed)
    pairs = synth.make_pairs(asm, n_pairs, seed=seed + 1).numpy()
    if sort:   # coordinate-sorted input: long runs of equal keys exercise the warp aggregation
        order = np.lexsort((pairs[:, 3], pairs[:, 1], pairs[:, 2], pairs[:, 0]))
        pairs = np.ascontiguousarray(pairs[order])
    return asm, pairs


@pytest.mark.parametrize("sort", [False, True])
@pytest.mark.parametrize("flank_kb,nx_frac", [(500, 1.0), (5, 0.7)])
def test_links_match_oracle_1m(ctx, sort, flank_kb, nx_frac, counting_mode):
    from haphic_b200.links import LinkTable
    from oracle import haphic_oracle as orc
    asm, pairs = synth_case(6, 1200, 30000, 1_000_000, seed=11, sort=sort)
    rng = np.random.default_rng(3)
    in_nx = (rng.random(asm.n) < nx_frac).astype(np.uint8)
    rank = rank_of(asm.names)
    # a few records naming contigs that are not in the FASTA, and some negative ids
    pairs[::997, 0] = asm.n + 5
    pairs[5::1013, 2] = -1
    ref = orc.count_links_numpy(pairs, asm.lengths, rank, in_nx, flank_kb * 1000)
    tab = LinkTable(ctx, asm.lengths, rank, in_nx, flank_kb * 1000)     # no capacity hint: the table grows
    # stream in uneven chunks, half from host memory, half from device memory
    import torch
    cuts = [0, 1, 33, 70001, 400000, 400032, len(pairs)]
    for k in range(len(cuts) - 1):
        chunk = pairs[cuts[k]:cuts[k + 1]]
        if k % 2:
            chunk = torch.from_numpy(np.ascontiguousarray(chunk)).cuda()
        tab.add(chunk)
    info = tab.finish()
    assert info.n_used == ref["n_used"]
    check_against(ref, tab.fetch(), tab.fetch_ctg())
    tab.close()


def test_links_edge_cases(ctx):
    from haphic_b200.links import LinkTable
    lengths = np.array([1000, 2000, 3000, 10], np.int64)
    rank = np.array([2, 0, 1, 3], np.int32)       # name order differs from id order
    nx = np.array([1, 1, 0, 1], np.uint8)
    # empty stream
    tab = LinkTable(ctx, lengths, rank, nx, 100)
    info = tab.finish()
    assert info.nnz_full == 0 and info.n_records == 0
    f = tab.fetch()
    assert len(f["key_i"]) == 0
    tab.close()
    # only unusable records: intra-contig and unknown ids
    tab = LinkTable(ctx, lengths, rank, nx, 100)
    tab.add(np.array([[0, 5, 0, 9], [7, 1, 1, 1], [1, 1, -3, 1]], np.int32))
    info = tab.finish()
    assert info.nnz_full == 0 and info.n_used == 0 and info.n_records == 3
    tab.close()
    # hand-checked semantics: name order, flank borders (1-based), head/tail split, Nx mask
    rec = np.array([
        [0, 99, 1, 0],      # coords 100 / 1 : both inside flank=100 -> flank link; key ordered by rank: (1, 0)
        [0, 100, 1, 0],     # coord 101 on ctg0 (len 1000): outside both flanks -> full only
        [0, 900, 1, 1999],  # coord 901 > 1000-100 -> flank; coord 2000 > 1900 -> flank; both tails
        [1, 999, 0, 499],   # coord_1 = 1000 (2*1000 > 2000 false -> H), coord_0 = 500 (2*500 > 1000 false -> H)
        [1, 1000, 0, 500],  # coord_1 = 1001 -> T ; coord_0 = 501 -> T
        [2, 0, 0, 0],       # ctg2 not in Nx -> never a flank link; key (2, 0) since rank[2]=1 < rank[0]=2
        [3, 4, 1, 50],      # ctg3 len 10 <= 2*flank: always flank; key (1, 3)
    ], np.int32)
    tab = LinkTable(ctx, lengths, rank, nx, 100)
    tab.add(rec)
    tab.finish()
    f = tab.fetch()
    keys = list(zip(f["key_i"].tolist(), f["key_j"].tolist()))
    assert keys == [(1, 0), (2, 0), (1, 3)]
    assert f["full"].tolist() == [5, 1, 1]
    assert f["flank"].tolist() == [2, 0, 1]
    assert f["first_full"].tolist() == [0, 5, 6]
    assert f["first_flank"].tolist() == [0, 0xFFFFFFFF, 6]
    # HT of pair (1,0): records 0:(H,H) 1:(H,H) 2:(T,T) 3:(H,H) 4:(T,T)   [i = ctg1, j = ctg0]
    assert f["ht"][0].tolist() == [3, 0, 0, 2]
    assert tab.fetch_ctg().tolist() == [2, 3, 0, 1]
    tab.close()


def test_links_sharded_merge_equals_single(ctx):
    """Two shards of the stream counted separately and merged give the single-stream table."""
    from haphic_b200.links import LinkTable
    asm, pairs = synth_case(4, 400, 30000, 300_000, seed=21)
    rank = rank_of(asm.names)
    nx = np.ones(asm.n, np.uint8)
    one = LinkTable(ctx, asm.lengths, rank, nx, 500000)
    one.add(pairs)
    one.finish()
    want, want_tot = one.fetch(), one.fetch_ctg()
    half = len(pairs) // 2
    a = LinkTable(ctx, asm.lengths, rank, nx, 500000)
    b = LinkTable(ctx, asm.lengths, rank, nx, 500000)
    b.add(pairs[half:], stream_offset=half)
    b.finish()
    ent, tot, nrec, nused = b.export()
    a.add(pairs[:half], stream_offset=0)
    a.merge(ent, tot, nrec, nused)
    info = a.finish()
    got = a.fetch()
    assert info.n_records == len(pairs)
    for k in want:
        assert np.array_equal(want[k], got[k]), k
    assert np.array_equal(want_tot, a.fetch_ctg())
    for t in (one, a, b):
        t.close()


@pytest.mark.parametrize("world", [2, 3])
def test_links_routed_equals_single(ctx, world):
    """The routed protocol of dist.routed_link_build emulated on one GPU: shards of the stream are routed to the
    owner of their contig pair, counted into disjoint partitions, and the adopted union equals the single table
    (dict order restored by the fetch), index and matrix included."""
    import torch
    from haphic_b200.links import LinkTable
    asm, pairs = synth_case(4, 400, 30000, 300_000, seed=23)
    pairs[::991, 0] = asm.n + 3                   # ids outside the FASTA are dropped by the router
    rank = rank_of(asm.names)
    nx = (np.random.default_rng(1).random(asm.n) < 0.8).astype(np.uint8)
    keep = np.ones(asm.n, np.uint8)
    one = LinkTable(ctx, asm.lengths, rank, nx, 20000)
    one.add(pairs)
    info1 = one.finish()
    want, want_tot = one.fetch(), one.fetch_ctg()
    want_index, want_nl = one.linked_index(keep)
    m1 = one.to_matrix(keep, np.nonzero(want_index < 0)[0].astype(np.int32)).to_scipy()
    cuts = np.linspace(0, len(pairs), world + 1).astype(int)
    cuts[1] += 7                                   # uneven shards
    tabs = [LinkTable(ctx, asm.lengths, rank, nx, 20000) for _ in range(world)]
    routed = []
    for r in range(world):
        shard = torch.from_numpy(np.ascontiguousarray(pairs[cuts[r]:cuts[r + 1]])).cuda()
        rec_out, pos_out, counts = tabs[r].route(shard, int(cuts[r]), world)
        assert sum(counts) <= len(shard)
        routed.append((rec_out, pos_out, np.concatenate([[0], np.cumsum(counts)])))
    parts, tots, n_used = [], [], 0
    for d in range(world):                          # "all-to-all": destination d takes its group from every source
        for r in range(world):
            rec_out, pos_out, off = routed[r]
            tabs[d].add_routed(rec_out[off[d]:off[d + 1]].contiguous(), pos_out[off[d]:off[d + 1]].contiguous())
        part = tabs[d].finish_partition()
        n_used += int(part.n_used)
        ent, tot, _, _ = tabs[d].export()
        parts.append(ent)
        tots.append(tot)
    keys = [set(map(tuple, p[:, :2].cpu().numpy().tolist())) for p in parts]
    for a in range(world):
        for b in range(a + 1, world):
            assert not (keys[a] & keys[b])          # disjoint partitions
    whole = torch.cat(parts)
    tot = torch.stack(tots).sum(0)
    for d in (0, world - 1):
        info = tabs[d].adopt(whole, tot, len(pairs), n_used, len(pairs))
        assert (info.n_records, info.n_used, info.nnz_full, info.nnz_flank) == \
               (info1.n_records, info1.n_used, info1.nnz_full, info1.nnz_flank)
        index, nl = tabs[d].linked_index(keep)     # works on the unordered list
        assert nl == want_nl and np.array_equal(index, want_index)
        m = tabs[d].to_matrix(keep, np.nonzero(index < 0)[0].astype(np.int32)).to_scipy()
        assert (m != m1).nnz == 0
        got = tabs[d].fetch()                       # restores dict insertion order
        for k in want:
            assert np.array_equal(want[k], got[k]), k
        assert np.array_equal(want_tot, tabs[d].fetch_ctg())
    for t in tabs + [one]:
        t.close()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_matrix_from_links_matches_reference_golden(ctx, tag):
    """dict_to_matrix: first-seen indices and the symmetric CSC with self loops (and, for case b,
    normalize_by_nlinks values within 1e-6 relative)."""
    from haphic_b200.links import LinkTable
    from tests.util import csc_from
    g = load_golden("links_{}.npz".format(tag))
    tab = LinkTable(ctx, g["lengths"], rank_of(g["names"].tolist()), g["in_nx"], int(g["flank_kb"]) * 1000)
    tab.add(g["pairs"])
    tab.finish()
    keep = g["filtered"]
    index, n_linked = tab.linked_index(keep)
    idx_ref = g["matrix_index"]
    linked = index >= 0
    assert np.array_equal(index[linked], idx_ref[linked])
    assert n_linked == int(linked.sum())
    tail_ids = np.nonzero((idx_ref >= 0) & ~linked)[0]
    tail = tail_ids[np.argsort(idx_ref[tail_ids])]
    assert np.array_equal(idx_ref[tail], n_linked + np.arange(len(tail)))
    normalize = "flank_norm_vals" in g.files
    mat = tab.to_matrix(keep, tail, normalize_by_nlinks=normalize)
    got = mat.to_scipy()
    ref = csc_from(g, "link", got.shape[0])
    assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)
    if normalize:
        assert np.allclose(got.data, ref.data, rtol=1e-6, atol=0)
    else:
        assert np.array_equal(got.data, ref.data)
    mat.close()
    tab.close()


def test_fragment_mode_matches_reference_golden(ctx):
    """parse_alignments (contigs split into bins): flank links and per-fragment totals are keyed by fragments
    (second table in fragment mode); full / HT stay contig-level."""
    from haphic_b200.links import LinkTable, name_rank
    g = load_golden("links_bins.npz")
    names = g["names"].tolist()
    frag_names = g["frag_names"].tolist()
    flank_bp = int(g["flank_kb"]) * 1000
    pairs = g["pairs"]
    ftab = LinkTable(ctx, g["frag_len"], name_rank(frag_names), g["frag_in_nx"], flank_bp,
                     frags=dict(ctg_rank=name_rank(names), frag_base=g["frag_base"], bin_size=int(g["bin_size"])))
    ctab = LinkTable(ctx, g["lengths"], name_rank(names), np.zeros(len(names), np.uint8), flank_bp)
    for lo in range(0, len(pairs), 17000):
        ftab.add(pairs[lo:lo + 17000])
        ctab.add(pairs[lo:lo + 17000])
    ftab.finish()
    ctab.finish()
    f = ftab.fetch()
    sel = np.nonzero(f["flank"] > 0)[0]
    sel = sel[np.argsort(f["first_flank"][sel], kind="stable")]
    assert np.array_equal(np.stack([f["key_i"][sel], f["key_j"][sel]], 1), g["flank_keys"])
    assert np.array_equal(f["flank"][sel].astype(np.int64), g["flank_vals"])
    tot = ftab.fetch_ctg()
    want = np.zeros(len(frag_names), np.int64)
    want[g["frag_link_ids"]] = g["frag_link_vals"]
    assert np.array_equal(tot, want)
    c = ctab.fetch()
    assert np.array_equal(np.stack([c["key_i"], c["key_j"]], 1), g["full_keys"])
    assert np.array_equal(c["full"].astype(np.int64), g["full_vals"])
    assert int(c["flank"].sum()) == 0
    got_ht = ht_dict(c["key_i"], c["key_j"], c["ht"])
    assert got_ht == {tuple(k): int(v) for k, v in zip(g["HT_keys"].tolist(), g["HT_vals"].tolist())}
    ftab.close()
    ctab.close()


@pytest.mark.parametrize("bad_pos", [-1, None])
def test_fragment_mode_refuses_positions_outside_the_contig(ctx, bad_pos):
    """A position outside its (split) contig -- .pairs position 0 -> -1, or beyond the last bin -- names a bin that does
    not exist.  The reference fails with a KeyError (frag_len_dict['ctg_binK']); the table must fail loudly too instead
    of crediting a neighbouring contig's fragment."""
    from haphic_b200._lib import HHError
    from haphic_b200.links import LinkTable, name_rank
    g = load_golden("links_bins.npz")
    names = g["names"].tolist()
    frag_names = g["frag_names"].tolist()
    base = g["frag_base"]
    split = int(np.nonzero(np.diff(base) > 1)[0][-1])          # the last split contig
    other = 0 if split != 0 else 1
    pos = bad_pos if bad_pos is not None else int(g["lengths"][split]) + 5 * int(g["bin_size"])
    rec = np.array([[split, pos, other, 10]], dtype=np.int32)
    ftab = LinkTable(ctx, g["frag_len"], name_rank(frag_names), g["frag_in_nx"], int(g["flank_kb"]) * 1000,
                     frags=dict(ctg_rank=name_rank(names), frag_base=base, bin_size=int(g["bin_size"])))
    ftab.add(g["pairs"][:1000])
    ftab.add(rec)
    with pytest.raises(HHError) as e:
        ftab.finish()
    assert "outside" in str(e.value)
    ftab.close()


@pytest.mark.parametrize("tag,topN", [("a", 10), ("b", 10), ("a", 4)])
def test_rank_sums_match_reference_filter_fragments(ctx, tag, topN):
    """The rank-sum statistic of filter_fragments (864-892) against the numbers the UNMODIFIED reference computed for the
    same record stream (tests/golden/ranksum_*.npz, read out of filter_fragments' own frame by make_ranksum_golden.py):
    device table -> matrix with the reference's fragment index -> hh_matrix_rank_sums."""
    from haphic_b200.links import LinkTable
    g = load_golden("links_{}.npz".format(tag))
    rs = load_golden("ranksum_{}_top{}.npz".format(tag, topN))
    tab = LinkTable(ctx, g["lengths"], rank_of(g["names"].tolist()), g["in_nx"], int(g["flank_kb"]) * 1000)
    tab.add(g["pairs"])
    tab.finish()
    n = len(g["lengths"])
    keep = np.zeros(n, np.uint8)
    keep[rs["frag_ids"]] = 1
    index, n_linked = tab.linked_index(keep)
    ref_index = np.full(n, -1, np.int64)
    ref_index[rs["frag_ids"]] = rs["frag_index"]
    linked = index >= 0
    assert np.array_equal(index[linked], ref_index[linked])                   # first-seen indices (327-349)
    tail = rs["frag_ids"][np.argsort(rs["frag_index"], kind="stable")]        # unlinked fragments in the reference's order (355-359)
    tail = tail[~linked[tail]].astype(np.int32)
    mat = tab.to_matrix(keep, tail, add_self_loops=False)
    got = mat.rank_sums(int(rs["topN"]))
    want = np.full(mat.n, -1, np.int64)
    want[ref_index[rs["rank_ids"]]] = rs["rank_sums"]
    sel = want >= 0
    assert sel.sum() == len(rs["rank_ids"]) and np.array_equal(got[sel], want[sel])
    mat.close()
    tab.close()
