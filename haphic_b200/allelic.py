"""Allele-aware link removal of `haphic cluster --remove_allelic_links / --remove_concentrated_links`
(scripts/HapHiC_cluster.py:419-471 record_coord_pairs / the two ratios, 474-692 remove_allelic_HiC_links).

Host-side graph logic (networkx cliques, Hungarian matching) exactly as far as results go; the per-read-pair part --
the first ``max_read_pairs`` coordinate pairs of every contig pair in stream order -- is taken from the same usable
record array the CLM writer uses (vectorised numpy: a stable sort by contig pair keeps stream order inside a pair)."""

from __future__ import annotations

import logging
from array import array
from collections import defaultdict
from math import inf

import numpy as np

logger = logging.getLogger("haphic_b200.cluster")


# ------------------------------------------------------------------------------------------------
# the two per-pair statistics (419-451)
# ------------------------------------------------------------------------------------------------

def cal_concordance_ratio(coord_list, shorter_len, nwindows):
    """Share of a contig pair's links on its most populated diagonal (y - x) or anti-diagonal (y + x), in windows of
    shorter_len // nwindows bp (419-428).  ``coord_list`` = [x0, y0, x1, y1, ...] (1-based coordinates)."""
    width = shorter_len // nwindows
    if width == 0:
        raise ZeroDivisionError("integer division or modulo by zero")      # what the reference's `//` raises
    c = np.asarray(coord_list, dtype=np.int64)
    x, y = c[0::2], c[1::2]
    npairs = len(x)
    best = 0
    for v in ((y - x) // width, (y + x) // width):
        best = max(best, int(np.unique(v, return_counts=True)[1].max()))
    return best / npairs


def cal_concentration_adj_ratio(coord_list, bin_width=10000):
    """(1 - share of links in over-populated x bins) * (1 - same for y) (431-451): a bin is over-populated when it
    holds >= 10x the median count of the occupied bins."""
    c = np.asarray(coord_list, dtype=np.int64)
    npairs = len(c) // 2
    out = 1.0
    for v in (c[0:2 * npairs:2], c[1:2 * npairs:2]):
        counts = np.unique(v // bin_width, return_counts=True)[1]
        med = np.median(counts)
        out_share = int(counts[counts >= 10 * med].sum()) / npairs
        out = out * (1 - out_share)
    return out


# ------------------------------------------------------------------------------------------------
# record_coord_pairs (454-471) over the whole record array
# ------------------------------------------------------------------------------------------------

def coord_pair_dict(rec, names, ctg_len, rank, args, pos_int_type="int32", skip_below=0):
    """ctg_coord_dict as parse_alignments(_for_ctgs) returns it (1609-1611, 1652-1653): for every contig pair in
    first-seen order either the ``array`` of its (coord_i, coord_j) values in stream order (fewer than
    ``args.max_read_pairs`` links) or ``[concordance_ratio, adj_ratio]`` computed from the first max_read_pairs links.

    ``rec`` holds the usable inter-contig records (int32 [m, 4], stream order, ids into ``names``).
    ``skip_below`` > 0 leaves out the pairs with fewer links than that: remove_allelic_HiC_links only writes a debug
    line for pairs under ``min_read_pairs``, so run() skips them unless --verbose (10^7 tiny Python arrays otherwise)."""
    code = "i" if pos_int_type == "int32" else "l"
    out = defaultdict(lambda: array(code))
    if len(rec) == 0:
        return out
    n_names = len(names)
    r = rec.astype(np.int64)
    swap = rank[r[:, 0]] > rank[r[:, 2]]
    i = np.where(swap, r[:, 2], r[:, 0])
    j = np.where(swap, r[:, 0], r[:, 2])
    ci = np.where(swap, r[:, 3], r[:, 1]) + 1          # 1-based, oriented like the (name-sorted) pair
    cj = np.where(swap, r[:, 1], r[:, 3]) + 1
    key = i * n_names + j
    order = np.argsort(key, kind="stable")
    ks = key[order]
    starts = np.concatenate([[0], np.nonzero(np.diff(ks))[0] + 1])
    ends = np.concatenate([starts[1:], [len(ks)]])
    first_seen = np.argsort(order[starts], kind="stable")
    if skip_below > 0:
        first_seen = first_seen[(ends - starts)[first_seen] >= skip_below]
    limit = int(args.max_read_pairs)
    xy = np.empty((len(ks), 2), dtype=np.int64)
    xy[:, 0] = ci[order]
    xy[:, 1] = cj[order]
    ctg_len = np.asarray(ctg_len)
    for s in first_seen.tolist():
        b, e = int(starts[s]), int(ends[s])
        k = int(ks[b])
        a_id, b_id = k // n_names, k % n_names
        pair = (names[a_id], names[b_id])
        if e - b < limit:
            out[pair] = array(code, xy[b:e].reshape(-1).tolist())
            continue
        coords = xy[b:b + limit].reshape(-1)
        data = None
        if args.remove_allelic_links:
            shorter = int(min(ctg_len[a_id], ctg_len[b_id]))
            data = [cal_concordance_ratio(coords, shorter, args.nwindows), 1]
        if args.remove_concentrated_links:
            if data is not None:
                # with both flags the reference evaluates the adjustment on the dict entry it has just replaced by
                # [concordance_ratio, 1] (465-469), i.e. on ONE pseudo coordinate pair -- which yields 1.0; kept as is
                data[1] = cal_concentration_adj_ratio(data)
            else:
                data = [0, cal_concentration_adj_ratio(coords)]
        out[pair] = data
    return out


def ctg_pair_to_frag_dict(rec, names, rank, frag_names, frag_base, frag_rank, bin_size):
    """ctg_pair_to_frag of parse_alignments (1731-1733) for the inter-contig records: contig pair -> set of the
    fragment (bin) pairs its links fall into."""
    out = defaultdict(set)
    if len(rec) == 0:
        return out
    r = rec.astype(np.int64)
    swap = rank[r[:, 0]] > rank[r[:, 2]]
    i = np.where(swap, r[:, 2], r[:, 0])
    j = np.where(swap, r[:, 0], r[:, 2])
    ci = np.where(swap, r[:, 3], r[:, 1]) + 1
    cj = np.where(swap, r[:, 1], r[:, 3]) + 1
    frag_base = np.asarray(frag_base, dtype=np.int64)
    frag_rank = np.asarray(frag_rank)

    def to_frag(c, coord):
        nb = frag_base[c + 1] - frag_base[c]
        split = nb > 1
        return frag_base[c] + np.where(split, -(-coord // bin_size) - 1, 0), split      # ceil(coord / bin_size) - 1

    fi, si = to_frag(i, ci)
    fj, sj = to_frag(j, cj)
    flip = (si | sj) & (frag_rank[fi] > frag_rank[fj])             # sort by bin name when a bin is involved (1719-1720)
    fi, fj = np.where(flip, fj, fi), np.where(flip, fi, fj)
    combos = np.unique(np.stack([i, j, fi, fj], 1), axis=0)
    for a, b, fa, fb in combos.tolist():
        out[(names[a], names[b])].add((frag_names[fa], frag_names[fb]))
    return out


# ------------------------------------------------------------------------------------------------
# remove_allelic_HiC_links (474-692)
# ------------------------------------------------------------------------------------------------

def _weakest_edge(graph):
    """First edge of minimal weight in the graph's edge order, self loops ignored (507-519)."""
    best = (None, None, inf)
    for u, v, w in graph.edges(data="weight"):
        if u != v and w < best[2]:
            best = (u, v, w)
    assert best[0] is not None
    return best[0], best[1]


def _split_oversized(graph, cliques, ploidy, visited, out):
    """Cliques larger than the ploidy lose their weakest edge and are replaced by the maximal cliques of what is
    left, recursively (521-548); ``visited`` avoids splitting the same clique twice."""
    from networkx import Graph, find_cliques
    for clique in cliques:
        clique = tuple(clique)
        if len(clique) <= ploidy:
            out.add(clique)
            continue
        if clique in visited:
            continue
        visited.add(clique)
        view = graph.subgraph(clique)
        u, v = _weakest_edge(view)
        sub = Graph(view)                   # subgraph views are frozen
        sub.remove_edge(u, v)
        _split_oversized(sub, find_cliques(sub), ploidy, visited, out)


def allele_groups(inter_allele_dict, allelic_ctg_set, ploidy, dict_to_matrix):
    """Groups of mutually allelic contigs (599-616): the allelic pairs themselves for ploidy 2, else the maximal
    cliques of the allelic-link graph cut down to at most ``ploidy`` members."""
    if ploidy <= 2:
        return set(inter_allele_dict.keys())
    from networkx import Graph, find_cliques
    matrix, index = dict_to_matrix(inter_allele_dict, allelic_ctg_set)
    name_of = {i: c for c, i in index.items()}
    graph = Graph(matrix)
    cliques = set()
    _split_oversized(graph, find_cliques(graph), ploidy, set(), cliques)
    return {tuple(sorted(name_of[i] for i in clique)) for clique in cliques}


def remove_allelic_HiC_links(fa_dict, ctg_coord_dict, full_link_dict, args, flank_link_dict=None, filtered_frags=None,
                             ctg_pair_to_frag=None, logger=logger, dict_to_matrix=None):
    """Deletes from ``full_link_dict`` (and from ``flank_link_dict`` restricted to ``filtered_frags``) 1) the links
    between allelic contigs (concordance ratio above the cutoff) and 2) the links between contigs of two allele
    groups that are not each other's best match (maximum-weight bipartite matching).  Returns the fragments that
    still have a flank link, or None without a flank_link_dict -- the reference's contract (474-692)."""
    from scipy.optimize import linear_sum_assignment
    if dict_to_matrix is None:
        from .cluster import dict_to_matrix
    logger.info("Removing Hi-C links between alleic contig pairs...")
    ploidy = args.remove_allelic_links
    cutoff = args.concordance_ratio_cutoff
    inter_allele_dict, allelic_ctg_set = dict(), set()

    def drop(pair):
        del full_link_dict[pair]
        if not flank_link_dict:
            return
        if ctg_pair_to_frag:
            for frag_pair in ctg_pair_to_frag[pair]:
                if frag_pair in flank_link_dict and frag_pair[0] in filtered_frags and frag_pair[1] in filtered_frags:
                    del flank_link_dict[frag_pair]
        elif pair in flank_link_dict and pair[0] in filtered_frags and pair[1] in filtered_frags:
            del flank_link_dict[pair]

    # 1) allelic contig pairs by concordance ratio (578-597)
    for pair, data in ctg_coord_dict.items():
        if isinstance(data, list):
            ratio = data[0]
        elif len(data) >= args.min_read_pairs * 2:
            ratio = cal_concordance_ratio(data, min(fa_dict[pair[0]][1], fa_dict[pair[1]][1]), args.nwindows)
        else:
            ratio = 0
        logger.debug("{} {} links={} concordance_ratio={}".format(pair[0], pair[1], full_link_dict[pair], ratio))
        if ratio > cutoff:
            inter_allele_dict[pair] = full_link_dict[pair]
            allelic_ctg_set.update(pair)
            drop(pair)

    groups = allele_groups(inter_allele_dict, allelic_ctg_set, ploidy, dict_to_matrix)
    groups_of = defaultdict(set)
    for group in groups:
        for ctg in group:
            groups_of[ctg].add(group)

    # 2) links between two allele groups that the maximum matching does not select (618-675)
    solutions = dict()

    def best_partner(group_pair):
        """column chosen for every row of the (square, zero-padded) link matrix between the two groups"""
        if group_pair not in solutions:
            g1, g2 = group_pair
            size = max(len(g1), len(g2))
            links = np.zeros((size, size), dtype=int)
            for a, c1 in enumerate(g1):
                for b, c2 in enumerate(g2):
                    links[a, b] = full_link_dict.get(tuple(sorted((c1, c2))), 0)
            solutions[group_pair] = linear_sum_assignment(-links)[1]
        return solutions[group_pair]

    nonmax = set()
    for pair in full_link_dict:
        c1, c2 = pair
        if c1 not in groups_of or c2 not in groups_of:
            continue
        hit = False
        for g1 in groups_of[c1]:
            for g2 in groups_of[c2]:
                group_pair = tuple(sorted((g1, g2)))
                cols = best_partner(group_pair)
                if c1 in group_pair[0]:
                    assert c2 in group_pair[1]
                    row, col = group_pair[0].index(c1), group_pair[1].index(c2)
                else:
                    assert c2 in group_pair[0]
                    row, col = group_pair[0].index(c2), group_pair[1].index(c1)
                if cols[row] != col:
                    hit = True
                    break
            if hit:
                break
        if hit:
            nonmax.add(pair)
    for pair in nonmax:
        logger.debug("{} {} links={} non-maximum matching".format(pair[0], pair[1], full_link_dict[pair]))
        drop(pair)

    # fragments left without any flank link (677-692)
    if flank_link_dict:
        remaining = set()
        for f1, f2 in flank_link_dict:
            if f1 in filtered_frags and f2 in filtered_frags:
                remaining.add(f1)
                remaining.add(f2)
        removed = filtered_frags - remaining
        logger.info("Removing isolated fragments after filtering out allelic Hi-C links...")
        logger.info("{} fragments removed, {} fragments kept".format(len(removed), len(remaining)))
        for frag in removed:
            logger.debug("Fragment {} is isolated and removed".format(frag))
        return remaining
    return None
