#!/usr/bin/env python3
"""Rank sums of filter_fragments (HapHiC_cluster.py:864-892) as the UNMODIFIED reference computes them.

The reference does not return them: they are read out of filter_fragments' own frame at the moment it calls
numpy.quantile on them (a spy installed on `HapHiC_cluster.quantile`), together with the fragment set and the matrix
index it used.  Inputs = the record stream of tests/golden/links_{a,b}.npz.  Run in the build container:

    python tests/golden/make_ranksum_golden.py
"""
import os
import sys

if os.environ.get("PYTHONHASHSEED") != "0":
    os.environ["PYTHONHASHSEED"] = "0"
    os.execv(sys.executable, [sys.executable] + sys.argv)

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import numpy as np
from make_golden import import_reference, make_args


def main():
    ref = import_reference()
    ref.logger.setLevel(100)
    for tag, topN in (("a", 10), ("b", 10), ("a", 4)):
        g = np.load(os.path.join(HERE, "links_{}.npz".format(tag)))
        names = g["names"].tolist()
        flank_d = {(names[i], names[j]): int(v) for (i, j), v in zip(g["flank_keys"].tolist(), g["flank_vals"].tolist())}
        ctg_links = {names[i]: int(v) for i, v in zip(g["ctg_link_ids"].tolist(), g["ctg_link_vals"].tolist())}
        nx = {n for n, f in zip(names, g["in_nx"].tolist()) if f}
        re_sites = {n: 1000 for n in names}                       # far above --RE_site_cutoff: no fragment is removed by it
        captured = {}
        real_quantile = ref.quantile

        def spy(values, qs):
            loc = sys._getframe(1).f_locals
            if "rank_sum_list" in loc:
                captured["rank_sum_list"] = list(loc["rank_sum_list"])
                captured["frag_index_dict"] = dict(loc["frag_index_dict"])
            return real_quantile(values, qs)

        ref.quantile = spy
        try:
            a = make_args(topN=topN)
            ref.filter_fragments(nx, re_sites, a.RE_site_cutoff, ctg_links, a.density_lower, a.density_upper, a.topN, a.rank_sum_upper,
                                 a.rank_sum_hard_cutoff, flank_d, dict(), a.read_depth_upper, set())
        finally:
            ref.quantile = real_quantile
        rs = captured["rank_sum_list"]
        idx = captured["frag_index_dict"]
        name_to_id = {n: i for i, n in enumerate(names)}
        out = {
            "topN": np.int64(topN),
            "rank_ids": np.array([name_to_id[f] for f, _ in rs], np.int32),
            "rank_sums": np.array([v for _, v in rs], np.int64),
            # the fragments that entered the stage and the matrix index the reference gave each of them
            "frag_ids": np.array([name_to_id[f] for f in idx], np.int32),
            "frag_index": np.array([idx[f] for f in idx], np.int32),
        }
        fn = os.path.join(HERE, "ranksum_{}_top{}.npz".format(tag, topN))
        np.savez_compressed(fn, **out)
        print(fn, len(rs), "fragments, rank sums", out["rank_sums"].min(), "..", out["rank_sums"].max())


if __name__ == "__main__":
    main()
