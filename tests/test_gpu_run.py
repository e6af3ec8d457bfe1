"""GPU parity of the whole drop-in `haphic cluster` run (haphic_b200.cluster.run) against the
output files the unmodified reference wrote for the same inputs (tests/golden/run_*.npz)."""

import hashlib
import json
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from tests.util import load_golden

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r"""
import json, os, sys
sys.path.insert(0, {repo!r})
from haphic_b200 import cluster, synth, hicio
asm = synth.make_assembly({nchr}, {n_contigs}, {mean_len}, seed={seed})
pairs = synth.make_pairs(asm, {n_pairs}, seed={seed} + 1, homolog={homolog!r}).numpy()
synth.write_fasta(asm, "asm.fa", seed={seed} + 3)
if {bam!r}:
    hicio.write_bam("aln.bam", asm.names, asm.lengths.tolist(), pairs)
    aln = "aln.bam"
else:
    synth.write_pairs(asm, pairs, "aln.pairs")
    aln = "aln.pairs"
argv = ["asm.fa", aln, str({nchr})] + {extra!r}
args = cluster.parse_arguments(argv)
cluster.run(args, log_file="HapHiC_cluster.log")
"""


def run_case(tmp_path, g, bam, env_extra=None):
    nchr, n_contigs, mean_len, n_pairs = g["shape"].tolist()
    kw = json.loads(str(g["argkw"]))
    extra = []
    for k, v in kw.items():
        extra += ["--" + k, str(v)]
    homolog = json.loads(str(g["homolog"])) if "homolog" in g else None
    code = DRIVER.format(repo=REPO, nchr=nchr, n_contigs=n_contigs, mean_len=mean_len, n_pairs=n_pairs, seed=int(g["seed"]),
                         bam=bam, extra=extra, homolog=tuple(homolog) if homolog else None)
    env = dict(os.environ, PYTHONHASHSEED="0")      # the reference's set-iteration orders (fixtures used seed 0)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return r


@pytest.mark.parametrize("tag,bam", [("c1", False), ("c1_nx80", False), ("c1_nx80", True), ("bins", False), ("bins", True),
                                     ("allelic4", False)])
def test_cluster_run_matches_reference_files(tmp_path, tag, bam):
    g = load_golden("run_{}.npz".format(tag))
    run_case(tmp_path, g, bam)
    want = json.loads(str(g["files_json"]))
    got = {}
    for root, _d, files in os.walk(tmp_path):
        for fn in files:
            p = os.path.relpath(os.path.join(root, fn), tmp_path)
            if p.startswith("inflation_") and p.endswith(".txt"):
                with open(os.path.join(root, fn)) as f:
                    got[p] = f.read()
    assert sorted(got) == sorted(want)
    for p in sorted(want):
        assert got[p] == want[p], p
    with open(tmp_path / "HapHiC_cluster.log") as f:
        log = f.read()
    rec = [ln.split("] ", 1)[1] for ln in log.splitlines() if "[recommend_inflation]" in ln]
    assert rec == g["recommend_lines"].tolist()
    conv = [ln.split("] ", 1)[1] for ln in log.splitlines() if "[mcl]" in ln]
    assert conv == g["mcl_lines"].tolist()
    with open(tmp_path / "full_links.pkl", "rb") as f:
        full = pickle.load(f)
    assert type(full).__name__ == "defaultdict"
    assert sorted([a, b, int(v)] for (a, b), v in full.items()) == json.loads(str(g["full_links_sorted"]))
    assert [list(k) for k in full.keys()][:50] == [r[:2] for r in _insertion(g, full)][:50]
    with open(tmp_path / "HT_links.pkl", "rb") as f:
        ht = pickle.load(f)
    assert sorted([a, b, int(v)] for (a, b), v in ht.items()) == json.loads(str(g["HT_links_sorted"]))
    with open(tmp_path / "paired_links.clm") as f:
        assert f.read() == str(g["clm_text"])
    if not bam:
        with open(tmp_path / "alignments.bed") as f:
            bed = f.read()
        assert bed[:2000] == str(g["bed_head"])
        assert hashlib.sha1(bed.encode()).hexdigest() == str(g["bed_sha1"])


def _insertion(g, full):
    # the sorted golden list does not carry insertion order; the clm text does (pairs with >= 2 links
    # appear in full_link_dict order): use it to check the head of the order
    seen, out = set(), []
    for ln in str(g["clm_text"]).splitlines():
        a, b = ln.split("\t")[0].split(" ")
        k = (a[:-1], b[:-1])
        if k not in seen:
            seen.add(k)
            out.append([k[0], k[1]])
    keys = [list(k) for k in full.keys() if tuple(k) in seen]
    assert keys == [k for k in out if tuple(k) in full]        # (allelic link removal deletes pairs after the clm is written)
    return [list(k) for k in full.keys()]


def test_cluster_run_on_several_gpus_writes_the_same_files(tmp_path):
    """HAPHIC_GPUS=N spreads the inflation sweep of run_mcl_clustering over N devices of one process; every output file and
    the machine-read log lines must equal the reference's (= the single-GPU run's)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    g = load_golden("run_c1.npz")
    run_case(tmp_path, g, False, env_extra={"HAPHIC_GPUS": str(min(4, torch.cuda.device_count()))})
    want = json.loads(str(g["files_json"]))
    got = {}
    for root, _d, files in os.walk(tmp_path):
        for fn in files:
            p = os.path.relpath(os.path.join(root, fn), tmp_path)
            if p.startswith("inflation_") and p.endswith(".txt"):
                with open(os.path.join(root, fn)) as f:
                    got[p] = f.read()
    assert sorted(got) == sorted(want)
    for p in sorted(want):
        assert got[p] == want[p], p
    with open(tmp_path / "HapHiC_cluster.log") as f:
        log = f.read()
    assert [ln.split("] ", 1)[1] for ln in log.splitlines() if "[recommend_inflation]" in ln] == g["recommend_lines"].tolist()
    assert [ln.split("] ", 1)[1] for ln in log.splitlines() if "[mcl]" in ln] == g["mcl_lines"].tolist()


def test_c4_shaped_run_matches_reference_digests(tmp_path):
    """BASELINE configs[3] in shape (4 haplotypes, `--remove_allelic_links 4`) at 10 000 contigs / 5M pairs: the unmodified
    reference needed 378 s for this run (tests/golden/make_golden.py run_c4); its output files are kept as SHA-1 digests.
    Every file, dict and machine-read log line must match; iteration counts may differ by one round (a different but equally
    valid fp32 summation order inside the expansion moves the fp32 convergence test), the clusters may not."""
    import re
    g = load_golden("rundigest_c4_10k.npz")
    run_case(tmp_path, g, False)

    def sha(text):
        return hashlib.sha1(text.encode()).hexdigest()

    want = json.loads(str(g["files_json"]))
    got = {}
    for root, _d, files in os.walk(tmp_path):
        for fn in files:
            p = os.path.relpath(os.path.join(root, fn), tmp_path)
            if p.startswith("inflation_") and p.endswith(".txt"):
                with open(os.path.join(root, fn)) as f:
                    got[p] = sha(f.read())
    assert sorted(got) == sorted(want)
    bad = [p for p in sorted(want) if got[p] != want[p]]
    assert not bad, bad[:10]
    with open(tmp_path / "full_links.pkl", "rb") as f:
        full = pickle.load(f)
    assert len(full) == int(g["n_full"])
    assert sha(json.dumps(sorted([[a, b, int(v)] for (a, b), v in full.items()]))) == str(g["full_links_sha1"])
    assert sha(json.dumps([[a, b] for (a, b) in full.keys()])) == str(g["full_links_order_sha1"])      # dict insertion order
    with open(tmp_path / "HT_links.pkl", "rb") as f:
        ht = pickle.load(f)
    assert len(ht) == int(g["n_HT"])
    assert sha(json.dumps(sorted([[a, b, int(v)] for (a, b), v in ht.items()]))) == str(g["HT_links_sha1"])
    with open(tmp_path / "paired_links.clm") as f:
        assert sha(f.read()) == str(g["clm_sha1"])
    with open(tmp_path / "alignments.bed") as f:
        assert sha(f.read()) == str(g["bed_sha1"])
    with open(tmp_path / "HapHiC_cluster.log") as f:
        log = f.read()
    assert [ln.split("] ", 1)[1] for ln in log.splitlines() if "[recommend_inflation]" in ln] == g["recommend_lines"].tolist()
    rounds = [int(re.search(r"after (\d+) rounds", ln).group(1)) for ln in log.splitlines() if "[mcl]" in ln]
    want_rounds = [int(re.search(r"after (\d+) rounds", ln).group(1)) for ln in g["mcl_lines"].tolist()]
    assert len(rounds) == len(want_rounds) and all(abs(a - b) <= 1 for a, b in zip(rounds, want_rounds)), (rounds, want_rounds)
