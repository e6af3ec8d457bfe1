"""Wall clock of a real `haphic cluster` run (haphic_b200.cluster.run) from FASTA + .pairs files to every output file, at
BASELINE.json configs[2] size by default (50k contigs / 200M pairs, `--Nx 100 --bin_size 0`, the default 20-inflation sweep).

    python scripts/run_wall.py [--contigs N --pairs P --nchr C --gpus G --keep DIR]

Prints one JSON line: seconds to fabricate the inputs (not part of the run), the run's wall clock, and the three wall-clock
figures the reference itself logs (matrix construction 2940-2941, Markov clustering 2951-2953, program total 2958-2959).
"""
import argparse
import json
import os
import re
import shutil
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--contigs", type=int, default=50000)
    ap.add_argument("--pairs", type=int, default=200_000_000)
    ap.add_argument("--nchr", type=int, default=24)
    ap.add_argument("--mean-len", type=int, default=20000)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--threads", type=int, default=16)
    ap.add_argument("--keep", default=None)
    ap.add_argument("--extra", default="")
    a = ap.parse_args()
    import ctypes as C
    import numpy as np
    from haphic_b200 import cluster, hicio, synth
    from haphic_b200._lib import check, load, ptr

    work = a.keep or tempfile.mkdtemp(prefix="haphic_run_")
    os.makedirs(work, exist_ok=True)
    t0 = time.time()
    asm = synth.make_assembly(a.nchr, a.contigs, a.mean_len, seed=12345)
    fa = os.path.join(work, "asm.fa")
    synth.write_fasta(asm, fa, seed=12348)
    pairs_path = os.path.join(work, "aln.pairs")
    blob = hicio.names_blob(asm.names)
    CH = 20_000_000
    for lo in range(0, a.pairs, CH):
        hi = min(a.pairs, lo + CH)
        rec = np.ascontiguousarray(synth.make_pairs_range(asm, lo, hi, seed=12346, device="cuda").cpu().numpy())
        check(load().hh_pairs_write(os.fsencode(pairs_path), blob, asm.n, ptr(rec), len(rec), lo, 1 if lo else 0, a.threads))
    t_inputs = time.time() - t0
    sizes = {"fasta_bytes": os.path.getsize(fa), "pairs_bytes": os.path.getsize(pairs_path)}

    out = os.path.join(work, "01.cluster")
    os.makedirs(out, exist_ok=True)
    os.chdir(out)
    if a.gpus > 1:
        os.environ["HAPHIC_GPUS"] = str(a.gpus)
    argv = [fa, pairs_path, str(a.nchr), "--Nx", "100", "--bin_size", "0", "--threads", str(a.threads)] + a.extra.split()
    args = cluster.parse_arguments(argv)
    t1 = time.time()
    cluster.run(args, log_file="HapHiC_cluster.log")
    t_run = time.time() - t1
    log = open("HapHiC_cluster.log").read()

    def grab(pat):
        m = re.search(pat, log)
        return float(m.group(1)) if m else None

    files = sorted(os.listdir("."))
    line = {
        "what": "cluster.run(): FASTA + .pairs -> HT_links.pkl, paired_links.clm, full_links.pkl, alignments.bed, inflation_*/ (all files)",
        "contigs": a.contigs, "pairs": a.pairs, "gpus": a.gpus, "host_threads": a.threads, "host_cores": os.cpu_count(),
        "argv": argv[2:], "inputs_s": round(t_inputs, 1), **sizes, "run_wall_s": round(t_run, 2),
        "log_matrix_s": grab(r"Hi-C linking matrix was constructed in ([\d.]+)s"),
        "log_mcl_s": grab(r"of Markov clustering finished in ([\d.]+)s"),
        "log_total_s": grab(r"Program finished in ([\d.]+)s"),
        "recommended": re.findall(r"You could try inflation from ([\d.]+)", log)[:1],
        "n_inflation_dirs": len([f for f in files if f.startswith("inflation_")]),
        "output_bytes": {f: os.path.getsize(f) for f in files if os.path.isfile(f)},
    }
    print("RUNWALL " + json.dumps(line), flush=True)
    os.chdir(REPO)
    if not a.keep:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
