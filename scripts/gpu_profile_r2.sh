#!/bin/bash
# round-2 ncu evidence for profiles/: launch list of one bench step, full captures of the tensor-core GEMM, the dense
# iteration 0, the block-diagonal iteration and the partitioned link counting.  Numbers printed under ncu are never bench values.
# The .ncu-rep files are summarised on the box (only gpurun_out/ travels back, at most 64 MiB) and deleted, except one
# single-kernel report of the GEMM.
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
S="python scripts/summarize_profiles.py r02 --outdir gpurun_out"
echo "== launch list (one warm-up step + one step)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hh_ -c 4000 --csv --log-file gpurun_out/r02_launches_raw.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 --no-default-sweep > gpurun_out/r02_launches_bench.log 2>&1
$S --launches gpurun_out/r02_launches_raw.csv
rm -f gpurun_out/r02_launches_raw.csv
echo "== full capture: MCL kernels (normalise, densify, GEMM, clip correction, dense iteration 0, block iteration)"
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:hh_k_syrk|hh_k_iter0|hh_k_slot_from_csc|hh_k_col_win|hh_k_blk|hh_k_clip|hh_k_gemm_densify" -c ${MCL_CAPTURES:-12} -f \
    -o gpurun_out/r02_prof_mcl python scripts/prof_c3.py > gpurun_out/r02_prof_mcl.log 2>&1
$S --rep gpurun_out/r02_prof_mcl.ncu-rep
rm -f gpurun_out/r02_prof_mcl.ncu-rep
if [ -z "$SKIP_LINKS" ]; then
echo "== full capture: partitioned link counting"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hh_k_part -c 4 -f -o gpurun_out/r02_prof_links \
    python scripts/prof_c3.py > gpurun_out/r02_prof_links.log 2>&1
$S --rep gpurun_out/r02_prof_links.ncu-rep
rm -f gpurun_out/r02_prof_links.ncu-rep
fi
if [ -n "$KEEP_SYRK_REP" ]; then
echo "== single-kernel report of the GEMM (kept)"
PAIRS=40000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:hh_k_syrk -c 1 -f -o gpurun_out/r02_prof_syrk \
    python scripts/prof_c3.py > /dev/null 2>&1
fi
ls -la gpurun_out/
