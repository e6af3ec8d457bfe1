#!/bin/bash
# ncu evidence for profiles/: per-launch device times of one bench step (our kernels only), and full
# captures of the dominant kernels.  Numbers printed under ncu are never bench values.
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 0"
echo "== launch list"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hh_ -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py $ARGS > gpurun_out/launches_bench.log 2>&1
tail -1 gpurun_out/launches_bench.log | cut -c1-200
echo "== full capture: column kernels -- pre-expansion, dense iteration 0, then the first window / big-list expansions"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:hh_k_col -c 6 -f -o gpurun_out/prof_col \
    python bench.py $ARGS > gpurun_out/prof_col.log 2>&1
tail -1 gpurun_out/prof_col.log | cut -c1-200
echo "== full capture: link insert kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hh_k_links_insert -c 1 -f -o gpurun_out/prof_links \
    python bench.py $ARGS > gpurun_out/prof_links.log 2>&1
tail -1 gpurun_out/prof_links.log | cut -c1-200
ls -la gpurun_out/ | grep ncu-rep
