#!/bin/bash
# ncu evidence for profiles/: per-launch device times of one bench step, and a full capture of the
# dominant kernels.  Numbers printed under ncu are never bench values.
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
ARGS="${BENCH_ARGS:---steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 0}"
echo "== launch list"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hh_ -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py $ARGS > gpurun_out/launches_bench.log 2>&1
tail -3 gpurun_out/launches_bench.log
echo "== full capture: MCL column kernel: pre-expansion, dense iteration 0, first sparse expansion"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:hh_k_col -s 1 -c 3 -o gpurun_out/prof_col \
    python bench.py $ARGS > gpurun_out/prof_col.log 2>&1
tail -2 gpurun_out/prof_col.log
echo "== full capture: link insert kernel, 1 launch"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hh_k_links_insert -c 1 -o gpurun_out/prof_links \
    python bench.py $ARGS > gpurun_out/prof_links.log 2>&1
tail -2 gpurun_out/prof_links.log
ls -la gpurun_out/
