#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:hh_k_col -s 1 -c 3 -f -o gpurun_out/prof_col2 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 0 > gpurun_out/prof_col2.log 2>&1
tail -2 gpurun_out/prof_col2.log | cut -c1-300
