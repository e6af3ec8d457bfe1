#!/bin/bash
# quick iteration: parity tests + one C3 bench line (+ optional launch list)
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 2>&1 | tail -25 | tee gpurun_out/tests.log
echo "== bench C3"
timeout 1200 python bench.py --steps 1 --warmup 1 ${BENCH_EXTRA} 2>&1 | tail -3 | tee gpurun_out/bench_c3.log
if [ "$1" == "launches" ]; then
echo "== launch list"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hh_ -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 0 > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log
fi
