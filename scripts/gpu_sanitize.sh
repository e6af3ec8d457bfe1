#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== memcheck: smoke + link tests + small mcl tests"
timeout 1200 compute-sanitizer --tool memcheck --print-limit 10 --error-exitcode 9 python -m pytest tests/test_gpu_links.py tests/test_gpu_mcl.py -m gpu -q -p no:cacheprovider -k "golden or edge or fragment or rank_sums or step_interface" 2>&1 | tail -15 | tee gpurun_out/memcheck.log
echo "== racecheck: golden MCL (small) + smoke"
timeout 1200 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_mcl.py -m gpu -q -p no:cacheprovider -k "golden and links_a or step_interface or rank_sums" 2>&1 | tail -25 | tee gpurun_out/racecheck.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee -a gpurun_out/racecheck.log
echo "== bench N=1 e2e check"
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','e2e','stage_ms')}, d['mcl']['value'], d['mcl']['e2e'])"
