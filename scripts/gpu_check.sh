#!/bin/bash
# One gpurun call: parity tests, smoke, sanitizer pass, first bench lines.  Everything is logged
# under gpurun_out/ so the results survive the call.
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,driver_version --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os; print('host cores', os.cpu_count())" >> gpurun_out/gpu.txt
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 -rA 2>&1 | tail -60 | tee gpurun_out/tests.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20 | tee gpurun_out/smoke.log
if [ "$1" != "nosan" ]; then
echo "== compute-sanitizer memcheck (smoke)"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -40 | tee gpurun_out/sanitizer.log
fi
echo "== bench C2 (10k contigs / 50M pairs)"
timeout 900 python bench.py --contigs 10000 --nchr 16 --mean-len 30000 --pairs 50000000 --steps 1 --warmup 1 2>&1 | tail -5 | tee gpurun_out/bench_c2.log
echo "== bench C3 (50k contigs / 200M pairs)"
timeout 1200 python bench.py --steps 1 --warmup 1 2>&1 | tail -5 | tee gpurun_out/bench_c3.log
echo "== done"
