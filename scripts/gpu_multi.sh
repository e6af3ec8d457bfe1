#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 2>&1 | tail -8 | tee gpurun_out/tests.log
echo "== bench N=1"
timeout 900 python bench.py --steps 1 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_n1.log | cut -c1-1500
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 1 --warmup 1 2>&1 | tail -12 | tee gpurun_out/bench_n$N.log | cut -c1-1500
