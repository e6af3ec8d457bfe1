#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 $BENCH_EXTRA 2> gpurun_out/bench_n$N.err | grep -v "^\*\|OMP_NUM" | tail -6 | tee gpurun_out/bench_n$N.log | cut -c1-2500
grep "e2e build sections\|Error\|error" gpurun_out/bench_n$N.err | head -20
