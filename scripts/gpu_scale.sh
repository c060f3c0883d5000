#!/bin/bash
# usage: bash scripts/gpu_scale.sh N   (inside gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
run() {  # tag, extra env..., then bench args after --
  tag=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/r2_scale_${tag}_n$N.log 2>&1
  echo "$tag rc=$?"; grep "^{" gpurun_out/r2_scale_${tag}_n$N.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d.get(k) for k in ('value', 'ms_per_step', 'comm_exposed_ms', 'dp_check')}, 'e2e', d['e2e']['value'])"
}
run simmim --steps 20 --warmup 5
VITSSL_DP_CHUNKS=6 run simmim_c6 --steps 20 --warmup 5
run dino --workload dino --steps 10 --warmup 3
