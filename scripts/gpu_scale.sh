#!/bin/bash
# usage: bash scripts/gpu_scale.sh N [workloads...]   (inside gpurun --gpus N)
N=${1:-2}; shift
mkdir -p gpurun_out
run() {  # tag, then bench args
  tag=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/r2_scale_${tag}_n$N.log 2>&1
  echo "$tag rc=$?"; grep "^{" gpurun_out/r2_scale_${tag}_n$N.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d.get(k) for k in ('value', 'ms_per_step', 'comm_exposed_ms', 'dp_check')}, 'e2e', d['e2e']['value'])"
}
for w in ${@:-simmim dino}; do
  case $w in
    simmim) run simmim --steps 20 --warmup 5 ;;
    dino) run dino --workload dino --steps 10 --warmup 3 ;;
    vitb) run vitb --arch vit_b --steps 10 --warmup 3 ;;
    dino_vitb) run dino_vitb --workload dino --arch vit_b --batch 64 --steps 6 --warmup 3 ;;
  esac
done
