#!/bin/bash
for i in 1 2; do
for pair in 1 0; do
VITSSL_GEMM_PAIR=$pair python bench.py --workload dino --steps 10 --warmup 3 --no-cpu-baseline --no-torch-baseline 2>&1 | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('dino pair=$pair', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'gemm', d['roofline']['frac'], d['roofline']['ms_per_step_in_kernel'], d['clocks']['sm_mhz'])"
done; done
