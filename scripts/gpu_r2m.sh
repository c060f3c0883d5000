#!/bin/bash
python -m pytest tests/test_gemm.py tests/test_models.py tests/test_baseline_shapes.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
for i in 1 2; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline 2>&1 | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('simmim', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'gemm', d['roofline']['frac'], d['roofline']['ms_per_step_in_kernel'], d['clocks']['sm_mhz'])"
VITSSL_GEMM_PAIR=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline 2>&1 | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('simmim nopair', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'gemm', d['roofline']['frac'], d['roofline']['ms_per_step_in_kernel'], d['clocks']['sm_mhz'])"
done
python bench.py --workload dino --steps 10 --warmup 3 --no-cpu-baseline --no-torch-baseline 2>&1 | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('dino', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'gemm', d['roofline']['frac'])"
