#!/bin/bash
for pdl in 0 1 0 1; do
VITSSL_PDL=$pdl python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('simmim PDL=$pdl', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['clocks']['sm_mhz'])"
done
for pdl in 0 1; do
VITSSL_PDL=$pdl python bench.py --workload dino --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('dino PDL=$pdl', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['clocks']['sm_mhz'])"
done
VITSSL_PDL=1 python -m pytest tests/test_models.py tests/test_baseline_shapes.py tests/test_encoder_stack.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
