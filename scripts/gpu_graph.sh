#!/bin/bash
python -m pytest tests/test_round2_kernels.py -m gpu -q -x -p no:cacheprovider -k "graph_replay" 2>&1 | tail -8
python -m pytest tests/test_gemm.py tests/test_layernorm_kernels.py tests/test_models.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for g in 1 0; do
VITSSL_GRAPH=$g python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline 2>&1 | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('simmim graph=$g', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], 'host', d['host_issue_ms_per_step'], d['clocks']['sm_mhz'])"
done
