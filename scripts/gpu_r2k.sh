#!/bin/bash
python -m pytest tests/test_round2_kernels.py tests/test_models.py tests/test_reference_trainers.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
python scripts/profile_gaps.py 2>&1 | grep -v Warn | head -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_simmim.log 2>&1; echo "bench rc=$?"; python - <<'P'
import json
for f in ["gpurun_out/r2k_bench_simmim.log"]:
    d=json.loads(open(f).readline()); print({k:d[k] for k in ("value","ms_per_step","e2e","e2e_u8","host_issue_ms_per_step")})
P
python bench.py --workload dino --steps 10 --warmup 3 > gpurun_out/r2k_bench_dino.log 2>&1; echo "dino rc=$?"; python - <<'P'
import json
for f in ["gpurun_out/r2k_bench_dino.log"]:
    d=json.loads(open(f).readline()); print({k:d[k] for k in ("value","ms_per_step","e2e","e2e_u8","host_issue_ms_per_step")})
P
