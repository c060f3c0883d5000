#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_r2.log
python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2f_tests.log
VITSSL_ATTN_FWD_STAGED=0 python -m pytest tests/test_attention_kernels.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
for st in 0 1; do echo "== staged=$st"; VITSSL_ATTN_FWD_STAGED=$st ONLY=attn REPS=40 python scripts/bench_kernels.py; VITSSL_ATTN_FWD_STAGED=$st ONLY="attn fwd" S=37 B=768 REPS=40 python scripts/bench_kernels.py; done > gpurun_out/r2f_attn.log 2>&1; cat gpurun_out/r2f_attn.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/r2f_bench_simmim.log 2>&1; echo "bench rc=$?"; head -c 250 gpurun_out/r2f_bench_simmim.log; echo
VITSSL_ATTN_FWD_STAGED=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/r2f_bench_simmim_direct.log 2>&1; head -c 250 gpurun_out/r2f_bench_simmim_direct.log; echo
