#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_r2.log
python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2c_tests.log
python scripts/diag_attention_precision.py > gpurun_out/r2c_attn_precision.log 2>&1; cat gpurun_out/r2c_attn_precision.log
ONLY=ffn REPS=30 python scripts/bench_gemm.py > gpurun_out/r2c_gemm.log 2>&1; head -4 gpurun_out/r2c_gemm.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/r2c_bench_simmim.log 2>&1; echo "bench rc=$?"; head -c 300 gpurun_out/r2c_bench_simmim.log; echo
python bench.py --workload dino --steps 8 --warmup 3 --no-cpu-baseline --no-torch-baseline > gpurun_out/r2c_bench_dino.log 2>&1; echo "dino rc=$?"; head -c 300 gpurun_out/r2c_bench_dino.log; echo
