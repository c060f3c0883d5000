#!/bin/bash
# one GPU call: full GPU test-suite, LayerNorm / GEMM micro-benchmarks of the round-2 kernels, short bench
mkdir -p gpurun_out; rm -f gpurun_out/parity_r2.log
python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2b_tests.log
for cfg in "VITSSL_LN_IMPL=0" "VITSSL_LN_IMPL=1" "VITSSL_LN_IMPL=1 VITSSL_LN_BWD_PREFETCH=0" "VITSSL_LN_IMPL=1 VITSSL_LN_BWD_CAP=4" "VITSSL_LN_IMPL=1 VITSSL_LN_BWD_CAP=12" "VITSSL_LN_IMPL=1 VITSSL_LN_BWD_CAP=24"; do
  echo "== $cfg"; env $cfg ONLY=ln REPS=30 python scripts/bench_kernels.py
done > gpurun_out/r2b_ln.log 2>&1
cat gpurun_out/r2b_ln.log
ONLY=ffn REPS=30 python scripts/bench_gemm.py > gpurun_out/r2b_gemm.log 2>&1; cat gpurun_out/r2b_gemm.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-baseline > gpurun_out/r2b_bench_simmim.log 2>&1; echo "bench rc=$?"; head -c 300 gpurun_out/r2b_bench_simmim.log
