#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_r2.log
python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2d_tests.log
python scripts/diag_attention_precision.py > gpurun_out/r2d_attn_precision.log 2>&1; head -8 gpurun_out/r2d_attn_precision.log
ONLY=attn REPS=30 python scripts/bench_kernels.py > gpurun_out/r2d_attn.log 2>&1; cat gpurun_out/r2d_attn.log
WORKLOAD=simmim OPT=vitssl python scripts/profile_step.py > gpurun_out/r2d_prof_simmim.log 2>&1; head -5 gpurun_out/r2d_prof_simmim.log
WORKLOAD=simmim OPT=torch python scripts/profile_step.py > gpurun_out/r2d_prof_simmim_t.log 2>&1; head -5 gpurun_out/r2d_prof_simmim_t.log
WORKLOAD=dino OPT=vitssl python scripts/profile_step.py > gpurun_out/r2d_prof_dino.log 2>&1; head -5 gpurun_out/r2d_prof_dino.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/r2d_bench_simmim.log 2>&1; echo "bench rc=$?"; head -c 300 gpurun_out/r2d_bench_simmim.log; echo
python bench.py --workload dino --steps 8 --warmup 3 --no-cpu-baseline --no-torch-baseline > gpurun_out/r2d_bench_dino.log 2>&1; echo "dino rc=$?"; head -c 300 gpurun_out/r2d_bench_dino.log; echo
