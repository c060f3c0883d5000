#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_r2.log
python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2j_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench_simmim.log 2>&1; echo "bench rc=$?"; head -c 250 gpurun_out/r2j_bench_simmim.log; echo
python bench.py --workload dino --steps 10 --warmup 3 > gpurun_out/r2j_bench_dino.log 2>&1; echo "dino rc=$?"; head -c 250 gpurun_out/r2j_bench_dino.log; echo
timeout 600 python bench.py --impl reference --device cuda --torch-compile --steps 5 > gpurun_out/r2j_ref_compile_simmim.log 2>&1; echo "compile rc=$?"; tail -c 500 gpurun_out/r2j_ref_compile_simmim.log; echo
timeout 600 python bench.py --impl reference --device cuda --torch-compile --workload dino --steps 5 > gpurun_out/r2j_ref_compile_dino.log 2>&1; echo "compile dino rc=$?"; tail -c 500 gpurun_out/r2j_ref_compile_dino.log; echo
