#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_r2.log
python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2e_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench_simmim.log 2>&1; echo "bench rc=$?"; head -c 250 gpurun_out/r2e_bench_simmim.log; echo
VITSSL_PDL=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/r2e_bench_simmim_pdl.log 2>&1; echo "pdl rc=$?"; head -c 250 gpurun_out/r2e_bench_simmim_pdl.log; echo
python bench.py --steps 20 --warmup 5 --optimizer torch --no-cpu-baseline --no-torch-baseline > gpurun_out/r2e_bench_simmim_torchopt.log 2>&1; echo "torchopt rc=$?"; head -c 250 gpurun_out/r2e_bench_simmim_torchopt.log; echo
