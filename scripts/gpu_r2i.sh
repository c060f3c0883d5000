#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_r2.log
python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2i_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench_simmim.log 2>&1; echo "bench rc=$?"; head -c 250 gpurun_out/r2i_bench_simmim.log; echo
python bench.py --workload dino --steps 10 --warmup 3 > gpurun_out/r2i_bench_dino.log 2>&1; echo "dino rc=$?"; head -c 250 gpurun_out/r2i_bench_dino.log; echo
python bench.py --arch vit_b --steps 10 --warmup 3 > gpurun_out/r2i_bench_vitb.log 2>&1; echo "vitb rc=$?"; head -c 250 gpurun_out/r2i_bench_vitb.log; echo
python bench.py --workload dino --arch vit_b --batch 64 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_dino_vitb.log 2>&1; echo "dino vitb rc=$?"; head -c 250 gpurun_out/r2i_bench_dino_vitb.log; echo
python -c "import __graft_entry__ as g; g.smoke()"
