#!/bin/bash
# GPU box: plain bench run, then the ncu launch list of the same command (gpurun_out/r2_launches.csv;
# summarise with scripts/summarize_launches.py, which also writes profiles/traffic.json).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-torch-baseline"
$CMD > gpurun_out/r2_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 900 -c 700 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launch.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/r2_ncu_launch.log | cut -c1-200
