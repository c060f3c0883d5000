#!/bin/bash
# GPU box: plain bench run, then the ncu launch list of the same command (profiles/r1_launches.csv;
# summarise with scripts/summarize_launches.py, which also writes profiles/traffic.json).
cd $GRAFT_REPO_ROOT
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1100 -c 700 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_ncu_launch.log 2>&1
tail -2 gpurun_out/r1b_ncu_launch.log | cut -c1-200
