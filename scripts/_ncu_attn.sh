cd $GRAFT_REPO_ROOT
export REPS=2
ONLY="attn bwd" ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 2 -c 1 -f -o gpurun_out/prof_attn_bwd2 python scripts/bench_kernels.py > gpurun_out/ncu4.log 2>&1
