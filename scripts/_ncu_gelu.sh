set -x
cd $GRAFT_REPO_ROOT
export REPS=2
ONLY="ffn1 fwd gelu  M" ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 1 -f -o gpurun_out/prof_gemm_gelu2 python scripts/bench_gemm.py > gpurun_out/ncu3_a.log 2>&1
