set -x
cd $GRAFT_REPO_ROOT
export REPS=2
ONLY="ffn1 fwd gelu  M" python scripts/bench_gemm.py > gpurun_out/ncu2_plain.log 2>&1 || exit 1
ONLY="ffn1 fwd gelu  M" ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 1 -f -o gpurun_out/prof_gemm_gelu python scripts/bench_gemm.py > gpurun_out/ncu2_a.log 2>&1
ONLY="ffn2 dgrad dgelu M" ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 1 -f -o gpurun_out/prof_gemm_dgelu python scripts/bench_gemm.py > gpurun_out/ncu2_b.log 2>&1
ONLY="ffn wgrad FxD" ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 1 -f -o gpurun_out/prof_gemm_wgrad python scripts/bench_gemm.py > gpurun_out/ncu2_c.log 2>&1
REPS=20 python scripts/bench_gemm.py > gpurun_out/gemm_all.log 2>&1
