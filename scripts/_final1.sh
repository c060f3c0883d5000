cd $GRAFT_REPO_ROOT
python bench.py --steps 20 --warmup 5 > gpurun_out/r1c_bench_n1.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1c_bench_reference.log 2>&1
export REPS=2
ONLY="ffn1 fwd gelu  M" ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 1 -f -o gpurun_out/prof_r1c_gemm_gelu python scripts/bench_gemm.py > /dev/null 2>&1
ONLY="ffn wgrad FxD" ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 1 -f -o gpurun_out/prof_r1c_gemm_wgrad python scripts/bench_gemm.py > /dev/null 2>&1
ONLY="attn bwd" ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 2 -c 1 -f -o gpurun_out/prof_r1c_attn_bwd python scripts/bench_kernels.py > /dev/null 2>&1
ONLY="attn fwd" ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -f -o gpurun_out/prof_r1c_attn_fwd python scripts/bench_kernels.py > /dev/null 2>&1
ONLY="ln bwd" ncu --set full --clock-control none --import-source on -k regex:ln_bwd -s 2 -c 1 -f -o gpurun_out/prof_r1c_ln_bwd python scripts/bench_kernels.py > /dev/null 2>&1
timeout 100 python scripts/bench_gemm.py > gpurun_out/r1c_gemm_shapes.log 2>&1
timeout 100 python scripts/bench_kernels.py > gpurun_out/r1c_kernels.log 2>&1
tail -1 gpurun_out/r1c_bench_n1.log | cut -c1-250
