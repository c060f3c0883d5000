#!/bin/bash
# GPU box: one `ncu --set full` capture per hot kernel at the ViT-S/16 B=256 shapes (reports land in
# gpurun_out/, key rows are summarised in profiles/README.md). Each micro-benchmark command first
# runs plain and must exit 0 before the same command runs under ncu.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
export REPS=2
cap() {  # name, kernel regex, script, ONLY filter, launches to skip
  ONLY="$4" python "$3" > "gpurun_out/cap_$1.log" 2>&1 &&
  ONLY="$4" ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$5" -c 1 -f \
    -o "gpurun_out/r2_prof_$1" python "$3" > /dev/null 2>&1
  echo "$1 rc=$?"
}
cap gemm_gelu_d gemm_tcgen05 scripts/bench_gemm.py    "ffn1 fwd gelu_d" 3
cap gemm_mul    gemm_tcgen05 scripts/bench_gemm.py    "ffn2 dgrad mul" 3
cap gemm_qkv    gemm_tcgen05 scripts/bench_gemm.py    "qkv fwd" 3
cap gemm_wgrad  gemm_tcgen05 scripts/bench_gemm.py    "ffn wgrad FxD" 3
cap attn_bwd    attn_bwd     scripts/bench_kernels.py "attn bwd" 2
cap attn_fwd    attn_fwd     scripts/bench_kernels.py "attn fwd" 5
cap attn_delta  attn_delta   scripts/bench_kernels.py "attn bwd" 2
cap ln_bwd      ln_bwd4      scripts/bench_kernels.py "ln bwd" 2
cap ln_fwd      ln_fwd4      scripts/bench_kernels.py "ln fwd add" 5
