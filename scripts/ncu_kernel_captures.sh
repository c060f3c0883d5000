#!/bin/bash
# GPU box: one `ncu --set full` capture per hot kernel at the ViT-S/16 B=256 shapes (reports land in
# gpurun_out/, key rows are summarised by hand in profiles/README.md section 3). Run only after the
# same benchmarks exit 0 without ncu.
cd "${GRAFT_REPO_ROOT:-.}"
export REPS=2
cap() {  # name, kernel regex, script, ONLY filter, launches to skip
  ONLY="$4" ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$5" -c 1 -f \
    -o "gpurun_out/prof_$1" python "$3" > /dev/null 2>&1
}
cap gemm_gelu  gemm_tcgen05 scripts/bench_gemm.py    "ffn1 fwd gelu  M" 3
cap gemm_dgelu gemm_tcgen05 scripts/bench_gemm.py    "ffn2 dgrad dgelu M" 3
cap gemm_wgrad gemm_tcgen05 scripts/bench_gemm.py    "ffn wgrad FxD" 3
cap attn_bwd   attn_bwd     scripts/bench_kernels.py "attn bwd" 2
cap attn_fwd   attn_fwd     scripts/bench_kernels.py "attn fwd" 2
cap ln_bwd     ln_bwd       scripts/bench_kernels.py "ln bwd" 2
