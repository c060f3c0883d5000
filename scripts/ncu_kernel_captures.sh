#!/bin/bash
# GPU box: one `ncu --set full` capture per hot kernel at the ViT-S/16 B=256 shapes. Each
# micro-benchmark command first runs plain and must exit 0 before the same command runs under ncu.
# gpurun returns at most 64 MiB, so every report is reduced on the box to its raw metric table
# (gpurun_out/r2_prof_<name>_raw.csv) and, for the small reports, the per-line source table; the
# GEMM reports (19 MB each with embedded source) are deleted after the export.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
export REPS=2
cap() {  # name, kernel regex, script, ONLY filter, launches to skip, keep report (0/1)
  ONLY="$4" python "$3" > "gpurun_out/cap_$1.log" 2>&1 &&
  ONLY="$4" ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$5" -c 1 -f \
    -o "gpurun_out/r2_prof_$1" python "$3" > /dev/null 2>&1
  echo "$1 rc=$?"
  ncu -i "gpurun_out/r2_prof_$1.ncu-rep" --page raw --csv > "gpurun_out/r2_prof_$1_raw.csv" 2>/dev/null
  ncu -i "gpurun_out/r2_prof_$1.ncu-rep" --page source --csv 2>/dev/null | head -c 3000000 > "gpurun_out/r2_prof_$1_source.csv"
  if [ "$6" != "1" ]; then rm -f "gpurun_out/r2_prof_$1.ncu-rep"; fi
}
cap gemm_gelu_d gemm_tcgen05 scripts/bench_gemm.py    "ffn1 fwd gelu_d" 3 0
cap gemm_mul    gemm_tcgen05 scripts/bench_gemm.py    "ffn2 dgrad mul" 3 0
cap gemm_qkv    gemm_tcgen05 scripts/bench_gemm.py    "qkv fwd" 3 0
cap attn_bwd    attn_bwd     scripts/bench_kernels.py "attn bwd" 2 1
cap attn_fwd    attn_fwd     scripts/bench_kernels.py "attn fwd" 5 1
cap attn_delta  attn_delta   scripts/bench_kernels.py "attn bwd" 2 0
cap ln_bwd      ln_bwd4      scripts/bench_kernels.py "ln bwd" 2 1
cap ln_fwd      ln_fwd4      scripts/bench_kernels.py "ln fwd add" 5 0
du -sh gpurun_out
