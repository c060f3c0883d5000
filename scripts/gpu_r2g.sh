#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_attention_kernels.py tests/test_models.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
ONLY=attn REPS=40 python scripts/bench_kernels.py; ONLY="attn" S=37 B=768 REPS=40 python scripts/bench_kernels.py
bash scripts/ncu_launch_list.sh
bash scripts/ncu_kernel_captures.sh
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
