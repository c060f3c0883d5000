#!/bin/bash
python -m pytest tests/test_attention_kernels.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
for stg in 4096 8192; do
echo "== S=196 lag=0 stg=$stg"; VITSSL_ATTN_FWD_LAG=0 VITSSL_ATTN_FWD_STG=$stg ONLY="attn fwd" python scripts/bench_kernels.py 2>&1 | grep "attn fwd"
done
echo "== S=196 lag=3000 stg=4096"; VITSSL_ATTN_FWD_LAG=3000 VITSSL_ATTN_FWD_STG=4096 ONLY="attn fwd" python scripts/bench_kernels.py 2>&1 | grep "attn fwd"
echo "== S=37 B=768 direct"; VITSSL_ATTN_FWD_LAG=0 S=37 B=768 ONLY="attn fwd" python scripts/bench_kernels.py 2>&1 | grep "attn fwd"
echo "== S=37 B=768 staged"; VITSSL_ATTN_FWD_LAG=0 VITSSL_ATTN_FWD_STAGED=2 S=37 B=768 ONLY="attn fwd" python scripts/bench_kernels.py 2>&1 | grep "attn fwd"
echo "== trace"; VITSSL_ATTN_FWD_LAG=0 VITSSL_ATTN_FWD_STG=4096 VITSSL_LIB=vit-ssl_b200/lib/libvitssl_b200_trace.so python scripts/trace_attn_fwd.py | tail -2
