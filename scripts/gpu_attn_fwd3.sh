#!/bin/bash
python -m pytest tests/test_attention_kernels.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
for cfg in "0 4096" "0 8192" "3000 4096" "3000 8192" "4500 8192"; do set -- $cfg
echo "== S=196 lag=$1 stg=$2"; VITSSL_ATTN_FWD_LAG=$1 VITSSL_ATTN_FWD_STG=$2 ONLY="attn fwd" python scripts/bench_kernels.py 2>&1 | grep "attn fwd"
done
echo "== S=37 B=768 direct"; VITSSL_ATTN_FWD_LAG=0 S=37 B=768 ONLY="attn fwd" python scripts/bench_kernels.py 2>&1 | grep "attn fwd"
echo "== S=37 B=768 staged"; VITSSL_ATTN_FWD_LAG=0 VITSSL_ATTN_FWD_STAGED=2 S=37 B=768 ONLY="attn fwd" python scripts/bench_kernels.py 2>&1 | grep "attn fwd"
for cfg in "0 4096 2" "0 8192 2" "0 8192 1"; do set -- $cfg
echo "== trace lag=$1 stg=$2 slots=$3"; VITSSL_ATTN_FWD_LAG=$1 VITSSL_ATTN_FWD_STG=$2 VITSSL_ATTN_FWD_SLOTS=$3 VITSSL_LIB=vit-ssl_b200/lib/libvitssl_b200_trace.so python scripts/trace_attn_fwd.py | tail -1
done
