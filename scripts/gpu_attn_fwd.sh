#!/bin/bash
python -m pytest tests/test_attention_kernels.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
echo "== S=196"; ONLY="attn fwd" python scripts/bench_kernels.py 2>&1 | grep "attn fwd"
echo "== S=196 shared lo staging"; VITSSL_ATTN_FWD_STG=4096 ONLY="attn fwd" python scripts/bench_kernels.py 2>&1 | grep "attn fwd"
echo "== S=197"; S=197 ONLY="attn fwd" python scripts/bench_kernels.py 2>&1 | grep "attn fwd"
echo "== S=37 B=768 direct"; S=37 B=768 ONLY="attn fwd" python scripts/bench_kernels.py 2>&1 | grep "attn fwd"
echo "== S=37 B=768 staged"; VITSSL_ATTN_FWD_STAGED=2 S=37 B=768 ONLY="attn fwd" python scripts/bench_kernels.py 2>&1 | grep "attn fwd"
echo "== trace"; VITSSL_LIB=vit-ssl_b200/lib/libvitssl_b200_trace.so python scripts/trace_attn_fwd.py | tail -3
VITSSL_ATTN_FWD_STAGED=2 VITSSL_LIB=vit-ssl_b200/lib/libvitssl_b200_trace.so S=37 B=768 python scripts/trace_attn_fwd.py | tail -2
