#!/bin/bash
python -m pytest tests/test_attention_kernels.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
echo "== S=196"; ONLY="attn" python scripts/bench_kernels.py 2>&1 | grep "attn"
echo "== S=197"; S=197 ONLY="attn" python scripts/bench_kernels.py 2>&1 | grep "attn"
echo "== S=37 B=768"; S=37 B=768 ONLY="attn" python scripts/bench_kernels.py 2>&1 | grep "attn"
echo "== S=128"; S=128 ONLY="attn" python scripts/bench_kernels.py 2>&1 | grep "attn"
VITSSL_LIB=vit-ssl_b200/lib/libvitssl_b200_trace.so python scripts/trace_attn_bwd.py | head -5 | cut -c1-400
