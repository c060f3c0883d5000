"""Developer tool (GPU box): compute-bound square GEMMs through vitssl_gemm_bf16 vs torch.matmul (cuBLAS)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ssl_b200"))
import torch
from vit_core._backend import ops
bf = torch.bfloat16
import ast
SHAPES = ast.literal_eval(os.environ.get('SHAPES', '[(8192, 8192, 8192), (8192, 4096, 4096), (50176, 1536, 1536), (50176, 384, 384)]'))
for (M, N, K) in SHAPES:
    a = (torch.randn(M, K, device="cuda") * 0.1).to(bf)
    b = (torch.randn(N, K, device="cuda") * 0.1).to(bf)
    for name, fn in (("ours", lambda: ops.gemm(a, b)), ("cublas", lambda: torch.matmul(a, b.t()))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        print(f"{M}x{N}x{K} {name:7s} {us:9.1f} us {2.0*M*N*K/us*1e-6:8.1f} TFLOP/s", flush=True)
