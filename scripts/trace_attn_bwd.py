"""Developer tool (GPU box): per-iteration timeline of CTA 0 of the attention backward kernel.
Needs the library built with the trace hooks:
  VITSSL_EXTRA_NVCC_FLAGS=-DVITSSL_ATTN_TRACE python vit-ssl_b200/build.py --force"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ssl_b200"))
import torch
from vit_core._backend import ops, lib
B, S, H, D = int(os.environ.get("B", 256)), int(os.environ.get("S", 196)), 6, 384
bf = torch.bfloat16
qkv = (torch.randn(B, S, 3 * D, device="cuda") * 0.5).to(bf)
dctx = (torch.randn(B, S, D, device="cuda") * 0.5).to(bf)
dqkv = torch.empty_like(qkv)
ctx, lse, _lo = ops.attention_fwd(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], H, 0.125)
for _ in range(3):
    ops.attention_bwd(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], ctx, dctx, lse, H, 0.125,
                      dqkv[..., :D], dqkv[..., D:2 * D], dqkv[..., 2 * D:])
torch.cuda.synchronize()
n = 16 * 44
buf = (ctypes.c_longlong * n)()
l = lib.lib()
l.vitssl_debug_attn_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert l.vitssl_debug_attn_trace(ctypes.addressof(buf), n) == 0
t0 = buf[8]
names = {0: "mma:sdp_read", 1: "mma:sdp_issued", 2: "mma:pds_ready", 3: "mma:grads_issued", 8: "math:top", 9: "math:sdp_full",
         10: "math:c0done", 11: "math:pds_free", 12: "math:stored", 13: "drain:start", 14: "drain:tmem_ld", 15: "drain:dkv_stored"}
for g in range(4, 12):
    ev = sorted((buf[16 * g + k] - t0, names[k]) for k in names if buf[16 * g + k])
    print(f"g={g} it={g % 4}: " + "  ".join(f"{n}@{t}" for t, n in ev))
