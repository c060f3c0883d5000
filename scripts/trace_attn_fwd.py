"""Developer tool (GPU box): per-item timeline of CTA 0 of the attention forward kernel.
Needs the library built with the trace hooks:
  VITSSL_EXTRA_NVCC_FLAGS=-DVITSSL_ATTN_TRACE python vit-ssl_b200/build.py --force"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ssl_b200"))
import torch
from vit_core._backend import ops, lib
B, S, H, D = int(os.environ.get("B", 256)), int(os.environ.get("S", 196)), 6, 384
qkv = (torch.randn(B, S, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
for _ in range(3):
    ops.attention_fwd(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], H, 0.125)
torch.cuda.synchronize()
n = 4096 + 16 * 12
buf = (ctypes.c_longlong * n)()
l = lib.lib()
l.vitssl_debug_attn_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert l.vitssl_debug_attn_trace(ctypes.addressof(buf), n) == 0
names = {4: "ep:prev_stores_read", 5: "ep:staged", 6: "ep:fenced", 14: "ep:tma_issued", 15: "ep:mid_wait_done", 0: "mma:qk_landed", 1: "mma:tmem_free", 2: "mma:p_ready", 3: "mma:v_landed", 7: "sm:top", 8: "sm:s_ready",
         9: "sm:pass1_done", 10: "sm:pass2_done", 11: "sm:o_ready", 12: "sm:o_in_regs", 13: "sm:item_done"}
t0 = buf[4096 + 16 * 2 + 7]
for it in range(2, 8):
    ev = sorted((buf[4096 + 16 * it + k] - t0, names[k]) for k in names if buf[4096 + 16 * it + k])
    print(f"it={it}: " + "  ".join(f"{n}@{t}" for t, n in ev))
