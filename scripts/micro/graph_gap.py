"""Developer microbenchmark (GPU box): does a CUDA graph shorten the kernel-to-kernel gaps of the
persistent kernels? 96 dependent launches (GEMM -> LayerNorm pairs at the ViT-S shapes) timed as a
stream of launches and as one graph replay."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vit-ssl_b200"))
import torch
from vit_core._backend import ops
M, D = 50176, 384
x = torch.randn(M, D, device="cuda").to(torch.bfloat16)
w = (torch.randn(3 * D, D, device="cuda") * 0.05).to(torch.bfloat16)
wo = (torch.randn(D, D, device="cuda") * 0.05).to(torch.bfloat16)
outs = [torch.empty(M, 3 * D, device="cuda", dtype=torch.bfloat16) for _ in range(2)]
outo = [torch.empty(M, D, device="cuda", dtype=torch.bfloat16) for _ in range(2)]
N = 48

def chain():
    for i in range(N):
        ops.gemm(x, w, out=outs[i & 1])
        ops.gemm(outs[i & 1][:, :D], wo, out=outo[i & 1])

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

try:
    t_stream = timeit(chain)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        chain()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            chain()
    t_graph = timeit(g.replay)
    print(f"{2 * N} launches: stream {t_stream * 1e3:.1f} us, graph {t_graph * 1e3:.1f} us, "
          f"difference {(t_stream - t_graph) * 1e3 / (2 * N):.2f} us per launch")
except Exception as e:
    print("failed:", repr(e)[:300])
