#!/usr/bin/env python
"""bench.py — headline benchmark: pretraining images/s of ViT-S/16 SimMIM (and DINO) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload simmim|dino] [--arch vit_s|vit_b] [--batch B]

One "step" is the reference trainer's loop body, called through the drop-in API exactly as the
unmodified trainer calls it:
  simmim  utils/trainers/simmim_trainer.py:63-71 — zero_grad -> autocast(bf16): `model(x)` +
          `nn.L1Loss` -> GradScaler-scaled backward -> scaler.step(AdamW) -> scaler.update()
  dino    utils/trainers/dino_trainer.py:85-105 — same with `model(views, 2)`, DINOLoss on
          model.center, then `model.momentum_update_teacher(m)`
Default workload = BASELINE.json configs[1]: ViT-S/16 (D384 L12 H6 F1536), 224x224, mask ratio 0.6,
batch 256 per GPU, synthetic torch.rand images, random-init weights, dropout 0.1
(configs/base/model.yaml:7). `--workload dino` = configs[2] (2x224 + 6x96 crops, K = 65536, B = 128
per GPU, EMA teacher); `--arch vit_b` = configs[3]/[4] (B = 128). Under torchrun (N > 1) the batch
is sharded per rank (weak scaling), gradients are averaged inside the modules
(vit_core._backend.dp) and the DINO center is summed across ranks.

JSON keys beyond the base contract: `roofline` (tcgen05 GEMM family), `roofline_attn` (attention
forward / backward kernels), `roofline_hbm` (fused add+LayerNorm family) — all three measured on the
SAME code path as the timed region: CUDA events recorded around every launch by the C-side
sequencer (csrc/encoder.cu, vitssl_profile_*) and by the Python wrappers for the few launches
outside the encoder stack, in one extra step after the timed regions; `cpu_baseline` (the
reference's own modules from baseline/_ref on the host cores, bounded sample); `torch_gpu` (those
same unmodified modules on THIS GPU under autocast(bf16), eager — what the reference does today on
the same hardware); `e2e` (step fed from pinned host memory with the loss read back every step) and
`e2e_u8` (same with raw uint8 images, ToTensor on the device); `fused_objective` (the non-reference
`reconstruction_loss` entry); `gpu_launches`; `clocks`; `dp_check` and `comm_exposed_ms` at N > 1.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "vit-ssl_b200")

import torch  # noqa: E402

ARCHS = {
    "vit_s": dict(embed_dim=384, num_blocks=12, num_heads=6, mlp_dim=1536, patch_size=16),
    "vit_b": dict(embed_dim=768, num_blocks=12, num_heads=12, mlp_dim=3072, patch_size=16),
}
ARCH_NAME = {"vit_s": "ViT-S/16", "vit_b": "ViT-B/16"}
DINO_K = 65536


def metric_name(workload, arch, batch):
    if workload == "simmim":
        return f"pretrain images/sec {ARCH_NAME[arch]} SimMIM (224x224, mask 0.6, batch {batch}/GPU)"
    return f"pretrain images/sec {ARCH_NAME[arch]} DINO (2x224 + 6x96 crops, K=65536, EMA teacher, batch {batch}/GPU)"


def default_batch(workload, arch):
    return 256 if (workload == "simmim" and arch == "vit_s") else 128


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def traffic(family):
    """DRAM bytes per launch from the committed ncu launch list (profiles/traffic.json); None if absent."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[family]["dram_bytes_per_launch"]
    except Exception:
        return None


# ------------------------------------------------------------------------------------------
# algorithmic FLOPs per image: GEMMs only, 2mnk each; backward = 2x forward except
# the patch projection (no dX); the DINO teacher is forward-only
# ------------------------------------------------------------------------------------------
def flops_per_image(workload, arch):
    a = ARCHS[arch]
    D, L, F_ = a["embed_dim"], a["num_blocks"], a["mlp_dim"]
    P = 768

    def blocks(S):
        return L * (8 * S * D * D + 4 * S * S * D + 4 * S * D * F_)

    if workload == "simmim":
        N, n_m = 196, 117
        head = 2 * n_m * D * P
        patch = 2 * N * P * D
        return dict(fwd=blocks(N) + patch + head, step=3 * (blocks(N) + head) + 2 * patch)
    head_row = 2 * (D * 2048 + 2048 * 2048 + 2048 * D + D * DINO_K)
    g = blocks(197) + 2 * 196 * P * D
    l_ = blocks(37) + 2 * 36 * P * D
    student_fwd = 2 * g + 6 * l_ + 8 * head_row
    teacher_fwd = 2 * g + 2 * head_row
    patch = 2 * (2 * 196 + 6 * 36) * P * D
    return dict(fwd=student_fwd + teacher_fwd, step=3 * student_fwd - patch + teacher_fwd)


# ------------------------------------------------------------------------------------------
# reference arm: the UNMODIFIED reference (baseline/_ref) — on the host cores (cpu_baseline and the
# driver's `--impl reference`), or on this GPU under autocast (`--device cuda`, the torch_gpu key)
# ------------------------------------------------------------------------------------------
def _reference_step_factory(workload, arch, batch, device, compile_):
    """Model, step() and views built from the reference's own classes; falls back to the oracle port
    (CPU, SimMIM only) when baseline/_ref is absent."""
    sys.path.insert(0, ROOT)
    from baseline import refenv
    a = ARCHS[arch]
    kind = "reference"
    torch.manual_seed(42)
    if refenv.available():
        refenv.use_reference_vit_core()
        if workload == "simmim":
            from vit_core.ssl.simmim.model import SimMIMViT
            model = SimMIMViT(input_shape=(3, 224, 224), dropout=0.1, mask_ratio=0.6, **a)
            crit = torch.nn.L1Loss(reduction="mean")
        else:
            from vit_core.ssl.dino.loss import DINOLoss
            from vit_core.ssl.dino.model import DINOViT
            model = DINOViT(input_shape=(3, 224, 224), dropout=0.1, output_dim=DINO_K, center_momentum=0.9, **a)
            crit = DINOLoss(0.04, 0.1)
        model = model.to(device).train()
        run = torch.compile(model) if compile_ else model          # utils/model_builder.py:182-183
        opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-3)
        scaler = torch.amp.GradScaler("cuda", enabled=(device != "cpu"))
        if workload == "simmim":
            x = torch.rand(batch, 3, 224, 224, device=device)

            def step():
                opt.zero_grad(set_to_none=True)
                with torch.autocast(device_type="cuda", dtype=torch.bfloat16, enabled=(device != "cpu")):
                    pred, tgt = run(x)
                    loss = crit(pred, tgt)
                scaler.scale(loss).backward()
                scaler.step(opt)
                scaler.update()
                return loss.item()
        else:
            views = [torch.rand(batch, 3, 224, 224, device=device) for _ in range(2)] + \
                    [torch.rand(batch, 3, 96, 96, device=device) for _ in range(6)]

            def step():
                opt.zero_grad(set_to_none=True)
                with torch.autocast(device_type="cuda", dtype=torch.bfloat16, enabled=(device != "cpu")):
                    t, s = run(views, 2)
                    loss = crit(t.view(2, batch, -1), s.view(8, batch, -1), model.center)
                scaler.scale(loss).backward()
                scaler.step(opt)
                scaler.update()
                model.momentum_update_teacher(0.996)
                return loss.item()
        return step, kind
    # ---- oracle port (reference install absent): same algorithm restated in oracle/vit_ref.py
    if workload != "simmim" or device != "cpu":
        raise RuntimeError("baseline/_ref is not installed (python baseline/install_reference.py)")
    from oracle import vit_ref
    D, L, H, F_, p = a["embed_dim"], a["num_blocks"], a["num_heads"], a["mlp_dim"], a["patch_size"]
    P, N = 3 * p * p, (224 // p) ** 2

    def lin(o, i):
        return torch.randn(o, i) * (1.0 / i ** 0.5)

    w = {"projection.weight": lin(D, P), "projection.bias": torch.zeros(D), "mask_token": torch.randn(1, 1, D),
         "positional_embedding": torch.rand(1, N, D), "simmim_head.weight": lin(P, D), "simmim_head.bias": torch.zeros(P)}
    for i in range(L):
        b = f"encoder_blocks.{i}."
        for n in ("w_query", "w_key", "w_value", "final_linear"):
            w[b + f"self_attention.{n}.weight"] = lin(D, D)
        w[b + "feed_forward.linear_in.weight"], w[b + "feed_forward.linear_in.bias"] = lin(F_, D), torch.zeros(F_)
        w[b + "feed_forward.linear_out.weight"], w[b + "feed_forward.linear_out.bias"] = lin(D, F_), torch.zeros(D)
        for n in ("layer_norm1", "layer_norm2"):
            w[b + n + ".weight"], w[b + n + ".bias"] = torch.ones(D), torch.zeros(D)
    for v in w.values():
        v.requires_grad_(True)
    opt = torch.optim.AdamW(list(w.values()), lr=1e-4, weight_decay=1e-3)
    x = torch.rand(batch, 3, 224, 224)

    def step():
        opt.zero_grad(set_to_none=True)
        perms = torch.stack([torch.randperm(N) for _ in range(batch)])
        mask = vit_ref.mask_from_perms(perms, N, 0.6)
        pred, tg = vit_ref.simmim_forward(w, x, mask, patch_size=p, num_blocks=L, num_heads=H)
        loss = vit_ref.l1_loss(pred, tg)
        loss.backward()
        opt.step()
        return float(loss.detach())
    return step, "port"


def run_reference_arm(args):
    """Prints ONE JSON line (rank 0 only). CPU: all host threads, a bounded batch (the full
    per-GPU batch would take minutes per step). `--device cuda`: the same unmodified modules on
    the GPU at the full batch."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    device = args.device
    batch = args.batch if device != "cpu" else (args.ref_batch or (16 if args.workload == "simmim" else 4))
    if args.arch == "vit_b" and device == "cpu" and not args.ref_batch:
        batch = max(2, batch // 2)
    steps = max(1, min(args.steps, 4 if device == "cpu" else 10))
    warm = 1 if device == "cpu" else 3
    torch.set_num_threads(os.cpu_count() or 1)
    try:
        step, kind = _reference_step_factory(args.workload, args.arch, batch, device, args.torch_compile)
    except Exception as e:  # the reference could not be set up: say so, exit 0
        print(json.dumps({"impl": "reference", "unavailable": f"{type(e).__name__}: {e}"[:300]}), flush=True)
        return
    times = []
    for it in range(warm + steps):
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = step()
        if device != "cpu":
            torch.cuda.synchronize()
        if it >= warm:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    v = batch / (ms / 1e3)
    cores = os.cpu_count() or 1
    where = "CPU eager fp32" if device == "cpu" else ("B200 autocast(bf16) " + ("torch.compile" if args.torch_compile else "eager"))
    body = {"simmim": "fwd + nn.L1Loss + bwd + AdamW", "dino": "fwd + DINOLoss + bwd + AdamW + EMA"}[args.workload]
    line = {
        "impl": "reference", "metric": metric_name(args.workload, args.arch, args.batch), "value": round(v, 3),
        "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": round(ms, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if device == "cpu" else "bf16",
        "data": "synthetic",
        "config": {"workload": f"{ARCH_NAME[args.arch]} {args.workload} step, UNMODIFIED reference modules "
                               f"(baseline/_ref) on {where}" if kind == "reference" else
                               f"{ARCH_NAME[args.arch]} {args.workload} step, oracle port of the reference, {where}",
                   "batch_per_step": batch, "device": device, "loss": round(float(loss), 5)},
        "cpu_baseline": {"value": round(v, 3), "unit": "images/s", "cores": cores if device == "cpu" else 0, "kind": kind,
                         "sample": f"{steps} timed steps of batch {batch} (+{warm} warm-up), {body}"},
        "e2e": {"value": round(v, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def reference_subprocess(args, device, timeout=900, compile_=False):
    """Run the reference arm in its own process (its package is also called `vit_core`)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--arch", args.arch,
           "--batch", str(args.batch), "--device", device, "--steps", "3"]
    if compile_:
        cmd.append("--torch-compile")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": (r.stderr or "no output")[-300:]}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


# ------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 50 ms in the background; `window(t0, t1)` summarises the rows whose
    host timestamps fall inside a timed region (started well before it: nvidia-smi needs ~0.5 s)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def wait_ready(self, timeout):
        t_end = time.time() + timeout
        while self.proc is not None and len(self.rows) < 2 and time.time() < t_end and self.proc.poll() is None:
            time.sleep(0.05)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def window(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        for ts, r in self.rows:
            if ts < t0 or ts > t1 + 0.05:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def gemm_algorithmic_bytes(kind):
    """"gemm|MxNxK|a_mn=. b_mn=. epi=.": operands read once + result written once."""
    try:
        _, shp, flags = kind.split("|")
        M_, N_, K_ = (int(v) for v in shp.split("x"))
        fl = dict(kv.split("=") for kv in flags.split())
        out_b = 4 if fl.get("a_mn") == "1" else 2  # weight gradients are fp32
        return 2 * (M_ * K_ + N_ * K_) + M_ * N_ * out_b + (2 * M_ * N_ if fl.get("epi") in ("2", "3") else 0)
    except Exception:
        return None


def run_gpu_arm(args):
    sys.path.insert(0, PKG)
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from vit_core._backend import dp, lib, ops
    from vit_core.optim import FusedAdamW

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    lib.ensure_device()
    dev = torch.device("cuda", local)
    B, arch, wl = args.batch, ARCHS[args.arch], args.workload
    torch.manual_seed(42)
    if wl == "simmim":
        from vit_core.ssl.simmim import SimMIMViT
        model = SimMIMViT(input_shape=(3, 224, 224), dropout=args.dropout, mask_ratio=0.6, **arch).to(dev)
        crit = torch.nn.L1Loss(reduction="mean")                  # configs/simmim/training.yaml:2-5
        shapes = [(B, 3, 224, 224)]
    else:
        from vit_core.ssl.dino import DINOViT
        from vit_core.ssl.dino.loss import DINOLoss
        model = DINOViT(input_shape=(3, 224, 224), dropout=args.dropout, output_dim=DINO_K, center_momentum=0.9, **arch).to(dev)
        crit = DINOLoss(0.04, 0.1)
        shapes = [(B, 3, 224, 224)] * 2 + [(B, 3, 96, 96)] * 6
    model.train()
    if world > 1:
        dp.attach(model)
    trainable = [p for p in model.parameters() if p.requires_grad]
    if args.optimizer == "vitssl":   # training.optimizer.name=VitsslAdamW (utils/train_utils.py:27)
        opt = FusedAdamW(trainable, lr=1e-4, weight_decay=1e-3)
    else:
        opt = torch.optim.AdamW(trainable, lr=1e-4, weight_decay=1e-3)
    scaler = torch.amp.GradScaler("cuda")
    torch.manual_seed(1000 + rank)
    n_host = 3
    host = [[torch.rand(*s).pin_memory() for s in shapes] for _ in range(n_host)]
    host_u8 = [[(h * 255).round().to(torch.uint8).pin_memory() for h in hs] for hs in host]
    dev_in = [[h.to(dev) for h in hs] for hs in host]
    in_bytes = sum(h.numel() * 4 for h in host[0])

    phases = []  # host seconds per phase of every step (diagnostic)

    def step(batch, objective="dropin"):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if wl == "simmim":
                if objective == "fused":
                    loss = model.reconstruction_loss(batch[0])
                else:
                    pred, tgt = model(batch[0])
                    loss = crit(pred, tgt)
            else:
                t, s = model(batch, 2)
                loss = crit(t.view(2, B, -1), s.view(8, B, -1), model.center)
        t1 = time.perf_counter()
        scaler.scale(loss).backward()
        t2 = time.perf_counter()
        scaler.step(opt)
        t3 = time.perf_counter()
        scaler.update()
        if wl == "dino":
            model.momentum_update_teacher(0.996)
        t4 = time.perf_counter()
        phases.append((round((t1 - t0) * 1e3, 2), round((t2 - t1) * 1e3, 2), round((t3 - t2) * 1e3, 2), round((t4 - t3) * 1e3, 2)))
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    import gc
    gc_log = []       # (generation, ms) of every collector run inside a timed() call

    def _gc_cb(phase, info, _t=[0.0]):
        if phase == "start":
            _t[0] = time.perf_counter()
        else:
            gc_log.append((info.get("generation"), round((time.perf_counter() - _t[0]) * 1e3, 2)))
    gc.callbacks.append(_gc_cb)
    step_events = []  # per-step device times of the last timed() call (diagnostic: are slow regions a few stalls?)
    host_steps = []   # per-step host issue times of the same call

    def timed(n, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        del host_steps[:]
        del gc_log[:]
        del phases[:]
        ops.HOST_TRACE = []
        barrier()
        e0.record()
        marks[0].record()
        out = None
        for i in range(n):
            th = time.perf_counter()
            out = fn(i)
            marks[i + 1].record()
            host_steps.append((time.perf_counter() - th) * 1e3)
        e1.record()
        barrier()
        step_events[:] = [marks[i].elapsed_time(marks[i + 1]) for i in range(n)]
        return max_over_ranks(e0.elapsed_time(e1)), out

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        # nvidia-smi attaches to the driver for ~0.5 s (seconds in the first process on a fresh box) and
        # kernel launches crawl meanwhile: a timed region that overlaps it measured 16.8 instead of 14.8
        # ms/step. Wait for its first rows before anything is timed.
        sampler.wait_ready(20.0)
    barrier()
    # warm-up = a rehearsal of the timed loop itself (same retention of the previous step's loss object,
    # same run-ahead of the host): the caching allocator must reach its steady state here. With a bare
    # `step()` loop it still grew by three cudaMalloc calls in the second timed step (1-340 ms each)
    timed(args.warmup, lambda i: step(dev_in[i % n_host]))

    def settle(run_two_steps):
        """Extra untimed steps until nothing one-off is left to happen inside a timed region: the library
        captures a stack call into a CUDA graph at the second sighting of its argument set (a few ms each;
        with chunked data-parallel backward there are more sets, and they need not all repeat within W
        steps), and the caching allocator may still grow. At most 8 extra steps; reported as `settle_steps`."""
        extra = 0
        for _ in range(4):
            before = (lib.graph_stats()[0], torch.cuda.memory_stats().get("num_device_alloc", 0))
            run_two_steps()
            extra += 2
            changed = (lib.graph_stats()[0], torch.cuda.memory_stats().get("num_device_alloc", 0)) != before
            if world > 1:  # every rank runs the same number of steps (they hold collectives)
                flag = torch.tensor([int(changed)], device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
                changed = bool(flag.item())
            if not changed:
                break
        return extra

    settle_steps = settle(lambda: timed(2, lambda i: step(dev_in[i % n_host])))

    # ---- timed region 1: device-resident inputs (3 rotating batches + GBs of activations >> L2) ----
    lib.launch_count(reset=True)
    ms0 = torch.cuda.memory_stats()
    t_host0 = time.time()
    ms_total, loss = timed(args.steps, lambda i: step(dev_in[i % n_host]))
    launches = lib.launch_count()
    t_host1 = time.time()
    per_step = {"device_ms": [round(v, 2) for v in step_events], "host_issue_ms": [round(v, 2) for v in host_steps],
                "loadavg_1min": round(os.getloadavg()[0], 2), "gc_runs": [g for g in gc_log if g[1] >= 1.0], "phases_fwd_bwd_opt_upd_ms": phases[:4], "stack_bwd_c_ms": [v for _, v in (ops.HOST_TRACE or [])[:4]],
                "cuda_malloc_calls": torch.cuda.memory_stats().get("num_device_alloc", 0) - ms0.get("num_device_alloc", 0),
                "cuda_free_calls": torch.cuda.memory_stats().get("num_device_free", 0) - ms0.get("num_device_free", 0),
                "reserved_gb": round(torch.cuda.memory_reserved() / 2**30, 2)}
    ops.HOST_TRACE = None
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    final_loss = float(loss.detach())

    # host time to ISSUE one step into an idle stream (no synchronisation inside, launch queue empty,
    # so the host never waits for the device): the margin by which the step is GPU-bound
    issue = []
    for i in range(3):
        barrier()
        th0 = time.perf_counter()
        step(dev_in[i % n_host])
        issue.append((time.perf_counter() - th0) * 1e3)
    host_issue_ms = statistics.median(issue)
    barrier()

    # ---- timed region 2: end to end (pinned host -> device each step, loss read back each step) ----
    copy_stream = torch.cuda.Stream()
    h2d_gbps = []
    e2e_diag = []

    def e2e_run(host_sets):
        bufs = [[torch.empty(h.shape, dtype=h.dtype, device=dev) for h in host_sets[0]] for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[i % 2])
                for b_, h in zip(bufs[i % 2], host_sets[i % n_host]):
                    b_.copy_(h, non_blocking=True)
                ready[i % 2].record(copy_stream)

        # pinned host -> device bandwidth of this box (reported next to e2e: a slow link shows here first)
        with torch.cuda.stream(copy_stream):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 0.0
            for _ in range(3):
                c0.record(copy_stream)
                for b_, h in zip(bufs[0], host_sets[0]):
                    b_.copy_(h, non_blocking=True)
                c1.record(copy_stream)
                c1.synchronize()
                nbytes = sum(h.numel() * h.element_size() for h in host_sets[0])
                best = max(best, nbytes / (c0.elapsed_time(c1) * 1e-3) / 1e9)
        h2d_gbps.append(round(best, 1))
        for ev in free:
            ev.record()
        walls = []

        def loop(nsteps):
            # the trainers' loop: next batch's copy in flight on the copy stream, step, loss read back
            prefetch(0)
            for i in range(nsteps):
                tw = time.perf_counter()
                if i + 1 < nsteps:
                    prefetch(i + 1)
                torch.cuda.current_stream().wait_event(ready[i % 2])
                l = step(bufs[i % 2])
                free[i % 2].record()
                _ = l.item()  # device -> host read of the step's loss, as the trainer does every step
                walls.append(round((time.perf_counter() - tw) * 1e3, 2))

        # W warm-up steps of the SAME loop (copies, look-ahead and loss retention included): on a fresh box
        # the first process measured 1.3-1.9 ms/step more in this region with a simpler warm-up
        loop(max(2, args.warmup))
        settle(lambda: loop(2))
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
        del gc_log[:]
        del walls[:]
        t0.record()
        loop(args.steps)
        t1.record()
        barrier()
        ms = max_over_ranks(t0.elapsed_time(t1))
        e2e_diag.append({"wall_ms": walls, "cuda_malloc_calls": torch.cuda.memory_stats().get("num_device_alloc", 0) - m0,
                         "gc_runs": [g for g in gc_log if g[1] >= 1.0]})
        return world * B * args.steps / (ms / 1e3), ms / args.steps

    e2e_value, e2e_ms = e2e_run(host)
    e2e8_value, e2e8_ms = e2e_run(host_u8)

    # ---- the fused-objective entry (non-reference API), device-resident, for comparison ----
    fused = None
    if wl == "simmim":
        timed(max(2, args.warmup), lambda i: step(dev_in[i % n_host], "fused"))
        settle(lambda: timed(2, lambda i: step(dev_in[i % n_host], "fused")))
        ms_f, _ = timed(args.steps, lambda i: step(dev_in[i % n_host], "fused"))
        fused = {"value": round(world * B * args.steps / (ms_f / 1e3), 1), "unit": "images/s",
                 "ms_per_step": round(ms_f / args.steps, 3), "api": "SimMIMViT.reconstruction_loss(x) (targets never materialised for the caller)"}

    clocks = None
    if rank == 0:
        clocks = sampler.window(t_host0, t_host1)
        if (clocks.get("samples") or 0) < 3:
            clocks = sampler.window(t_host0, time.time())
        sampler.stop()

    # ---- exposed communication: time the main stream spends waiting for the all-reduce stream ----
    comm_exposed = None
    if world > 1:
        dp.WAIT_TRACE = []
        step(dev_in[0])
        barrier()
        tr, dp.WAIT_TRACE = dp.WAIT_TRACE, None
        comm_exposed = max_over_ranks(sum(a.elapsed_time(b) for a, b in tr))

    # ---- instrumented step: per-launch CUDA events on the SAME path as the timed region ----
    ops.PROFILE = []
    lib.profile_begin()
    step(dev_in[0])
    barrier()
    recs = lib.profile_collect()
    py, ops.PROFILE = ops.PROFILE, None
    recs += [(k, a.elapsed_time(b), w) for (k, a, b, w) in py]
    roof = roof_hbm = roof_attn = None
    if rank == 0:
        pk = peaks()
        gemm = [(k, ms, w) for (k, ms, w) in recs if k.startswith("gemm")]
        ln = [(ms, w) for (k, ms, w) in recs if k == "add_layernorm"]
        if gemm:
            t = sum(ms for _, ms, _ in gemm); f = sum(w for _, _, w in gemm)
            ach = f / (t * 1e-3) / 1e12
            gb = [v for v in (gemm_algorithmic_bytes(k) for k, _, _ in gemm) if v]
            by_shape = {}
            for k, ms, w in gemm:
                e = by_shape.setdefault(k, [0, 0.0, 0.0]); e[0] += 1; e[1] += ms; e[2] += w
            slow = sorted(by_shape.items(), key=lambda kv: -kv[1][1])[:6]
            roof = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (all linear layers, fwd+dgrad+wgrad)",
                    "achieved": round(ach, 1), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                    "frac": round(ach / pk["tf_sustained"], 4), "traffic": traffic("gemm_tcgen05_kernel"),
                    "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read+write, mean over the family; profiles/traffic.json)",
                    "algorithmic_bytes_per_launch": round(sum(gb) / len(gb)) if gb else None,
                    "peak_source": pk["src"] + " sustained bf16", "launches_per_step": len(gemm),
                    "ms_per_step_in_kernel": round(t, 3), "share_of_step": round(t / ms_step, 3),
                    "timed_on": "the timed step's own path (CUDA events from the C sequencer)",
                    "top_shapes": [{"kind": k, "launches": v[0], "ms": round(v[1], 3),
                                    "tflops": round(v[2] / (v[1] * 1e-3) / 1e12, 1)} for k, v in slow]}
        att = {}
        for name in ("attn_fwd", "attn_bwd"):
            r_ = [(ms, w) for (k, ms, w) in recs if k == name]
            if r_:
                t = sum(ms for ms, _ in r_); f = sum(w for _, w in r_)
                att[name] = {"achieved": round(f / (t * 1e-3) / 1e12, 1), "frac": round(f / (t * 1e-3) / 1e12 / pk["tf_sustained"], 4),
                             "launches_per_step": len(r_), "ms_per_step_in_kernel": round(t, 3), "us_per_launch": round(1e3 * t / len(r_), 1),
                             "traffic": traffic(name + "_kernel")}
        if att:
            t = sum(v["ms_per_step_in_kernel"] for v in att.values())
            f = sum(w for (k, ms, w) in recs if k in ("attn_fwd", "attn_bwd"))
            roof_attn = {"bound": "tensor", "kernel": "attn_fwd_kernel / attn_bwd_kernel (tcgen05 QK^T, PV, dV, dK, dQ; softmax in registers)",
                         "achieved": round(f / (t * 1e-3) / 1e12, 1), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                         "frac": round(f / (t * 1e-3) / 1e12 / pk["tf_sustained"], 4), "flops": "4*B*H*S^2*64 forward, 2.5x backward",
                         "share_of_step": round(t / ms_step, 3),
                         "traffic": traffic("attn_bwd_kernel"), "traffic_unit": "DRAM bytes per launch of attn_bwd_kernel (ncu; profiles/traffic.json)", **att}
        if ln:
            t = sum(ms for ms, _ in ln); by = sum(w for _, w in ln)
            ach = by / (t * 1e-3) / 1e9
            roof_hbm = {"bound": "hbm", "kernel": "ln_fwd_kernel/ln_bwd_kernel (fused residual-add + dropout + LayerNorm)",
                        "achieved": round(ach, 1), "peak": pk["hbm"], "unit": "GB/s", "frac": round(ach / pk["hbm"], 4),
                        "traffic": traffic("ln_kernel"),
                        "traffic_unit": "DRAM bytes per launch (ncu, mean over ln_fwd/ln_bwd launches; profiles/traffic.json)",
                        "algorithmic_bytes_per_launch": round(by / len(ln)), "peak_source": pk["src"],
                        "launches_per_step": len(ln), "ms_per_step_in_kernel": round(t, 3), "share_of_step": round(t / ms_step, 3)}

    # ---- N > 1: a DINO micro-step must leave identical center and gradients on every rank ----
    dp_check = None
    if world > 1:
        dp_check = dino_dp_check(dev, rank, world)

    if rank == 0:
        cpu = torch_gpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = reference_subprocess(args, "cpu")
            cpu = r.get("cpu_baseline") or {"unavailable": r.get("unavailable")}
        if world == 1 and not args.no_torch_baseline:
            r = reference_subprocess(args, "cuda")
            if "value" in r:
                torch_gpu = {"value": r["value"], "unit": "images/s", "ms_per_step": r["ms_per_step"], "batch": r["config"]["batch_per_step"],
                             "what": r["config"]["workload"], "speedup_vs_it": round(value / r["value"], 2)}
            else:
                torch_gpu = {"unavailable": r.get("unavailable")}
        fl = flops_per_image(wl, args.arch)
        body = {"simmim": "model(x) + nn.L1Loss (the drop-in path, simmim_trainer.py:66-67) + GradScaler bwd + AdamW",
                "dino": "model(views, 2) + DINOLoss(center) (dino_trainer.py:86-99) + GradScaler bwd + AdamW + EMA teacher"}[wl]
        line = {
            "metric": metric_name(wl, args.arch, B), "value": round(value, 1), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{ARCH_NAME[args.arch]} {wl} pretraining step: {body}",
                       "batch_per_gpu": B, "global_batch": B * world, "dropout": args.dropout, "parallelism": f"dp{world}",
                       "optimizer": "vit_core.optim.FusedAdamW (torch.optim.VitsslAdamW, config-selectable)" if args.optimizer == "vitssl" else "torch.optim.AdamW",
                       "api": "unmodified-trainer call sequence on the vit_core drop-in modules",
                       "l2_policy": "3 rotating input batches and GBs of per-step activations exceed the 126 MB L2",
                       "mask": "bit-exact reference RNG sequence (B sequential torch.randperm)" if wl == "simmim" else None},
            "model_tflops": round(value * fl["step"] / 1e12, 1),
            "mfu_vs_sustained_bf16": round(value / world * fl["step"] / 1e12 / peaks()["tf_sustained"], 4),
            "e2e": {"value": round(e2e_value, 1), "unit": "images/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": round(e2e_ms, 3), "h2d_gbps": h2d_gbps[0], "per_step": e2e_diag[0], "input": "fp32 images from pinned host memory (ToTensor output, what the reference's loaders yield)"},
            "e2e_u8": {"value": round(e2e8_value, 1), "unit": "images/s", "h2d_bytes_per_step": in_bytes // 4, "d2h_bytes_per_step": 4,
                       "ms_per_step": round(e2e8_ms, 3), "h2d_gbps": h2d_gbps[1], "input": "raw uint8 images from pinned host memory, /255 inside the patch kernels (SURVEY 8(f)3)"},
            "gpu_launches": int(launches), "host_issue_ms_per_step": round(host_issue_ms, 3), "per_step": per_step,
            "settle_steps": settle_steps, "graphs": dict(zip(("captured", "replayed_calls"), lib.graph_stats())),
            "loss": round(final_loss, 5), "clocks": clocks,
        }
        for k, v in (("fused_objective", fused), ("roofline", roof), ("roofline_attn", roof_attn), ("roofline_hbm", roof_hbm),
                     ("cpu_baseline", cpu), ("torch_gpu", torch_gpu), ("dp_check", dp_check)):
            if v:
                line[k] = v
        if comm_exposed is not None:
            line["comm_exposed_ms"] = round(comm_exposed, 3)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def dino_dp_check(dev, rank, world):
    """One data-parallel DINO micro-step (student stack used by two passes, cross-rank center): the
    center and a checksum of every gradient must be bit-identical on all ranks afterwards."""
    import torch.distributed as dist
    from vit_core._backend import dp
    from vit_core.ssl.dino import DINOViT
    from vit_core.ssl.dino.loss import DINOLoss
    torch.manual_seed(7)
    m = DINOViT(num_blocks=2, input_shape=(3, 64, 64), embed_dim=128, patch_size=16, num_heads=2, mlp_dim=256,
                dropout=0.1, output_dim=4096, center_momentum=0.9).to(dev).train()
    dp.attach(m)
    g = torch.Generator().manual_seed(100 + rank)  # different data on every rank
    views = [torch.rand(4, 3, 64, 64, generator=g).to(dev) for _ in range(2)] + [torch.rand(4, 3, 32, 32, generator=g).to(dev) for _ in range(3)]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        t, s = m(views, 2)
        loss = DINOLoss(0.04, 0.1)(t.view(2, 4, -1), s.view(5, 4, -1), m.center)
    loss.backward()
    torch.cuda.synchronize()
    grads = torch.cat([p.grad.reshape(-1).double() for p in m.parameters() if p.grad is not None])
    sig = torch.stack([m.center.double().sum(), m.center.double().abs().sum(), grads.sum(), grads.abs().sum(), grads.square().sum()])
    allsig = [torch.zeros_like(sig) for _ in range(world)]
    dist.all_gather(allsig, sig)
    same_center = all(torch.equal(a[:2], allsig[0][:2]) for a in allsig)
    same_grads = all(torch.equal(a[2:], allsig[0][2:]) for a in allsig)
    return {"center_identical": bool(same_center), "grad_checksum_identical": bool(same_grads), "ranks": world,
            "center_abs_sum": float(sig[1]), "grad_abs_sum": float(sig[3])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="simmim", choices=["simmim", "dino"])
    ap.add_argument("--arch", default="vit_s", choices=list(ARCHS))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: BASELINE's)")
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--optimizer", default="vitssl", choices=["vitssl", "torch"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-baseline", action="store_true")
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"], help="reference arm only")
    ap.add_argument("--ref-batch", type=int, default=0, help="reference arm on CPU: images per step")
    ap.add_argument("--torch-compile", action="store_true", help="reference arm: wrap the model in torch.compile")
    args = ap.parse_args()
    if not args.batch:
        args.batch = default_batch(args.workload, args.arch)
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
