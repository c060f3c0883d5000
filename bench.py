#!/usr/bin/env python
"""bench.py — headline benchmark: pretraining images/s of ViT-S/16 SimMIM on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload simmim|dino]

One "step" is the reference trainer's loop body (utils/trainers/simmim_trainer.py:63-71):
zero_grad -> autocast(bf16) forward + L1 loss -> GradScaler-scaled backward -> AdamW step.
Workload at N=1 is BASELINE.json configs[1]: ViT-S/16 (D384 L12 H6 F1536), 224x224, mask ratio
0.6, batch 256 per GPU, synthetic torch.rand images, random-init weights, dropout 0.1 (the
reference default, configs/base/model.yaml:7). Under torchrun (N > 1) the batch is sharded per
rank (weak scaling) and gradients are averaged by vit_core._backend.dp.

JSON keys beyond the base contract: `roofline` (flops-weighted throughput of the tcgen05 GEMM
kernel family measured with CUDA events in a separate, instrumented pass of the same step),
`roofline_hbm` (the fused add+LayerNorm kernel, HBM-bound), `cpu_baseline` (the oracle port of the
reference step on the host cores, bounded sample), `e2e` (same step fed from pinned host memory
with the loss read back each step), `gpu_launches`, `clocks`.
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
sys.path.insert(0, os.path.join(ROOT, "vit-ssl_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

VIT_S = dict(embed_dim=384, num_blocks=12, num_heads=6, mlp_dim=1536, patch_size=16)
METRIC = "pretrain images/sec ViT-S/16 SimMIM (224x224, mask 0.6, batch 256/GPU)"


def traffic(family):
    """DRAM bytes per launch (mean over the family's launches of one step) from the committed ncu
    launch list (profiles/traffic.json, written by scripts/summarize_launches.py); None if absent."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))[family]["dram_bytes_per_launch"]
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's SimMIM training step, all host threads
# ------------------------------------------------------------------------------------------
def cpu_reference_step_rate(batch: int, steps: int, warmup: int, seed: int = 42):
    """images/s of the reference algorithm (oracle/vit_ref.py, fp32, eager) on the host cores."""
    from oracle import vit_ref
    torch.manual_seed(seed)
    torch.set_num_threads(os.cpu_count() or 1)
    D, L, H, F_, p = VIT_S["embed_dim"], VIT_S["num_blocks"], VIT_S["num_heads"], VIT_S["mlp_dim"], VIT_S["patch_size"]
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
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        perms = torch.stack([torch.randperm(N) for _ in range(batch)])
        mask = vit_ref.mask_from_perms(perms, N, 0.6)
        pred, tg = vit_ref.simmim_forward(w, x, mask, patch_size=p, num_blocks=L, num_heads=H)
        loss = vit_ref.l1_loss(pred, tg)
        loss.backward()
        opt.step()
        float(loss)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    return batch / (ms / 1e3), ms


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 16
    steps = max(1, min(args.steps, 4))
    warm = 1
    v, ms = cpu_reference_step_rate(batch, steps, warm)
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": "images/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ViT-S/16 SimMIM 224x224 mask 0.6 (reference algorithm, oracle port, CPU eager fp32)",
                   "batch_per_step": batch},
        "cpu_baseline": {"value": round(v, 3), "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} timed steps of batch {batch} (+{warm} warm-up), fwd+L1+bwd+AdamW"},
        "e2e": {"value": round(v, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


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
def simmim_flops_per_image():
    D, L, F_, N, P = 384, 12, 1536, 196, 768
    n_m = int(N * 0.6)
    blk = 8 * N * D * D + 4 * N * N * D + 4 * N * D * F_
    fwd = L * blk + 2 * N * P * D + 2 * n_m * D * P
    step = 3 * (L * blk + 2 * n_m * D * P) + 2 * (2 * N * P * D)  # patch proj has no dgrad
    return fwd, step


def run_gpu_arm(args):
    import torch.distributed as dist
    from vit_core._backend import dp, lib, ops
    from vit_core.ssl.simmim import SimMIMViT

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                                timeout=datetime.timedelta(seconds=180))
    lib.ensure_device()
    dev = torch.device("cuda", local)
    B = args.batch
    torch.manual_seed(42)
    model = SimMIMViT(num_blocks=VIT_S["num_blocks"], input_shape=(3, 224, 224), embed_dim=VIT_S["embed_dim"],
                      patch_size=16, num_heads=VIT_S["num_heads"], mlp_dim=VIT_S["mlp_dim"], dropout=args.dropout,
                      mask_ratio=0.6).to(dev)
    model.train()
    if world > 1:
        dp.attach(model)
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-3, fused=True)
    scaler = torch.amp.GradScaler("cuda")
    torch.manual_seed(1000 + rank)
    n_host = 3
    host = [torch.rand(B, 3, 224, 224).pin_memory() for _ in range(n_host)]
    dev_in = [h.to(dev) for h in host]

    def step(x):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = model.reconstruction_loss(x)
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
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

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(dev_in[i % n_host])
    barrier()

    # ---- timed region 1: device-resident inputs (B*3*224*224*4 B * 3 buffers = 462 MB > L2) ----
    lib.launch_count(reset=True)
    t_host0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        loss = step(dev_in[i % n_host])
    e1.record()
    barrier()
    launches = lib.launch_count()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    t_host1 = time.time()
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    final_loss = float(loss.detach())

    # ---- timed region 2: end to end (pinned host -> device each step, loss read back) ----
    copy_stream = torch.cuda.Stream()
    bufs = [torch.empty_like(dev_in[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[i % 2])
            bufs[i % 2].copy_(host[i % n_host], non_blocking=True)
            ready[i % 2].record(copy_stream)

    for ev in free:
        ev.record()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    prefetch(0)
    for i in range(args.steps):
        if i + 1 < args.steps:
            prefetch(i + 1)
        torch.cuda.current_stream().wait_event(ready[i % 2])
        l = step(bufs[i % 2])
        free[i % 2].record()
        _ = l.item()  # device -> host read of the step's loss
    t1.record()
    barrier()
    ms_e2e = max_over_ranks(t0.elapsed_time(t1))
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    clocks = None
    if rank == 0:
        # clocks under load: samples inside timed region 1 (and the e2e region right after it when
        # region 1 is shorter than a few sampling periods)
        clocks = sampler.window(t_host0, t_host1)
        if (clocks.get("samples") or 0) < 3:
            clocks = sampler.window(t_host0, time.time())
        sampler.stop()

    # ---- instrumented pass: per-kernel-family CUDA-event timing of the same step ----
    # (every rank runs the step — its gradient all-reduce is collective — rank 0 keeps the events)
    roof = roof_hbm = None
    ops.PROFILE = []
    step(dev_in[0])
    barrier()
    recs, ops.PROFILE = ops.PROFILE, None
    if rank == 0:
        pk = peaks()
        gemm = [(a.elapsed_time(b), w) for (k, a, b, w) in recs if k.startswith("gemm")]

        def gemm_bytes(kind):  # "gemm|MxNxK|a_mn=. b_mn=. epi=.": operands read once + result written once
            try:
                _, shp, flags = kind.split("|")
                M_, N_, K_ = (int(v) for v in shp.split("x"))
                fl = dict(kv.split("=") for kv in flags.split())
                out_b = 4 if fl.get("a_mn") == "1" else 2  # weight gradients are fp32
                return 2 * (M_ * K_ + N_ * K_) + M_ * N_ * out_b + (2 * M_ * N_ if fl.get("epi") in ("2", "3") else 0)
            except Exception:
                return None
        gemm_b = [v for v in (gemm_bytes(k) for (k, *_r) in recs if k.startswith("gemm")) if v]
        ln = [(a.elapsed_time(b), w) for (k, a, b, w) in recs if k == "add_layernorm"]
        if gemm:
            t = sum(x for x, _ in gemm); f = sum(w for _, w in gemm)
            ach = f / (t * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (all linear layers, fwd+dgrad+wgrad)",
                    "achieved": round(ach, 1), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                    "frac": round(ach / pk["tf_sustained"], 4), "traffic": traffic("gemm_tcgen05_kernel"),
                    "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read+write, mean over the family; profiles/traffic.json)",
                    "algorithmic_bytes_per_launch": round(sum(gemm_b) / len(gemm_b)) if gemm_b else None,
                    "peak_source": pk["src"] + " sustained bf16",
                    "launches_per_step": len(gemm), "ms_per_step_in_kernel": round(t, 3),
                    "share_of_step": round(t / ms_step, 3)}
        if ln:
            t = sum(x for x, _ in ln); by = sum(w for _, w in ln)
            ach = by / (t * 1e-3) / 1e9
            roof_hbm = {"bound": "hbm", "kernel": "ln_fwd_kernel/ln_bwd_kernel (fused residual-add + LayerNorm)",
                        "achieved": round(ach, 1), "peak": pk["hbm"], "unit": "GB/s", "frac": round(ach / pk["hbm"], 4),
                        "traffic": traffic("ln_kernel"),
                        "traffic_unit": "DRAM bytes per launch (ncu, mean over ln_fwd/ln_bwd launches; profiles/traffic.json)",
                        "algorithmic_bytes_per_launch": round(by / len(ln)),
                        "peak_source": pk["src"], "launches_per_step": len(ln),
                        "ms_per_step_in_kernel": round(t, 3)}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, ms = cpu_reference_step_rate(16, 3, 1)
            cpu = {"value": round(v, 3), "unit": "images/s", "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": "3 timed steps of batch 16 (+1 warm-up) of the same ViT-S/16 SimMIM step, oracle port, eager fp32"}
        fwd_f, step_f = simmim_flops_per_image()
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "ViT-S/16 SimMIM 224x224 mask 0.6 pretraining step (fwd + L1 + bwd + AdamW)",
                       "batch_per_gpu": B, "global_batch": B * world, "dropout": args.dropout, "parallelism": f"dp{world}",
                       "l2_policy": "3 rotating input batches of 154 MB each and >8 GB of per-step activations exceed the 126 MB L2",
                       "mask": "bit-exact reference RNG sequence (B sequential torch.randperm)"},
            "model_tflops": round(value * step_f / 1e12, 1),
            "mfu_vs_sustained_bf16": round(value / world * step_f / 1e12 / peaks()["tf_sustained"], 4),
            "e2e": {"value": round(e2e_value, 1), "unit": "images/s", "h2d_bytes_per_step": B * 3 * 224 * 224 * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": int(launches), "loss": round(final_loss, 5), "clocks": clocks,
        }
        if roof:
            line["roofline"] = roof
        if roof_hbm:
            line["roofline_hbm"] = roof_hbm
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
