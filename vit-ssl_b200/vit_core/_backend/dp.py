"""Data parallelism that lives inside the modules (the reference has none: SURVEY §5/§8(e)).

One process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch). Images shard across ranks,
so the only exchanges are
  1. the gradient all-reduce (mean) — bucketed, launched from post-accumulate-grad hooks on a
     side stream as buckets become ready and joined in an autograd-engine callback, so the
     reference's unmodified trainer (`scaler.step(optimizer)` right after `backward()`) sees
     reduced gradients;
  2. the DINO center: all-reduce of the per-rank column sums of the teacher logits;
  3. a parameter/buffer broadcast from rank 0 when a model is first used.
The wrapper-free design keeps `model.center`, `model.momentum_update_teacher` etc. reachable
(dino_trainer.py:99,105), which DistributedDataParallel's wrapper would hide.

The logic is device-agnostic (CPU tensors + gloo work too) so it is unit-tested without GPUs.
"""
from __future__ import annotations

import weakref
from typing import Iterable, List

import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors
from torch.autograd import Variable

DEFAULT_BUCKET_BYTES = 32 << 20
# bench.py sets this to a list to measure EXPOSED communication: (event, event) pairs recorded on the
# compute stream around every wait for the all-reduce stream (the span the compute stream is stalled)
WAIT_TRACE = None


def _wait_for(comm_stream) -> None:
    cur = torch.cuda.current_stream()
    if WAIT_TRACE is None:
        cur.wait_stream(comm_stream)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cur)
    cur.wait_stream(comm_stream)
    e1.record(cur)
    WAIT_TRACE.append((e0, e1))

_SYNCS: "weakref.WeakSet" = weakref.WeakSet()  # live GradSync objects (sync_for looks a parameter up here)


def world_size() -> int:
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def is_distributed() -> bool:
    return world_size() > 1


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def broadcast_module_(module: torch.nn.Module, src: int = 0) -> None:
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not is_distributed():
        return
    tensors = list(module.parameters()) + list(module.buffers())
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault((t.dtype, t.device), []).append(t)
    with torch.no_grad():
        for group in by_dtype.values():
            flat = _flatten_dense_tensors([t.detach() for t in group])
            dist.broadcast(flat, src=src)
            for t, f in zip(group, _unflatten_dense_tensors(flat, group)):
                # an in-place copy on the tensor itself (not `.data`) bumps its version counter, which
                # is what invalidates bf16 weight shadows cast by an earlier forward on this rank
                t.copy_(f)


class GradSync:
    """Bucketed, overlapped gradient averaging for one module."""

    def __init__(self, module: torch.nn.Module, bucket_bytes: int = DEFAULT_BUCKET_BYTES):
        self.world = world_size()
        self.bucket_bytes = bucket_bytes
        params = [p for p in module.parameters() if p.requires_grad]
        # backward produces gradients roughly in reverse registration order
        params = list(reversed(params))
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, size = [], 0
        for p in params:
            nbytes = p.numel() * p.element_size()
            if cur and size + nbytes > bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {}
        for bi, b in enumerate(self.buckets):
            for p in b:
                self.bucket_of[id(p)] = bi
        self.ready = [0] * len(self.buckets)
        self.launched = [False] * len(self.buckets)
        self.callback_queued = False
        self.comm_stream = None
        self.handles = []
        self.reduced_bytes = 0
        self.prereduced = set()  # ids of parameters whose gradient was averaged inside backward already
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook()) for p in params]
        _SYNCS.add(self)
        if self.world > 1 and any(p.is_cuda for p in params):
            # gradient buffers that a collective touches on another stream are released at
            # timing-dependent moments, so buffer addresses do not repeat from step to step and the
            # stack calls' CUDA-graph replay (csrc/encoder.cu) would capture a new graph every time
            # (measured at 2 GPUs: 70-80 captures in 45 steps, 15.0 instead of 14.7 ms/step)
            from . import lib as _lib
            _lib.graph_enable(False)

    def _make_hook(self):
        ref = weakref.ref(self)

        def hook(param):
            self_ = ref()
            if self_ is not None:
                self_._on_grad(param)

        return hook

    # -- hook path -------------------------------------------------------------------------
    def _on_grad(self, param) -> None:
        if self.world <= 1:
            return
        bi = self.bucket_of[id(param)]
        self.ready[bi] += 1
        if not self.callback_queued:
            Variable._execution_engine.queue_callback(self._finalize)
            self.callback_queued = True
        if self.ready[bi] == len(self.buckets[bi]):
            self._launch(bi)

    def prereduce(self, flat: torch.Tensor, params: Iterable[torch.Tensor]) -> None:
        """Average `flat` (a contiguous buffer holding the gradient contributions that an autograd
        node is about to return for `params`) across ranks, in place, asynchronously on the
        communication stream — called from inside that node's backward so the all-reduce overlaps
        the rest of the backward pass (the encoder stack is ONE node: without this its gradients,
        98 % of the model, would only become visible to the hooks when the whole stack is done).
        The hooks then skip these parameters. Linear, so it composes with gradient accumulation
        from several nodes (DINO's global and local passes share parameters)."""
        if self.world <= 1:
            return
        self.reduced_bytes += flat.numel() * flat.element_size()
        if not self.callback_queued:
            Variable._execution_engine.queue_callback(self._finalize)
            self.callback_queued = True
        if flat.is_cuda:
            if self.comm_stream is None:
                self.comm_stream = torch.cuda.Stream()
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                flat.div_(self.world)
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                flat.record_stream(self.comm_stream)
        else:
            flat.div_(self.world)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        for p in params:
            self.prereduced.add(id(p))

    def join(self) -> None:
        """Make the current stream wait for the reductions issued so far (the autograd engine is
        about to read / accumulate the tensors `prereduce` is averaging in place)."""
        if self.comm_stream is not None:
            _wait_for(self.comm_stream)

    def _launch(self, bi: int) -> None:
        grads = [p.grad for p in self.buckets[bi] if p.grad is not None and id(p) not in self.prereduced]
        self.launched[bi] = True
        if not grads:
            return
        self.reduce_async(grads)

    def reduce_async(self, grads: Iterable[torch.Tensor]) -> None:
        """Average `grads` across ranks in place; completes by the next `_finalize`."""
        grads = list(grads)
        if self.world <= 1 or not grads:
            return
        self.reduced_bytes += sum(g.numel() * g.element_size() for g in grads)
        if grads[0].is_cuda:
            if self.comm_stream is None:
                self.comm_stream = torch.cuda.Stream()
            cur = torch.cuda.current_stream()
            self.comm_stream.wait_stream(cur)
            with torch.cuda.stream(self.comm_stream):
                flat = _flatten_dense_tensors(grads)
                flat.div_(self.world)
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                for g, f in zip(grads, _unflatten_dense_tensors(flat, grads)):
                    g.copy_(f)
                    g.record_stream(self.comm_stream)
        else:
            flat = _flatten_dense_tensors(grads)
            flat.div_(self.world)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            for g, f in zip(grads, _unflatten_dense_tensors(flat, grads)):
                g.copy_(f)

    def _finalize(self) -> None:
        # params whose bucket never filled (unused this step) still get reduced so ranks agree
        for bi in range(len(self.buckets)):
            if not self.launched[bi] and self.ready[bi] > 0:
                self._launch(bi)
        if self.comm_stream is not None:
            _wait_for(self.comm_stream)
        self.ready = [0] * len(self.buckets)
        self.launched = [False] * len(self.buckets)
        self.callback_queued = False
        self.prereduced.clear()

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


def sync_for(params: Iterable[torch.Tensor]):
    """The live GradSync that owns (all of) the trainable tensors in `params`, or None."""
    if not is_distributed():
        return None
    ids = [id(p) for p in params if getattr(p, "requires_grad", False)]
    if not ids:
        return None
    for s in list(_SYNCS):
        if s._hooks and all(i in s.bucket_of for i in ids):
            return s
    return None


def attach(module: torch.nn.Module, bucket_bytes: int = DEFAULT_BUCKET_BYTES, broadcast: bool = True):
    """Idempotently enable data-parallel training for `module` (no-op with world size 1)."""
    if not is_distributed():
        return None
    sync = module.__dict__.get("_vitssl_grad_sync")
    if sync is not None:
        return sync
    if broadcast:
        broadcast_module_(module)
    sync = GradSync(module, bucket_bytes)
    object.__setattr__(module, "_vitssl_grad_sync", sync)
    return sync


def maybe_attach(module: torch.nn.Module) -> None:
    """Called from the top-level models' forward: under torchrun the drop-in becomes data
    parallel without any change to the reference's train.py / trainers."""
    if is_distributed() and module.training and "_vitssl_grad_sync" not in module.__dict__:
        attach(module)
