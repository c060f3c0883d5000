"""Loss scalars whose value travels to the host as soon as the kernel that produced it has run.

The reference's trainers read `loss.item()` once per step AFTER `scaler.step` / `scaler.update`
(simmim_trainer.py:69-76, dino_trainer.py:100-109). On a plain tensor that read is a device-to-host
copy enqueued behind the whole backward and the optimizer, so the host cannot start issuing the next
step until the GPU has drained — with ~1 ms of Python between that point and the next step's first
encoder kernel, the GPU idles ~0.9 ms of a 15 ms step (profiles/r2_gaps.md). The loss modules of this
package therefore return a `PrefetchedScalar`: an ordinary autograd tensor whose 4-byte copy into
pinned memory is enqueued right behind the loss kernel; `.item()` waits for THAT copy only and
returns the same value. Everything else (`backward`, arithmetic, `float()`, printing) is the plain
tensor's behaviour."""
import weakref

import torch

_RING = 32
_slots = None      # pinned [RING] fp32
_owners = [None] * _RING
_cursor = 0


class PrefetchedScalar(torch.Tensor):
    __torch_function__ = torch._C._disabled_torch_function_impl  # results of ops are plain tensors

    def _fetch(self):
        st = self.__dict__.get("_vitssl_prefetch")
        if st is None:
            return None
        if len(st) == 3:                       # (slot, event, ring) -> read once, remember the float
            slot, ev, ring = st
            ev.synchronize()
            self.__dict__["_vitssl_prefetch"] = (float(ring[slot]),)
            if _owners[slot] is not None and _owners[slot]() is self:
                _owners[slot] = None
        return self.__dict__["_vitssl_prefetch"][0]

    def item(self):
        v = self._fetch()
        return torch.Tensor.item(self) if v is None else v


def prefetch_scalar(t):
    """`t`: a one-element CUDA tensor (fp32). Returns `t` as a PrefetchedScalar (same storage, same
    autograd node) with its host copy in flight on the current stream."""
    global _slots, _cursor
    if not (torch.is_tensor(t) and t.is_cuda and t.numel() == 1 and t.dtype == torch.float32):
        return t
    if torch.cuda.is_current_stream_capturing():
        return t
    if _slots is None:
        _slots = torch.empty(_RING, dtype=torch.float32).pin_memory()
    slot = _cursor
    _cursor = (_cursor + 1) % _RING
    prev = _owners[slot]() if _owners[slot] is not None else None
    if prev is not None:
        prev._fetch()                          # a still-unread loss from 32 losses ago: read it before its slot is reused
    with torch.no_grad():
        _slots[slot:slot + 1].copy_(t.detach().reshape(1), non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    out = t.as_subclass(PrefetchedScalar)
    out.__dict__["_vitssl_prefetch"] = (slot, ev, _slots)
    _owners[slot] = weakref.ref(out)
    return out
