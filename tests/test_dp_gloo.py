"""CPU, world_size 2, gloo: host-side logic of the in-module data parallelism
(vit_core/_backend/dp.py): parameter broadcast, bucketed gradient averaging joined at the end
of backward (so an unmodified trainer sees reduced grads), and the cross-rank sum used by the
DINO center. The kernels are not involved; the N>1 GPU path reuses exactly this code over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "vit-ssl_b200"))
    from vit_core._backend import dp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)  # different init per rank: broadcast must fix it
        model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 8),
                                    torch.nn.LayerNorm(8))
        frozen = model[0].bias
        frozen.requires_grad = False
        v0 = model[2].weight._version
        sync = dp.attach(model, bucket_bytes=1024)  # tiny buckets -> several of them
        # the broadcast writes through the parameter itself: version-keyed caches (bf16 weight shadows
        # cast by a forward that ran before attach) see the change on every rank
        assert model[2].weight._version > v0
        assert len(sync.buckets) > 1
        w0 = [p.detach().clone() for p in model.parameters()]
        gathered = [torch.zeros_like(w0[0]) for _ in range(world)]
        dist.all_gather(gathered, w0[0])
        assert all(torch.equal(g, gathered[0]) for g in gathered), "broadcast failed"

        # per-rank shard of a global batch; mean loss over equal shards
        torch.manual_seed(7)
        X = torch.randn(world * 4, 16)
        Y = torch.randn(world * 4, 8)
        xs, ys = X[rank * 4:(rank + 1) * 4], Y[rank * 4:(rank + 1) * 4]
        for _ in range(2):  # two steps: bucket bookkeeping must reset
            model.zero_grad(set_to_none=True)
            ((model(xs) - ys) ** 2).mean().backward()
            got = [p.grad.clone() for p in model.parameters() if p.requires_grad]
        # single-process reference on the global batch
        ref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 8), torch.nn.LayerNorm(8))
        ref.load_state_dict(model.state_dict())
        ((ref(X) - Y) ** 2).mean().backward()
        want = [p.grad for n, p in ref.named_parameters() if n != "0.bias"]
        for a, b in zip(got, want):
            assert torch.allclose(a, b, atol=1e-6), (a - b).abs().max()
        assert frozen.grad is None

        # gradients averaged INSIDE an autograd node (the encoder stack's chunked backward): the
        # node hands its flat gradient buffer to prereduce() and the hooks must not reduce the
        # same parameters again; a second use of the same parameters in the graph accumulates
        class _Scaled(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x, w, b):
                ctx.save_for_backward(x, w)
                ctx.params = (w, b)
                return x @ w.t() + b

            @staticmethod
            def backward(ctx, gy):
                x, w = ctx.saved_tensors
                flat = torch.cat([(gy.t() @ x).reshape(-1), gy.sum(0)])
                s_ = dp.sync_for(ctx.params)
                assert s_ is sync2
                s_.prereduce(flat, ctx.params)
                s_.join()
                return gy @ w, flat[:w.numel()].view_as(w), flat[w.numel():]

        lin = torch.nn.Linear(16, 8)
        tail = torch.nn.Linear(8, 8)
        both = torch.nn.ModuleList([lin, tail])
        sync2 = dp.attach(both, bucket_bytes=256)
        for _ in range(2):
            both.zero_grad(set_to_none=True)
            y = _Scaled.apply(xs, lin.weight, lin.bias) + _Scaled.apply(2 * xs, lin.weight, lin.bias)
            ((tail(y) - ys) ** 2).mean().backward()
        ref_lin, ref_tail = torch.nn.Linear(16, 8), torch.nn.Linear(8, 8)
        ref_lin.load_state_dict(lin.state_dict()); ref_tail.load_state_dict(tail.state_dict())
        ((ref_tail(ref_lin(X) + ref_lin(2 * X)) - Y) ** 2).mean().backward()
        for a, b in zip([lin.weight.grad, lin.bias.grad, tail.weight.grad, tail.bias.grad],
                        [ref_lin.weight.grad, ref_lin.bias.grad, ref_tail.weight.grad, ref_tail.bias.grad]):
            assert torch.allclose(a, b, atol=1e-5), (a - b).abs().max()
        assert not sync2.prereduced  # cleared by the end-of-backward callback

        # center-style sum
        t = torch.full((5,), float(rank + 1))
        dp.all_reduce_sum_(t)
        assert torch.equal(t, torch.full((5,), float(sum(range(1, world + 1)))))
        assert dp.world_size() == world and dp.is_distributed()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, f"fail: {e!r}"))
    finally:
        dist.destroy_process_group()


def test_grad_sync_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_dp_is_inert_without_process_group():
    import sys
    from vit_core._backend import dp
    m = torch.nn.Linear(4, 4)
    assert dp.attach(m) is None and not dp.is_distributed() and dp.world_size() == 1
    t = torch.ones(3)
    assert torch.equal(dp.all_reduce_sum_(t), torch.ones(3))
