"""world_size-2 CPU (gloo) test of the data-parallel host logic (trainer.GradSync): per-layer ranges of the flat
gradient buffer are averaged across ranks as the backward hands them over, the head bucket follows in finish(), and
initial parameters are broadcast from rank 0. The NCCL path on GPUs uses the same code with a communication stream."""
import os
import socket
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Fake(torch.nn.Module):
    """Stand-in with the attributes GradSync touches: named_parameters(), buffers(), _fused.flat_g / comm_hook."""

    def __init__(self, rank):
        super().__init__()
        torch.manual_seed(100 + rank)                       # different init per rank -> broadcast must fix it
        self.fc_list = torch.nn.Linear(8, 4)                # "head" (not fused) parameters
        self.ie_demo = torch.nn.Linear(2, 8)
        flat_w = torch.randn(64)
        self.ie_time = torch.nn.Linear(1, 16)               # fused-path names: live inside the flat buffer
        self.ie_time.weight.data = flat_w[:16].view(16, 1)
        self.ie_time.bias.data = flat_w[16:32]
        object.__setattr__(self, "_fused", types.SimpleNamespace(flat_g=torch.zeros(64), flat_w=flat_w, comm_hook=None))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from medical_tri_modal_pilot_b200.trainer import GradSync
        model = _Fake(rank)
        sync = GradSync(model, overlap=False)
        assert model.grad_sync is sync and len(sync.head_params) == 4
        # broadcast: every rank now holds rank 0's parameters
        ref = _Fake(0)
        for (n, p), (_, r) in zip(model.named_parameters(), ref.named_parameters()):
            assert torch.equal(p, r), n
        # backward simulation: ranges become final in reverse layer order
        fp = model._fused
        fp.flat_g.copy_(torch.arange(64, dtype=torch.float32) * (rank + 1))
        assert fp.grad_post_scale == 1.0 / world      # the averaging factor rides on FusedPath's un-scaling pass
        for a, b in ((32, 64), (16, 32), (0, 16)):
            fp.flat_g[a:b].mul_(fp.grad_post_scale)    # what FusedPath._range_done does before handing the range over
            fp.comm_hook(a, b)
        for k, p in enumerate(sync.head_params):
            p.grad = torch.full_like(p, float((rank + 1) * (k + 1)))
        sync.finish()
        mean_scale = sum(r + 1 for r in range(world)) / world
        assert torch.allclose(fp.flat_g, torch.arange(64, dtype=torch.float32) * mean_scale)
        for k, p in enumerate(sync.head_params):
            assert torch.allclose(p.grad, torch.full_like(p, mean_scale * (k + 1)))
        assert sync.n_collectives == 4
        # early head bucket: the gradients that exist when dL/dCLS reaches the fused backward go out then (FusedPath calls
        # head_hook), the rest in finish() -- every gradient is averaged exactly once
        for k, p in enumerate(sync.head_params):
            p.grad = torch.full_like(p, float((rank + 1) * (k + 1))) if k < 2 else None
        fp.head_hook()
        assert sync.n_collectives == 5
        for k, p in enumerate(sync.head_params):
            if k >= 2:
                p.grad = torch.full_like(p, float((rank + 1) * (k + 1)))
        sync.finish()
        assert sync.n_collectives == 6
        for k, p in enumerate(sync.head_params):
            assert torch.allclose(p.grad, torch.full_like(p, mean_scale * (k + 1))), k
        fp.head_hook()                       # next step, all four ready early: one bucket, nothing left for finish()
        sync.finish()
        assert sync.n_collectives == 7
        for k, p in enumerate(sync.head_params):     # identical on every rank already: the average changes nothing
            assert torch.allclose(p.grad, torch.full_like(p, mean_scale * (k + 1))), k
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradsync_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_gradsync_single_rank_is_a_noop():
    port = _free_port()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        from medical_tri_modal_pilot_b200.trainer import GradSync
        model = _Fake(0)
        sync = GradSync(model)
        model._fused.flat_g.fill_(2.0)
        model._fused.comm_hook(0, 64)
        sync.finish()
        assert torch.equal(model._fused.flat_g, torch.full((64,), 2.0)) and sync.n_collectives == 0
    finally:
        dist.destroy_process_group()
