"""Data-parallel exchange on real GPUs: two NCCL ranks, each running the fused forward/backward on HALF of a
batch with GradSync's overlapped range-by-range all-reduce, must end with the same averaged gradient as ONE rank running the
whole batch (the loss is a sum over samples, so sum_of_halves / world == full / 2). Needs >= 2 GPUs (skipped on the 1-GPU
box; run with `gpurun --gpus 2`). The CPU-side bucket logic is covered by tests/test_gradsync_gloo.py."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(B):
    from golden_util import fixture_inputs, fp16_representable, load_fixture
    from test_model_parity_gpu import build_model
    fx = load_fixture("tri_nl2_multi_B32_L40")
    sd, batch, cfg = fixture_inputs(fx)
    sd = fp16_representable(sd)
    model = build_model(cfg, sd, B).train()
    return model, batch


def _fused_grads(model, b, R):
    model.zero_grad(set_to_none=True)
    cls = model._fused(b["x"], b["input_lengths"], b["txts"], b["txt_lengths"], b["img_feats"], b["img_time"], b["txt_time"],
                       b["missing"])
    cls.backward(R)
    sync = getattr(model, "grad_sync", None)
    if sync is not None:
        sync.finish()
    torch.cuda.synchronize()
    fp = model._fused
    return fp.flat_g[: fp.live_end()].detach().clone()


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from builder.trainer import GradSync
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    model, batch = _build(16)
    sync = GradSync(model)
    B = batch["x"].shape[0]
    half = slice(rank * B // 2, (rank + 1) * B // 2)
    himg = slice(rank * 3 * B // 2, (rank + 1) * 3 * B // 2)
    b = {k: (v[himg] if k == "img_feats" else v[half]).cuda().contiguous() for k, v in batch.items()}
    R = (torch.randn(B, 256, generator=torch.Generator().manual_seed(9)) * 0.02)[half].cuda()
    g = _fused_grads(model, b, R)
    assert sync.n_collectives >= 3
    torch.save(g.cpu(), os.path.join(out_dir, f"g{rank}.pt"))
    dist.barrier()
    sync.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_allreduce_equals_single_rank_full_batch(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g0, g1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    assert torch.equal(g0, g1)                                   # every rank holds the same averaged gradient
    model, batch = _build(32)
    b = {k: v.cuda() for k, v in batch.items()}
    R = (torch.randn(32, 256, generator=torch.Generator().manual_seed(9)) * 0.02).cuda()
    g_full = _fused_grads(model, b, R).cpu() / 2
    cos = torch.nn.functional.cosine_similarity(g0.double(), g_full.double(), dim=0).item()
    rel = ((g0 - g_full).norm() / g_full.norm()).item()
    assert cos > 0.9999 and rel < 1e-2, (cos, rel)


def _head_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from builder.trainer import GradSync
    from test_model_parity_gpu import run_model
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    model, batch = _build(16)
    sync = GradSync(model)
    B = batch["x"].shape[0]
    half = slice(rank * B // 2, (rank + 1) * B // 2)
    himg = slice(rank * 3 * B // 2, (rank + 1) * 3 * B // 2)
    b = {k: (v[himg] if k == "img_feats" else v[half]).contiguous() for k, v in batch.items()}
    r = torch.randn(B, 1, generator=torch.Generator().manual_seed(5))[half].cuda()
    res = {}
    for mode in (True, False):            # the three-launch head (head.py) / the stock PyTorch modules
        model.fused_head = mode
        model.zero_grad(set_to_none=True)
        out, _ = run_model(model, b)
        (out * r).sum().backward()
        sync.finish()
        torch.cuda.synchronize()
        res[mode] = {n: p.grad.detach().cpu().clone() for n, p in model.named_parameters()
                     if p.grad is not None and not n.startswith("fusion_transformer.")}
        res[mode]["flat"] = model._fused.flat_g[: model._fused.live_end()].detach().cpu().clone()
    torch.save(res, os.path.join(out_dir, f"h{rank}.pt"))
    dist.barrier()
    sync.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_head_gradients_fused_head_equals_stock_modules(tmp_path):
    """Whole model under GradSync on two ranks: the gradients of the classifier-head parameters (averaged early, during the
    fused backward) are the same on both ranks and the same whether the head ran as csrc/head.cu or as the stock modules."""
    import torch.multiprocessing as mp
    port = 31500 + os.getpid() % 2000
    mp.spawn(_head_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    h0, h1 = torch.load(tmp_path / "h0.pt"), torch.load(tmp_path / "h1.pt")
    names = [n for n in h0[True] if n != "flat"]
    assert any(n.startswith("fc_list.0") for n in names) and any(n.startswith("ie_demo") for n in names), names
    for mode in (True, False):
        for n in h0[mode]:
            assert torch.equal(h0[mode][n], h1[mode][n]), (mode, n)          # averaged: identical on every rank
    # element-wise, relative to the tensor's largest element with a floor tied to the largest head gradient: BatchNorm in
    # training mode makes the loss invariant to layer_norms_after_concat.bias and fc_list.0.bias (a constant shift of every
    # row is removed with the batch mean), so those two gradients are rounding noise around an exact zero in BOTH modes
    scale = max(h0[False][n].abs().max().item() for n in names)
    for n in h0[True]:
        a, c = h0[True][n].double().flatten(), h0[False][n].double().flatten()
        err = (a - c).abs().max().item() / max(c.abs().max().item(), 1e-3 * scale)
        assert err < 5e-3, (n, err)
        if c.abs().max().item() > 1e-3 * scale:
            cos = torch.nn.functional.cosine_similarity(a, c, dim=0).item()
            assert cos > 0.9995, (n, cos)

