"""Kernel-by-kernel parity on a B200, through the C ABI (include/tmp_b200.h via ctypes), against torch fp32 restatements of
the reference ops (cases live in tools/gpu_kernel_check.py so they can also be run stand-alone under gpurun).
Integer work (lengths, masks, feature-id gather) is checked bit-exact; floating point within the tolerance in each case."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import gpu_kernel_check as kc  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", kc.CASES)
def test_kernel_case(case):
    out = kc.run_case(case)
    bad = {k: v for k, v in out.items() if isinstance(v, dict) and not v.get("finite", True)}
    assert out["ok"], (out.get("bad"), {k: v for k, v in out.items() if k != "kv" and not isinstance(v, dict)})
    assert not bad


def test_adamw_matches_torch():
    import torch
    from medical_tri_modal_pilot_b200 import ops
    torch.manual_seed(0)
    n = 1 << 20
    w = torch.randn(n, device="cuda"); w_ref = w.clone().requires_grad_(True)
    m = torch.zeros_like(w); v = torch.zeros_like(w)
    opt = torch.optim.AdamW([w_ref], lr=3e-3, weight_decay=1e-2)
    for t in range(1, 4):
        g = torch.randn(n, device="cuda")
        w_ref.grad = g.clone()
        opt.step()
        ops.adamw_step(w, g, m, v, 3e-3, 0.9, 0.999, 1e-8, 1e-2, t)
    assert (w - w_ref.detach()).abs().max().item() < 1e-5


def test_flat_adamw_and_torch_adamw_agree_on_a_train_step():
    """FlatAdamW (one kernel over the flat buffers + torch AdamW for the head) == torch.optim.AdamW fed the SAME
    gradients, including leaving the dead (grad is None) parameters untouched."""
    import torch
    from golden_util import fixture_inputs, fixture_names, load_fixture
    from test_model_parity_gpu import build_model, run_model
    from medical_tri_modal_pilot_b200.optim import FlatAdamW
    fx = load_fixture(fixture_names()[0])
    sd, batch, cfg = fixture_inputs(fx)
    B = batch["x"].shape[0]
    model = build_model(cfg, sd, B).train()
    opt = FlatAdamW(model, lr=1e-3, weight_decay=1e-2)
    named = {k: p for k, p in model.named_parameters() if not k.startswith("img_encoder.")}
    ref_p = {k: p.detach().clone().requires_grad_(True) for k, p in named.items()}
    ref_opt = torch.optim.AdamW(list(ref_p.values()), lr=1e-3, weight_decay=1e-2)
    for _ in range(2):
        opt.zero_grad()
        out, b = run_model(model, batch)
        torch.nn.BCEWithLogitsLoss()(out.squeeze(), b["y"]).backward()
        for k, p in named.items():
            ref_p[k].grad = None if p.grad is None else p.grad.detach().clone()
        opt.step()
        ref_opt.step()
    n_dead = 0
    for k, p in named.items():
        a, r = p.detach(), ref_p[k].detach()
        assert (a - r).abs().max().item() <= 1e-6 + 1e-5 * r.abs().max().item(), k
        if ref_p[k].grad is None:
            n_dead += 1
            assert torch.equal(a, sd[k].cuda()), k
    assert n_dead >= 20        # last-layer img/txt blocks (--mbt-only-vslt 1), rmse_layer, prelu, unused LayerNorm
