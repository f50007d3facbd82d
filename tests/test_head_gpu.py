"""Classifier head in training mode: the three-launch sm_100a head (csrc/head.cu, head.py) against the
stock PyTorch modules it stands in for (reference tri_mbt_vsltcls.py:176-177, :248-255): logits, BatchNorm running
statistics, the gradient that flows back into the fused path and every parameter gradient; bit-reproducible (no atomics)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _head_modules(seed):
    from medical_tri_modal_pilot_b200.config import make_args
    from builder.models import get_model
    args = make_args(transformer_num_layers=2, multiimages=1, mbt_only_vslt=1, input_types="vslt_img_txt", imgtxt_time=1,
                     dropout=0.1, batch_size=8, img_pretrain="No")
    args.device = torch.device("cuda")
    torch.manual_seed(seed)
    model = get_model(args)(args).cuda().train()
    g = torch.Generator(device="cuda").manual_seed(seed)
    with torch.no_grad():          # non-trivial affine parameters / running statistics
        for n in ("layer_norms_after_concat", "ie_demo.1", "fc_list.1"):
            mod = model.get_submodule(n)
            mod.weight.copy_(1 + 0.3 * torch.randn(256, device="cuda", generator=g))
            mod.bias.copy_(0.2 * torch.randn(256, device="cuda", generator=g))
        model.fc_list[1].running_mean.copy_(0.1 * torch.randn(256, device="cuda", generator=g))
        model.fc_list[1].running_var.copy_(0.5 + torch.rand(256, device="cuda", generator=g))
    return model


def _torch_head(model, cls_out, age, gen):
    demo = model.ie_demo(torch.cat([age.unsqueeze(1), gen.unsqueeze(1)], dim=1).float())
    return model.fc_list(torch.cat([model.layer_norms_after_concat(cls_out), demo], dim=1))


@pytest.mark.parametrize("B", [4, 16, 64, 100])
def test_fused_head_matches_the_pytorch_modules(B):
    from medical_tri_modal_pilot_b200 import head, ops
    model = _head_modules(B)
    ref = copy.deepcopy(model)
    g = torch.Generator(device="cuda").manual_seed(100 + B)
    cls = (torch.randn(B, 256, device="cuda", generator=g) * 1.5 + 0.3)
    age = torch.rand(B, device="cuda", generator=g)
    gen = (torch.rand(B, device="cuda", generator=g) > 0.5).float()
    dlog = torch.randn(B, 1, device="cuda", generator=g)

    c1 = cls.clone().requires_grad_(True)
    assert head.usable(model, c1)
    out = head.fused_head(model, c1, age, gen)
    out.backward(dlog)
    c2 = cls.clone().requires_grad_(True)
    out_ref = _torch_head(ref, c2, age, gen)
    out_ref.backward(dlog)

    # relative to the reference's largest element, with an absolute floor (B = 2, the smallest batch BatchNorm accepts, is
    # left out: its output is +-1 whatever the input, every gradient upstream of it is rounding noise times 1/sqrt(var + eps))
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp_min(2e-2)).item()
    assert out.shape == (B, 1) and rel(out, out_ref) < 1e-5, rel(out, out_ref)
    assert rel(c1.grad, c2.grad) < 2e-4, rel(c1.grad, c2.grad)
    pr = dict(ref.named_parameters())
    for n, p in model.named_parameters():
        if n in ops.HEAD_PARAM_ORDER:
            assert p.grad is not None and p.grad.shape == pr[n].grad.shape, n
            if n == "fc_list.0.bias":          # mathematically zero behind a train-mode BatchNorm: rounding noise on both sides
                assert p.grad.abs().max() < 1e-5 * max(1.0, dlog.abs().max().item())
                continue
            assert rel(p.grad, pr[n].grad) < 2e-4, (n, rel(p.grad, pr[n].grad))
    bn, bn_ref = model.fc_list[1], ref.fc_list[1]
    assert rel(bn.running_mean, bn_ref.running_mean) < 1e-5 and rel(bn.running_var, bn_ref.running_var) < 1e-5
    assert int(bn.num_batches_tracked) == int(bn_ref.num_batches_tracked) == 1

    # bit-reproducible: a second model with the same state gives identical logits and gradients
    m2 = _head_modules(B)
    c3 = cls.clone().requires_grad_(True)
    out2 = head.fused_head(m2, c3, age, gen)
    out2.backward(dlog)
    assert torch.equal(out2, out) and torch.equal(c3.grad, c1.grad)
    assert torch.equal(m2.fc_list[0].weight.grad, model.fc_list[0].weight.grad)


def test_fused_head_is_bypassed_in_eval_mode_and_by_the_switch(monkeypatch):
    from medical_tri_modal_pilot_b200 import head
    model = _head_modules(3)
    cls = torch.randn(8, 256, device="cuda")
    assert head.usable(model, cls)
    assert not head.usable(model.eval(), cls)
    model.train()
    assert head.usable(model, cls[:2]) and not head.usable(model, cls[:1])      # BatchNorm needs two rows
    out2 = head.fused_head(model, cls[:2].clone(), torch.rand(2, device="cuda"), torch.ones(2, device="cuda"))
    assert out2.shape == (2, 1) and bool(torch.isfinite(out2).all())
    monkeypatch.setenv("TMP_B200_FUSED_HEAD", "0")
    assert not head.usable(model, cls)
