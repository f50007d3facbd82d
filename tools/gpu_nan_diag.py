"""Diagnostic: fp16 gradient headroom of the fused backward at a fixture shape. Prints, per upstream-gradient scale, which
parameter gradients are non-finite and the largest |scaled gradient| per (layer, stream) residual gradient tensor."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import fixture_inputs, fp16_representable, load_fixture  # noqa: E402
from test_model_parity_gpu import build_model  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "tri_nl6_multi_B64_L1000"
fx = load_fixture(name)
sd, batch, cfg = fixture_inputs(fx)
sd = fp16_representable(sd)
B = batch["x"].shape[0]
model = build_model(cfg, sd, B).train()
b = {k: v.cuda() for k, v in batch.items()}
fp = model._fused
for scale in (0.02, 0.002):
    gen = torch.Generator().manual_seed(1234)
    R = torch.randn(B, 256, generator=gen) * scale
    model.zero_grad(set_to_none=True)
    fp.debug_trace = {}
    cls = fp(b["x"], b["input_lengths"], b["txts"], b["txt_lengths"], model.encode_images(b["img_feats"], None),
             b["img_time"], b["txt_time"], b["missing"])
    print("cls finite", bool(torch.isfinite(cls).all()), "max", cls.abs().max().item())
    cls.backward(R.cuda())
    bad = [k for k, p in model.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
    print(f"scale {scale}: non-finite grads in {len(bad)} tensors", bad[:12])
    for key in sorted(fp.debug_trace, key=lambda k: (-k[1], k[2])):
        t = fp.debug_trace[key]
        fin = torch.isfinite(t)
        mx = (t[fin].abs().max().item() * fp.grad_scale) if fin.any() else float("nan")
        print(f"   dX layer {key[1]} stream {key[2]}: max |scaled| {mx:10.1f}  non-finite {int((~fin).sum())}")
    for s in range(3):
        st = fp.ws[s]
        for nm in ("g_a", "g_hn", "g_h", "g_qkv", "g_xn"):
            t = st[nm].float()
            fin = torch.isfinite(t)
            print(f"   stream {s} {nm}: max |scaled| {t[fin].abs().max().item():10.1f} non-finite {int((~fin).sum())}")
