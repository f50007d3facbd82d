"""Diagnostic: per-layer dL/dX (gradient wrt every encoder-layer input, per stream) of the fused path vs the oracle's
autograd, to localise where a gradient mismatch enters. Run under gpurun."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from golden_util import fixture_inputs, load_fixture
from test_model_parity_gpu import build_model
from oracle import tri_mbt_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "tri_nl3_multi_B16_L150"
fx = load_fixture(name)
sd, batch, cfg = fixture_inputs(fx)
B = batch["x"].shape[0]
rec = {}
orig = O.encoder_layer
def hooked(sd_, prefix, x, mask, n_head):
    x = x.clone(); x.retain_grad()
    y = orig(sd_, prefix, x, mask, n_head)
    l, m = (int(v) for v in prefix.split("layer_stacks.")[1].split(".")[:2])
    rec[(l, m)] = x
    return y
O.encoder_layer = hooked
leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k and "positional" not in k}
full = dict(sd); full.update(leaves)
logits, aux = O.forward(full, batch, cfg, return_aux=True)
aux["vslt_out"].retain_grad()
O.loss_fn(logits, batch["y"]).backward()
d_cls = aux["vslt_out"].grad[:, 0].clone()
model = build_model(cfg, sd, B).train()
fp = model._fused
fp.debug_trace = {}
b = {k: v.cuda() for k, v in batch.items()}
cls = fp(b["x"], b["input_lengths"], b["txts"], b["txt_lengths"], model.encode_images(b["img_feats"], None), b["img_time"], b["txt_time"], b["missing"])
cls.backward(d_cls.cuda())
kv = fp.ctx["kv_len"].cpu()
for (tag, l, s), g in sorted(fp.debug_trace.items(), key=lambda kv_: (-kv_[0][1], kv_[0][2])):
    ref = rec[(l, s)].grad
    T = ref.shape[1]
    live = (torch.arange(T)[None, :] < kv[s][:, None])
    miss = batch["missing"]
    present = torch.ones(B, dtype=torch.bool) if s == 0 else ((miss != 2) & (miss != 3) if s == 1 else (miss != 1) & (miss != 3))
    def cos(a, r):
        return float((a * r).sum() / (a.norm() * r.norm() + 1e-30))
    a, r = g, ref
    sel = live & present[:, None]
    out = [f"dX l{l} s{s}: all cos {cos(a, r):.5f}"]
    out.append(f"live-rows cos {cos(a[sel], r[sel]):.5f} |a| {a[sel].norm():.3e} |r| {r[sel].norm():.3e}")
    out.append(f"dead rows |a| {a[~sel].norm():.2e} |r| {r[~sel].norm():.2e}")
    out.append(f"bott cos {cos(a[:, :4][present], r[:, :4][present]):.5f} cls cos {cos(a[:, 4][present], r[:, 4][present]):.5f} tok cos {cos((a[:, 5:] * sel[:, 5:, None]), (r[:, 5:] * sel[:, 5:, None])):.5f}")
    # column sums over live rows (bias-like)
    out.append(f"colsum cos {cos((a * sel[..., None]).sum((0, 1)), (r * sel[..., None]).sum((0, 1))):.5f}")
    # per-sample cos, worst
    pc = [(cos(a[i][sel[i]], r[i][sel[i]]), i, int(kv[s][i])) for i in range(B) if present[i]]
    out.append("worst samples " + str(sorted(pc)[:3]))
    print(" | ".join(out))
