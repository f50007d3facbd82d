"""ORACLE SUPPORT (test / analysis infrastructure, NOT product code): CPU emulation of the B200 path's PRECISION PLAN on
top of the fp32 oracle -- the oracle's algebra with fp16 rounding inserted at exactly the tensors the kernels store in 16
bits, forward (values) and backward (gradients, with the static gradient scale). Used by tests/ as the yardstick for the
end-to-end gradient leg ("as faithful as the storage plan allows") and as an analysis tool: which roundings cost
gradient cosine against the fp32 oracle, what an alternative plan (e.g. an fp32 residual stream) would buy.

    python -m oracle.precision_sim [fixture] [--plans all16,res32,...]

Finding recorded in DESIGN.md 2: with the upstream gradient dL/dCLS of the real loss, the head's train-mode BatchNorm1d
makes sum_b dL/dCLS_b = 0 (BatchNorm backward removes the batch mean), so every late-layer parameter gradient is a
difference of nearly equal per-sample terms; a 3e-4 relative perturbation of the forward values (any 16-bit plan, also the
reference's own fp16 autocast) moves those tensors to cosine 0.98-0.99 in the fp32 oracle itself, while the same plan
under a generic upstream gradient keeps EVERY tensor >= 0.9996.
"""
from __future__ import annotations

import argparse
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import tri_mbt_oracle as O  # noqa: E402

GRAD_SCALE = 4096.0


class _Q(torch.autograd.Function):
    """value -> fp16 in the forward, gradient -> fp16 (scaled) in the backward; either side optional."""

    @staticmethod
    def forward(ctx, x, qf, qb):
        ctx.qb = qb
        return x.half().float() if qf else x

    @staticmethod
    def backward(ctx, g):
        if ctx.qb:
            g = (g * GRAD_SCALE).half().float() / GRAD_SCALE
        return g, None, None


class Plan:
    """which tensors are rounded. keys: stream (X, h: the residual stream), act (xn, qkv, O, hn, a, P), gstream (g_y, g_h,
    g_x), gact (g_a, g_hn, g_qkv, g_xn, dS, P in bwd)."""

    def __init__(self, stream=True, act=True, gstream=True, gact=True):
        self.stream, self.act, self.gstream, self.gact = stream, act, gstream, gact

    def s(self, x):      # residual-stream tensor
        return _Q.apply(x, self.stream, self.gstream)

    def a(self, x):      # non-residual activation
        return _Q.apply(x, self.act, self.gact)


PLANS = {
    "fp32": Plan(False, False, False, False),
    "all16": Plan(True, True, True, True),
    "res32": Plan(False, True, False, True),        # fp32 residual stream, forward and backward
    "fres32": Plan(False, True, True, True),        # fp32 residual stream in the forward only
    "gres32": Plan(True, True, False, True),        # fp32 residual stream in the backward only
    "act32": Plan(True, False, True, False),        # only the residual stream is 16-bit
    "fwd16": Plan(True, True, False, False),        # all gradients fp32
    "bwd16": Plan(False, False, True, True),        # all forward values fp32
}


def mha(sd, prefix, x, mask, n_head, plan):
    B, T, Dm = x.shape
    d = Dm // n_head
    w = torch.cat([sd[f"{prefix}.{p}_proj.linear.weight"] for p in ("query", "key", "value")], 0)
    b = torch.cat([sd[f"{prefix}.{p}_proj.linear.bias"] for p in ("query", "key", "value")], 0)
    qkv = plan.a(F.linear(x, w, b))
    q, k, v = (t.view(B, T, n_head, d).permute(0, 2, 1, 3) for t in qkv.split(Dm, -1))
    s = q @ k.transpose(-1, -2) / math.sqrt(d)
    if mask is not None:
        s = s.masked_fill(mask[:, None, :1, :], -65504.0)
    p = plan.a(torch.softmax(s, -1))
    o = (p @ v).permute(0, 2, 1, 3).reshape(B, T, Dm)
    return plan.a(o)


def layer(sd, prefix, x, mask, n_head, plan):
    xn = plan.a(O.custom_layernorm(x, sd[f"{prefix}.attention_prenorm.gamma"], sd[f"{prefix}.attention_prenorm.beta"]))
    h = plan.s(mha(sd, f"{prefix}.self_attention", xn, mask, n_head, plan) + x)
    hn = plan.a(O.custom_layernorm(h, sd[f"{prefix}.feed_forward_prenorm.gamma"], sd[f"{prefix}.feed_forward_prenorm.beta"]))
    w1 = sd[f"{prefix}.feed_forward.w_1.weight"][:, :, 0]
    w2 = sd[f"{prefix}.feed_forward.w_2.weight"][:, :, 0]
    a = plan.a(torch.relu(F.linear(hn, w1, sd[f"{prefix}.feed_forward.w_1.bias"])))
    return plan.s(F.linear(a, w2, sd[f"{prefix}.feed_forward.w_2.bias"]) + h)


def forward_cls(sd, batch, cfg, plan):
    """CLS output of the fusion encoder [B,256] under `plan` (oracle forward with rounding sites)."""
    x = batch["x"].float()
    B = x.shape[0]
    nb = cfg.bottlenecks_n
    vs = O.umse_vslt_embedding(sd, x)
    txt = plan.a(F.linear(batch["txts"].float().half().float(), sd["txt_embedding.weight"], sd["txt_embedding.bias"]))
    img = plan.a(F.linear(batch["img_feats"].float().half().float(), sd["linear.weight"], sd["linear.bias"]))
    img = img + O.ie_branch(sd, "ie_time", batch["img_time"].float().reshape(-1).unsqueeze(1)).unsqueeze(1) + sd["ie_feat.weight"][18]
    txt = txt + O.ie_branch(sd, "ie_time", batch["txt_time"].float().unsqueeze(1)).unsqueeze(1) + sd["ie_feat.weight"][19]
    if cfg.multiimages == 1:
        img = img.reshape(-1, 147, 256)
    len_v, len_i, len_t = O.stream_lengths(batch["input_lengths"], batch["txt_lengths"], batch["img_time"].float(), cfg)
    P = "fusion_transformer"
    streams = [vs, img, txt]
    enc = []
    for m, s in enumerate(streams):
        y = torch.cat([sd[f"{P}.cls_token_per_modality.{m}"].expand(B, -1, -1), s], 1)
        y = F.layer_norm(y, (256,), sd[f"{P}.layer_norms_in.{m}.weight"], sd[f"{P}.layer_norms_in.{m}.bias"], 1e-5)
        if m == 2:
            y = y + O.positional_encoding(256, y.shape[1])
        enc.append(plan.s(y))
    bott = plan.s(sd[f"{P}.bottlenecks"].expand(B, -1, -1))
    lens = [len_v, len_i, len_t]
    masks = [None if lens[m] is None else O.attn_pad_mask(lens[m] + nb, enc[m].shape[1] + nb, enc[m].shape[1] + nb)
             for m in range(3)]
    missing = batch["missing"].long()
    for l in range(cfg.n_layers):
        last = cfg.vsltonly == 1 and l == cfg.n_layers - 1
        outs, bo = [], []
        for m in range(3):
            y = layer(sd, f"{P}.layer_stacks.{l}.{m}", torch.cat([bott, enc[m]], 1), masks[m], cfg.n_head, plan)
            bo.append(y[:, :nb]); outs.append(y[:, nb:])
            if last:
                break
        enc = outs
        if last:
            break
        st = torch.stack(bo)
        allb = torch.stack([st.mean(0), st[:2].mean(0), torch.stack([st[0], st[2]]).mean(0), st[0]])
        bott = plan.s(allb[missing, torch.arange(B)])
    return enc[0][:, 0]


def grads(sd, batch, cfg, plan, d_cls=None):
    """Parameter gradients under `plan`. d_cls given: gradient of <CLS output, d_cls> (the fusion encoder alone);
    d_cls None: the end-to-end BCE loss through the fp32 classifier head (as the B200 model runs it).
    Returns (grads, cls, logits or None)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point
              and "running" not in k and "positional_encoding" not in k}
    full = dict(sd); full.update(leaves)
    cls = forward_cls(full, batch, cfg, plan)
    logits = None
    if d_cls is not None:
        (cls * d_cls).sum().backward()
    else:
        demo = O.ie_branch(full, "ie_demo", torch.stack([batch["age"].float(), batch["gen"].float()], dim=1))
        logits = O.classifier_head(full, cls, demo, cfg)
        O.loss_fn(logits, batch["y"]).backward()
        logits = logits.detach()
    return {k: v.grad for k, v in leaves.items() if v.grad is not None}, cls.detach(), logits


def cosines(got, ref):
    rows = {}
    for k, r in ref.items():
        if k not in got:
            continue
        a, r = got[k].double().flatten(), r.double().flatten()
        if r.norm() < 1e-6:
            continue
        rows[k] = float(a @ r / (a.norm() * r.norm() + 1e-30))
    ga = torch.cat([got[k].double().flatten() for k in rows]); gr = torch.cat([ref[k].double().flatten() for k in rows])
    return rows, float(ga @ gr / (ga.norm() * gr.norm()))


def main():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import fixture_inputs, fp16_representable, load_fixture
    ap = argparse.ArgumentParser()
    ap.add_argument("fixture", nargs="?", default="tri_nl6_multi_B16_L260")
    ap.add_argument("--plans", default="all16,res32,fres32,gres32,act32,fwd16,bwd16")
    ap.add_argument("--worst", type=int, default=6)
    a = ap.parse_args()
    fx = load_fixture(a.fixture)
    sd, batch, cfg = fixture_inputs(fx)
    sd = fp16_representable(sd)
    torch.manual_seed(0)
    # the oracle's own dL/dCLS of the end-to-end loss
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point
              and "running" not in k and "positional_encoding" not in k}
    full = dict(sd); full.update(leaves)
    logits, aux = O.forward(full, batch, cfg, return_aux=True)
    aux["vslt_out"].retain_grad()
    O.loss_fn(logits, batch["y"]).backward()
    d_cls = aux["vslt_out"].grad[:, 0].detach().clone()
    g32, cls32, _ = grads(sd, batch, cfg, PLANS["fp32"], d_cls)
    for name in a.plans.split(","):
        g, cls, _ = grads(sd, batch, cfg, PLANS[name], d_cls)
        rows, glob = cosines(g, g32)
        vals = np.array(list(rows.values()))
        worst = sorted(rows.items(), key=lambda kv: kv[1])[: a.worst]
        rmse = float((cls - cls32).pow(2).mean().sqrt() / cls32.pow(2).mean().sqrt())
        print(f"{name:7s} cls rel-rmse {rmse:.2e} | grad cos: global {glob:.5f} median {np.median(vals):.5f} "
              f"min {vals.min():.4f} n<0.999 {int((vals < 0.999).sum())}/{len(vals)}")
        for k, v in worst:
            print(f"         {v:.4f} {k}")


if __name__ == "__main__":
    main()
