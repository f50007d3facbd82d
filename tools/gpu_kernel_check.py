"""Kernel-by-kernel numerics check on a real B200 (run through gpurun). Each case runs in its own subprocess with a
timeout so that a trapping / hanging kernel cannot take the whole run (or the box) down.

    python tools/gpu_kernel_check.py            # run all cases, write gpurun_out/kernel_check.json
    python tools/gpu_kernel_check.py --case gemm
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ["lengths", "umse", "layernorm", "prologue", "mix_colsum", "gemm", "wgrad", "attn_fwd", "attn_bwd", "precise"]


import torch as _t
ACT, GRD = _t.float16, _t.float16


def _err(a, b):
    import torch
    a = a.float(); b = b.float()
    d = (a - b).abs()
    denom = b.abs().max().clamp_min(1e-12)
    return {"max_abs": d.max().item(), "rel_to_max": (d.max() / denom).item(),
            "rmse_rel": (d.pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-12)).item(),
            "finite": bool(torch.isfinite(a).all().item())}


def case_lengths():
    import torch
    from medical_tri_modal_pilot_b200 import ops
    dev = "cuda"
    B = 7
    il = torch.tensor([40, 10, 25, 3, 300, 1, 0], device=dev)
    tl = torch.tensor([0, 5, 126, 128, 1, 64, 0], device=dev)
    it = torch.full((B, 3), 10.0, device=dev)
    filled = [0, 1, 2, 3, 1, 2, 0]
    for b, k in enumerate(filled):
        it[b, :k] = -float(b) - 0.5
    missing = torch.tensor([0, 1, 2, 3, 0, 0, 3], device=dev)
    Tv, Ti, Tt = 305, 152, 133
    kv = ops.build_lengths(il, tl, it, 3, 1, missing, 0, Tv, Ti, Tt).cpu()
    exp_v = [min(x + 5, Tv) for x in il.tolist()]
    exp_i = [49 * k + 5 for k in filled]
    exp_t = [4 if x == 0 else min(x + 3, 129) + 4 for x in tl.tolist()]
    ok = kv[0].tolist() == exp_v and kv[1].tolist() == exp_i and kv[2].tolist() == exp_t
    kv2 = ops.build_lengths(il, tl, it, 3, 1, missing, 1, Tv, Ti, Tt).cpu()
    ok2 = kv2[1].tolist() == [0 if m in (2, 3) else e for m, e in zip(missing.tolist(), exp_i)] and \
        kv2[2].tolist() == [0 if m in (1, 3) else e for m, e in zip(missing.tolist(), exp_t)]
    m = ops.materialize_mask(kv[0].to(dev).contiguous(), Tv).cpu()
    ar = torch.arange(Tv)
    expm = (ar[None, None, :] >= kv[0][:, None, None]).expand(B, Tv, Tv)
    return {"ok": bool(ok and ok2 and torch.equal(m, expm)), "kv": kv.tolist()}


def _branch_ref(s, w, b, g, be):
    import torch
    z = s[..., None] * w + b
    return torch.relu(torch.nn.functional.layer_norm(z, (256,), g, be, 1e-5))


def case_umse():
    import torch
    from medical_tri_modal_pilot_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    n = 5000
    x = torch.empty(n, 3, device=dev)
    x[:, 0] = -torch.rand(n, device=dev) * 24
    x[:, 1] = torch.rand(n, device=dev)
    x[:, 2] = torch.randint(0, 18, (n,), device=dev).float()
    mk = lambda: [torch.randn(256, device=dev), torch.randn(256, device=dev) * 0.5, 1 + 0.1 * torch.randn(256, device=dev),
                  0.1 * torch.randn(256, device=dev)]
    val4, tim4 = mk(), mk()
    W = torch.randn(20, 256, device=dev)
    E = ops.umse_embed(x, val4, tim4, W, torch.float32)
    ref = _branch_ref(x[:, 1], *val4) + _branch_ref(x[:, 0], *tim4) + W[x[:, 2].int().long()]
    out = {"fp32": _err(E, ref)}
    Eb = ops.umse_embed(x, val4, tim4, W, ACT)
    out["fp16"] = _err(Eb, ref)
    # gather bit-exactness: kill both LN-ReLU branches (gamma = beta = 0 -> relu(0) = 0)
    z4 = [val4[0], val4[1], torch.zeros(256, device=dev), torch.zeros(256, device=dev)]
    Eg = ops.umse_embed(x, z4, z4, W, torch.float32)
    out["gather_bit_exact"] = bool(torch.equal(Eg, W[x[:, 2].int().long()]))
    out["ok"] = out["fp32"]["max_abs"] < 2e-5 and out["fp16"]["rel_to_max"] < 2e-3 and out["gather_bit_exact"]
    return out


def _ln_ref(z, g, b):
    mean = z.mean(-1, keepdim=True)
    std = z.std(-1, keepdim=True)
    return g * (z - mean) / (std + 1e-6) + b


def case_layernorm():
    import torch
    from medical_tri_modal_pilot_b200 import ops
    torch.manual_seed(1)
    dev = "cuda"
    rows = 3001
    x = torch.randn(rows, 256, device=dev).half()
    o = torch.randn(rows, 256, device=dev).half()
    g = (1 + 0.1 * torch.randn(256, device=dev)); b = 0.1 * torch.randn(256, device=dev)
    y = torch.empty_like(x)
    ops.layernorm_fwd(x, g, b, y)
    res = {"fwd": _err(y, _ln_ref(x.float(), g, b))}
    h = torch.empty_like(x); y2 = torch.empty_like(x)
    ops.layernorm_fwd(x, g, b, y2, add=o, sum_out=h)
    href = (x.float() + o.float())
    res["add_sum"] = _err(h, href)
    res["add_fwd"] = _err(y2, _ln_ref(h.float(), g, b))
    # backward
    xf = x.float().requires_grad_(True); gp = g.clone().requires_grad_(True); bp = b.clone().requires_grad_(True)
    dy = torch.randn(rows, 256, device=dev).to(GRD)
    dres = torch.randn(rows, 256, device=dev).to(GRD)
    _ln_ref(xf, gp, bp).backward(dy.float())
    dx = torch.empty_like(dy); dg = torch.zeros(256, device=dev); db = torch.zeros(256, device=dev)
    ops.layernorm_bwd(dy, x, dres, g, dx, dg, db)
    res["bwd_dx"] = _err(dx, xf.grad + dres.float())
    res["bwd_dg"] = _err(dg, gp.grad)
    res["bwd_db"] = _err(db, bp.grad)
    dxd = torch.empty_like(dy); dx2 = torch.empty_like(dy)
    dg.zero_(); db.zero_()
    ops.layernorm_bwd(dy, x, dres, g, dx2, dg, db, dx_drop=dxd, drop_p=0.1, seed=7, salt=3)
    keep = (dxd != 0).float().mean().item()
    res["drop_keep_frac"] = keep
    kept = dxd != 0
    res["drop_scale"] = _err(dxd[kept].float(), (dx2.float() / 0.9)[kept])
    res["ok"] = all(res[k]["rel_to_max"] < 2e-2 for k in ("fwd", "add_fwd", "bwd_dx", "bwd_dg", "bwd_db")) and \
        abs(keep - 0.9) < 0.01
    return res


def case_prologue():
    import torch
    from medical_tri_modal_pilot_b200 import ops
    torch.manual_seed(2)
    dev = "cuda"
    res = {}
    mk = lambda: [torch.randn(256, device=dev), torch.randn(256, device=dev) * 0.5, 1 + 0.1 * torch.randn(256, device=dev),
                  0.1 * torch.randn(256, device=dev)]
    val4, tim4 = mk(), mk()
    W = torch.randn(20, 256, device=dev)
    cls = torch.randn(256, device=dev); bott = torch.randn(4, 256, device=dev)
    lg = 1 + 0.1 * torch.randn(256, device=dev); lb = 0.1 * torch.randn(256, device=dev)
    pe = torch.randn(200, 256, device=dev)
    # the last two: more rows than 32 per warp of a full grid (the backward walks a warp's row range in chunks of 32) and
    # an img stream that fills every SM
    for kind, B, n, drop in ((0, 5, 37, 0.0), (1, 5, 147, 0.0), (0, 7, 300, 0.25), (1, 9, 147, 0.25), (0, 40, 1000, 0.1),
                             (1, 64, 147, 0.1)):
        T = n + 5
        tag = (f"{kind}" if drop == 0.0 else f"{kind}_drop") + ("_big" if B * n > 5000 else "")
        leaves = [t.clone().requires_grad_(True) for t in (*val4, *tim4, W, cls, bott, lg, lb)]
        v4, t4, Wl, cl, bo, lgl, lbl = leaves[0:4], leaves[4:8], leaves[8], leaves[9], leaves[10], leaves[11], leaves[12]
        if kind == 0:
            x = torch.empty(B, n, 3, device=dev)
            x[..., 0] = -torch.rand(B, n, device=dev) * 24
            x[..., 1] = torch.rand(B, n, device=dev)
            x[..., 2] = torch.randint(0, 18, (B, n), device=dev).float()
            E = _branch_ref(x[..., 1], *v4) + _branch_ref(x[..., 0], *t4) + Wl[x[..., 2].long()]
            proj = times = None; n_slots = 0; feat = 0; use_pe = None
        else:
            x = None
            proj = torch.randn(B * n, 256, device=dev).half()
            projf = proj.float().requires_grad_(True)
            times = -torch.rand(B, 3, device=dev) * 24
            n_slots = 3; feat = 18; use_pe = pe[: n + 1].contiguous()
            te = _branch_ref(times, *t4)  # [B,3,256]
            E = projf.view(B, 3, 49, 256) + te[:, :, None, :] + Wl[feat]
            E = E.reshape(B, n, 256)
        seq = torch.cat([cl.expand(B, 1, 256), E], 1)
        y = torch.nn.functional.layer_norm(seq, (256,), lgl, lbl, 1e-5)
        if use_pe is not None:
            y = y + use_pe
        X0 = torch.empty(B, T, 256, device=dev, dtype=ACT)
        ops.stream_prologue_fwd(kind, B, n, x, val4 if kind == 0 else None, proj, times, n_slots, feat, tim4, W, cls,
                                bott, lg, lb, use_pe, drop, 11, 5, X0)
        if drop > 0:
            # the mask is a stateless hash: read it off the forward output, apply it to the reference, and require the
            # backward (same seed / salt) to use the identical mask
            keep = (X0[:, 4:] != 0).float()
            res[f"keep_frac{tag}"] = {"rel_to_max": abs(keep.mean().item() - (1 - drop)), "finite": True}
            y = y * keep / (1 - drop)
        ref = torch.cat([bo.expand(B, 4, 256), y], 1)
        res[f"fwd{tag}"] = _err(X0, ref)
        dX0 = torch.randn(B, T, 256, device=dev).to(GRD)
        ref.backward(dX0.float())
        g_val = torch.zeros(4, 256, device=dev); g_tim = torch.zeros(4, 256, device=dev)
        g_feat = torch.zeros(20, 256, device=dev); g_cls = torch.zeros(256, device=dev)
        g_bott = torch.zeros(4, 256, device=dev); g_ln = torch.zeros(2, 256, device=dev)
        dproj = torch.empty(B * n, 256, device=dev, dtype=GRD) if kind == 1 else None
        ops.stream_prologue_bwd(kind, B, n, x, val4 if kind == 0 else None, proj, times, n_slots, feat, tim4, W, cls,
                                bott, lg, lb, use_pe, drop, 11, 5, dX0, g_val if kind == 0 else None, g_tim, g_feat,
                                g_cls, g_bott, g_ln, dproj)
        if kind == 0:
            res[f"g_val{tag}"] = _err(g_val, torch.stack([l.grad for l in v4]))
        else:
            res[f"dproj{tag}"] = _err(dproj, projf.grad)
        res[f"g_tim{tag}"] = _err(g_tim, torch.stack([l.grad for l in t4]))
        res[f"g_feat{tag}"] = _err(g_feat, Wl.grad)
        res[f"g_cls{tag}"] = _err(g_cls, cl.grad)
        res[f"g_bott{tag}"] = _err(g_bott, bo.grad)
        res[f"g_ln{tag}"] = _err(g_ln, torch.stack([lgl.grad, lbl.grad]))
    res["ok"] = all(v["rel_to_max"] < 2e-2 for k, v in res.items() if isinstance(v, dict))
    return res


def case_mix_colsum():
    import torch
    from medical_tri_modal_pilot_b200 import ops
    torch.manual_seed(3)
    dev = "cuda"
    B = 9
    Ys = [torch.randn(B, T, 256, device=dev).half() for T in (45, 152, 133)]
    missing = torch.tensor([0, 1, 2, 3, 0, 1, 2, 3, 0], device=dev)
    bo = torch.stack([y[:, :4].float() for y in Ys])
    tri = bo.mean(0); vi = bo[:2].mean(0); vt = (bo[0] + bo[2]) / 2
    ref = torch.stack([tri, vi, vt, bo[0]])[missing, torch.arange(B, device=dev)]
    Yc = [y.clone() for y in Ys]
    ops.bottleneck_mix_fwd(*Yc, missing)
    res = {"mix_fwd": max(_err(y[:, :4], ref)["rel_to_max"] for y in Yc),
           "mix_rest_untouched": all(torch.equal(a[:, 4:], b[:, 4:]) for a, b in zip(Yc, Ys))}
    # bwd
    w = torch.tensor([[1 / 3, 1 / 3, 1 / 3], [.5, .5, 0], [.5, 0, .5], [1, 0, 0]], device=dev)[missing]  # [B,3]
    dY = [torch.randn(B, T, 256, device=dev).to(GRD) for T in (45, 152, 133)]
    g = sum(d[:, :4].float() for d in dY)
    dYc = [d.clone() for d in dY]
    ops.bottleneck_mix_bwd(*dYc, 1, missing)
    res["mix_bwd"] = max(_err(dYc[m][:, :4], g * w[:, m, None, None])["rel_to_max"] for m in range(3))
    dYc = [d.clone() for d in dY]
    ops.bottleneck_mix_bwd(*dYc, 0, missing)
    res["mix_bwd_vonly"] = max(_err(dYc[m][:, :4], dY[0][:, :4].float() * w[:, m, None, None])["rel_to_max"]
                               for m in range(3))
    # dropped copies of rows 0..3: must equal dropout_apply of the mixed tensor (same mask: seed, per-stream salt, element
    # index in the stream's [B*T, 256] matrix); rows >= 4 of the dropped tensors are not touched
    dYc = [d.clone() for d in dY]
    drp = [torch.full_like(d, 7.0) for d in dY]
    salts = (10, 14, 18)
    ops.bottleneck_mix_bwd(*dYc, 1, missing, dropped=tuple(drp), drop_p=0.1, seed=77, salts=salts)
    ok = True
    for m in range(3):
        full = torch.empty_like(dYc[m])
        ops.dropout_apply(dYc[m], full, 0.1, 77, salts[m])
        a_, f_ = drp[m][:, :4].float(), full[:, :4].float()      # the kernel drops the fp32 value, dropout_apply the rounded one
        ok = ok and torch.equal(a_ == 0, f_ == 0) and bool(((a_ - f_).abs() <= 2e-3 * f_.abs() + 1e-6).all()) \
            and bool((drp[m][:, 4:] == 7.0).all())
    res["mix_bwd_dropped_rows"] = ok
    for N in (256, 768, 1024):
        M = 4097
        dy = torch.randn(M, N, device=dev).to(GRD)
        out = torch.zeros(N, device=dev)
        ops.colsum(dy, out)
        res[f"colsum{N}"] = _err(out, dy.float().sum(0))["rel_to_max"]
    # dropout_apply statistics + determinism
    a = torch.ones(1 << 20, device=dev).to(GRD); o1 = torch.empty_like(a); o2 = torch.empty_like(a)
    ops.dropout_apply(a, o1, 0.1, 11, 5); ops.dropout_apply(a, o2, 0.1, 11, 5)
    res["drop_keep"] = (o1 != 0).float().mean().item(); res["drop_det"] = bool(torch.equal(o1, o2))
    res["ok"] = res["mix_bwd_dropped_rows"] and res["mix_fwd"] < 1e-2 and res["mix_bwd"] < 1e-2 and res["mix_bwd_vonly"] < 1e-2 and \
        res["mix_rest_untouched"] and all(res[f"colsum{N}"] < 1e-3 for N in (256, 768, 1024)) and \
        abs(res["drop_keep"] - 0.9) < 5e-3 and res["drop_det"]
    return res


def case_gemm():
    import torch
    from medical_tri_modal_pilot_b200 import ops
    torch.manual_seed(4)
    dev = "cuda"
    res = {}
    # (A dtype, B dtype, out dtype): the product uses fp16 everywhere; bf16 x bf16 stays supported by the kernel
    for tag, (da, db_, do) in {"fwd": (ACT, ACT, ACT), "bf16": (_t.bfloat16,) * 3}.items():
        # the last four shapes take the weight-stationary variant (K <= 256 / 512, >= 2 tiles per SM), incl. an M tail
        for (M, N, K) in [(128, 128, 64), (300, 256, 256), (4096, 768, 256), (1000, 1024, 256), (1001, 256, 1024),
                          (20000, 256, 768), (40001, 768, 256), (38000, 1024, 256), (40001, 384, 128),
                          (50000, 128, 384)]:
            A = torch.randn(M, K, device=dev).to(da)
            Bw = (torch.randn(N, K, device=dev) / K ** 0.5).to(db_)
            bias = torch.randn(N, device=dev)
            ref = A.float() @ Bw.float().t()
            out = torch.empty(M, N, device=dev, dtype=do)
            ops.gemm(A, Bw, out=out)
            res[f"{tag}_plain_{M}x{N}x{K}"] = _err(out, ref)
            ops.gemm(A, Bw, out=out, bias=bias, relu=True)
            res[f"{tag}_bias_relu_{M}x{N}x{K}"] = _err(out, torch.relu(ref + bias))
            resid = torch.randn(M, N, device=dev).to(ACT)
            of = torch.empty(M, N, device=dev)
            ops.gemm(A, Bw, out=out, out_f32=of, bias=bias, residual=resid)
            res[f"{tag}_bias_res_{M}x{N}x{K}"] = _err(out, ref + bias + resid.float())
            res[f"{tag}_f32out_{M}x{N}x{K}"] = _err(of, ref + bias + resid.float())
            gate = torch.randn(M, N, device=dev).to(ACT)
            ops.gemm(A, Bw, out=out, gate=gate, alpha=0.5)
            res[f"{tag}_gate_{M}x{N}x{K}"] = _err(out, 0.5 * ref * (gate.float() > 0))
            if tag != "fwd":
                continue
            # the compile-time specialised epilogues (gemm_tc05.cu TMP_EPI_MODES) that the cases above do not reach:
            # bias only; bias + GELU; bias + folded dropout + fp16 residual (checked against the GENERIC epilogue, which an
            # additional fp32 output selects: same dropout pattern, same values)
            ops.gemm(A, Bw, out=out, bias=bias)
            res[f"spec_bias_{M}x{N}x{K}"] = _err(out, ref + bias)
            ops.gemm(A, Bw, out=out, bias=bias, relu=2)
            res[f"spec_bias_gelu_{M}x{N}x{K}"] = _err(out, torch.nn.functional.gelu(ref + bias))
            o_s = torch.empty(M, N, device=dev, dtype=do); o_g = torch.empty_like(o_s)
            ops.gemm(A, Bw, out=o_s, bias=bias, residual=resid, drop_p=0.1, seed=11, salt=3)
            ops.gemm(A, Bw, out=o_g, out_f32=of, bias=bias, residual=resid, drop_p=0.1, seed=11, salt=3)
            res[f"spec_bias_drop_res_{M}x{N}x{K}"] = _err(o_s, o_g.float())
            kept = (of - resid.float()) != 0
            res[f"spec_bias_drop_res_vs_math_{M}x{N}x{K}"] = _err(o_s[kept], ((ref + bias) / 0.9 + resid.float())[kept])
    M, N, K = 2048, 1024, 256
    A = torch.randn(M, K, device=dev).half(); Bw = (torch.randn(N, K, device=dev) / 16).half()
    o0 = torch.empty(M, N, device=dev, dtype=ACT); o1 = torch.empty_like(o0); o2 = torch.empty_like(o0)
    ops.gemm(A, Bw, out=o0)
    ops.gemm(A, Bw, out=o1, drop_p=0.1, seed=3, salt=9); ops.gemm(A, Bw, out=o2, drop_p=0.1, seed=3, salt=9)
    kept = o1 != 0
    res["drop_keep"] = kept.float().mean().item()
    res["drop_det"] = bool(torch.equal(o1, o2))
    res["drop_scale"] = _err(o1[kept], (o0.float() / 0.9)[kept])
    # 1-bit ReLU/dropout gate: the FFN1 epilogue writes (result > 0) as a bit mask, the FFN2 input gradient consumes it
    mask_ok = True
    for (M, N, K) in [(300, 1024, 256), (40001, 1024, 256), (64, 1024, 256)]:
        A = torch.randn(M, K, device=dev).half(); Bw = (torch.randn(N, K, device=dev) / 16).half()
        bias = torch.randn(N, device=dev)
        a = torch.empty(M, N, device=dev, dtype=ACT)
        am = torch.full((M, N // 32), -1, device=dev, dtype=torch.int32)
        ops.gemm(A, Bw, out=a, bias=bias, relu=True, drop_p=0.1, seed=5, salt=2, mask_out=am)
        bits = ((am.unsqueeze(-1) >> torch.arange(32, device=dev)) & 1).reshape(M, N).bool()
        # the mask is taken from the fp32 result: it may be set where the fp16-rounded activation is exactly 0
        # (a positive value below the smallest fp16 subnormal) -- never the other way round
        mism = bits != (a > 0)
        res[f"mask_bits_{M}"] = {"rel_to_max": 0.0 if (mism.float().mean().item() < 1e-5 and bool((bits | ~mism).all())) else 1.0,
                                 "finite": True, "mismatch": int(mism.sum().item())}
        G = torch.randn(M, 256, device=dev).half(); Wt = (torch.randn(N, 256, device=dev) / 16).half()
        g1 = torch.empty(M, N, device=dev, dtype=ACT); g2 = torch.empty_like(g1)
        ops.gemm(G, Wt, out=g1, gate=a, alpha=1.0 / 0.9)
        ops.gemm(G, Wt, out=g2, gate=am, alpha=1.0 / 0.9)
        dif = (g1 != g2)
        res[f"mask_gate_{M}"] = {"rel_to_max": 0.0 if bool((dif == mism).all()) or dif.float().mean().item() < 1e-5 else 1.0,
                                 "finite": bool(torch.isfinite(g2).all()), "differs": int(dif.sum().item())}
    res["mask_gate_ok"] = mask_ok
    tol = lambda k: 2.5e-3 if k.startswith("fwd") else 1.5e-2
    res["bad"] = [k for k, v in res.items() if isinstance(v, dict) and not (v["rel_to_max"] < tol(k) and v["finite"])]
    res["ok"] = not res["bad"] and abs(res["drop_keep"] - 0.9) < 5e-3 and res["drop_det"] and mask_ok
    return res


def case_wgrad():
    import torch
    from medical_tri_modal_pilot_b200 import ops
    torch.manual_seed(5)
    dev = "cuda"
    res = {}
    for (M, N, K) in [(64, 128, 128), (1000, 256, 1024), (5000, 768, 256), (4097, 1024, 256), (333, 256, 768)]:
        dY = torch.randn(M, N, device=dev).to(GRD)
        X = torch.randn(M, K, device=dev).half()
        dW = torch.zeros(N, K, device=dev)
        ops.gemm_wgrad(dY, X, dW)
        res[f"{M}x{N}x{K}"] = _err(dW, dY.float().t() @ X.float())
        # fused bias gradient (column sums of dY from the staged smem tiles), accumulating into a non-zero buffer
        dW2 = torch.zeros(N, K, device=dev)
        db = torch.ones(N, device=dev)
        ops.gemm_wgrad(dY, X, dW2, dbias=db)
        res[f"{M}x{N}x{K}_dbias"] = _err(db, 1.0 + dY.float().sum(0))
        res[f"{M}x{N}x{K}_dW_with_dbias"] = _err(dW2, dY.float().t() @ X.float())
    # the bench shape, repeated: 53 k-blocks per split put the last dY block into ring stage 0, which the write-out reuses
    # (a summing warp racing another warp's staging box gave sporadic non-finite bias gradients before the named barrier)
    for (M, N, K) in [(64320, 1024, 256), (64320, 256, 1024), (64320, 768, 256)]:
        dY = torch.randn(M, N, device=dev).to(GRD)
        X = torch.randn(M, K, device=dev).half()
        ref_b = dY.float().sum(0)
        worst = None
        for rep in range(12):
            dW2 = torch.zeros(N, K, device=dev)
            db = torch.zeros(N, device=dev)
            ops.gemm_wgrad(dY, X, dW2, dbias=db)
            e = _err(db, ref_b)
            if worst is None or not e["finite"] or e["rel_to_max"] > worst["rel_to_max"]:
                worst = e
        res[f"{M}x{N}x{K}_dbias_x12"] = worst
    res["ok"] = all(v["rel_to_max"] < 5e-3 and v["finite"] for v in res.values() if isinstance(v, dict))
    return res


def case_precise():
    """fp32 mode building blocks (csrc/precise.cu): bf16x3-split GEMM / wgrad against fp64 matmul, fp32 CUDA-core attention
    forward / backward against torch fp64 autograd, fp32-storage LayerNorm."""
    import torch
    from medical_tri_modal_pilot_b200 import ops
    torch.manual_seed(11)
    dev = "cuda"
    res = {}
    for (M, N, K) in [(300, 256, 256), (1000, 1024, 256), (517, 256, 1024), (256, 768, 768)]:
        A = torch.randn(M, K, device=dev) * (10.0 ** torch.randint(-3, 3, (M, 1), device=dev).float())   # wide dynamic range
        W = torch.randn(N, K, device=dev) / K ** 0.5
        bias = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev)
        ops.gemm(A, W, out=out, bias=bias)
        ref = (A.double() @ W.double().t() + bias.double())
        res[f"gemm_{M}x{N}x{K}"] = _err(out.double() / A.abs().max(1, keepdim=True).values.double(),
                                        ref / A.abs().max(1, keepdim=True).values.double())
        # epilogue with fp32 gate + residual
        gate = torch.randn(M, N, device=dev)
        resid = torch.randn(M, N, device=dev)
        out2 = torch.empty(M, N, device=dev)
        A1 = torch.randn(M, K, device=dev)
        ops.gemm(A1, W, out=out2, gate=gate, residual=resid, alpha=1.5)
        ref2 = (1.5 * (A1.double() @ W.double().t())) * (gate > 0).double() + resid.double()
        res[f"gemm_gate_res_{M}x{N}x{K}"] = _err(out2.double(), ref2)
        dY = torch.randn(M, N, device=dev) * 1e-4      # gradient-like magnitudes, no scaling
        dW = torch.zeros(N, K, device=dev); db = torch.zeros(N, device=dev)
        ops.gemm_wgrad(dY, A1, dW, dbias=db)
        res[f"wgrad_{M}x{N}x{K}"] = _err(dW.double(), dY.double().t() @ A1.double())
        res[f"dbias_{M}x{N}x{K}"] = _err(db.double(), dY.double().sum(0))
    tol = {k: 2e-5 for k in res}
    # attention
    for (B, T, lens) in [(2, 128, [128, 77]), (3, 300, [300, 150, 4]), (2, 1005, [1005, 600]), (3, 54, [54, 0, 7])]:
        qkv = (torch.randn(B * T, 768, device=dev) * 1.5)
        kv = torch.tensor(lens, device=dev, dtype=torch.int32)
        O = torch.full((B * T, 256), 7.0, device=dev)
        lse = torch.zeros(B, 4, ops.lse_len(T), device=dev)
        ops.attn_fwd(qkv, kv, B, T, O, lse)
        q64 = qkv.double().requires_grad_(True)
        live = (torch.arange(T, device=dev)[None, :] < kv[:, None]).reshape(B * T, 1)
        ref = _attn_ref64(q64, kv, B, T) * live
        res[f"attn_fwd_T{T}"] = _err(O.double(), ref.detach())
        dO = torch.randn(B * T, 256, device=dev) * live
        ref.backward(dO.double())
        dq = torch.full((B * T, 768), 3.0, device=dev)
        delta = torch.zeros_like(lse)
        ops.attn_bwd(qkv, O, dO, kv, B, T, lse, delta, None, dq)
        res[f"attn_bwd_T{T}"] = _err(dq.double(), q64.grad)
        tol[f"attn_fwd_T{T}"] = tol[f"attn_bwd_T{T}"] = 2e-5
    # fp32-storage LayerNorm forward / backward
    rows = 777
    x = torch.randn(rows, 256, device=dev); o = torch.randn(rows, 256, device=dev)
    g = 1 + 0.1 * torch.randn(256, device=dev); bb = 0.1 * torch.randn(256, device=dev)
    h = torch.empty_like(x); y = torch.empty_like(x)
    ops.layernorm_fwd(x, g, bb, y, add=o, sum_out=h)
    res["ln_f32_fwd"] = _err(y, _ln_ref(x + o, g, bb)); tol["ln_f32_fwd"] = 1e-5
    xf = (x + o).requires_grad_(True)
    dy = torch.randn(rows, 256, device=dev); dres = torch.randn(rows, 256, device=dev)
    _ln_ref(xf, g, bb).backward(dy)
    dx = torch.empty_like(x); dg = torch.zeros(256, device=dev); dbt = torch.zeros(256, device=dev)
    ops.layernorm_bwd(dy, h, dres, g, dx, dg, dbt)
    res["ln_f32_bwd"] = _err(dx, xf.grad + dres); tol["ln_f32_bwd"] = 1e-5
    res["ok"] = all(res[k]["rel_to_max"] < tol[k] and res[k]["finite"] for k in tol)
    return res


def _attn_ref64(qkv, kv_len, B, T):
    import torch
    q, k, v = qkv.view(B, T, 3, 4, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) / 8.0
    mask = torch.arange(T, device=qkv.device)[None, None, None, :] >= kv_len[:, None, None, None]
    s = s.masked_fill(mask, -65504.0)
    p = torch.softmax(s, -1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B * T, 256)


def _attn_ref(qkv, kv_len, B, T):
    import torch
    q, k, v = qkv.float().view(B, T, 3, 4, 64).permute(2, 0, 3, 1, 4)  # [B,H,T,64]
    s = (q @ k.transpose(-1, -2)) / 8.0
    mask = torch.arange(T, device=qkv.device)[None, None, None, :] >= kv_len[:, None, None, None]
    s = s.masked_fill(mask, -65504.0)
    p = torch.softmax(s, -1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B, T, 256)


def case_attn_fwd():
    import torch
    from medical_tri_modal_pilot_b200 import ops
    torch.manual_seed(6)
    dev = "cuda"
    res = {}
    many = [300, 0, 150, 4, 299, 129, 128, 1] * 5   # 40 samples x 4 heads x 3 tiles = 480 work items (> SM count)
    for (B, T, lens) in [(2, 128, [128, 77]), (3, 300, [300, 150, 4]), (2, 1005, [1005, 600]), (2, 54, [54, 54]),
                         (40, 300, many), (3, 2005, [2005, 0, 1203]), (2, 2048, [2048, 777]), (2, 4096, [4096, 2500])]:
        qkv = (torch.randn(B * T, 768, device=dev) * 1.5).half()
        kv = torch.tensor(lens, device=dev, dtype=torch.int32)
        O = torch.full((B * T, 256), 7.0, device=dev, dtype=ACT)
        lse = torch.zeros(B, 4, ops.lse_len(T), device=dev)
        ops.attn_fwd(qkv, kv, B, T, O, lse)
        ref = _attn_ref(qkv, kv, B, T)
        live = (torch.arange(T, device=dev)[None, :] < kv[:, None])  # only live query rows are defined by parity
        e = _err(O.view(B, T, 256)[live], ref[live])
        res[f"B{B}_T{T}"] = e
    res["ok"] = all(v["rel_to_max"] < 3e-3 and v["finite"] for v in res.values() if isinstance(v, dict))
    return res


def case_attn_bwd():
    import torch
    from medical_tri_modal_pilot_b200 import ops
    torch.manual_seed(7)
    dev = "cuda"
    res = {}
    # last case: 480 (key tile, head, sample) items on 148 persistent CTAs -- several items per CTA, de-selected
    # samples (kv_len 0), dead and partial key tiles in between
    many = [300, 0, 150, 4, 299, 129, 128, 1] * 5
    # T = 2005 (BASELINE config 4) and S = 2048 / 4096 (config 5) with ragged and zero kv_len are checked too
    for (B, T, lens) in [(2, 128, [128, 77]), (3, 300, [300, 150, 4]), (2, 1005, [1005, 600]), (40, 300, many),
                         (3, 2005, [2005, 0, 1203]), (2, 2048, [2048, 777]), (2, 4096, [4096, 2500])]:
        qkv = (torch.randn(B * T, 768, device=dev)).half()
        kv = torch.tensor(lens, device=dev, dtype=torch.int32)
        live = (torch.arange(T, device=dev)[None, :] < kv[:, None])
        qf = qkv.float().requires_grad_(True)
        ref = _attn_ref(qf, kv, B, T)
        dO = torch.randn(B, T, 256, device=dev).to(GRD)
        dO = (dO * live[..., None]).contiguous()  # padding query rows carry exactly zero gradient
        ref.backward(dO.float())
        O = torch.empty(B * T, 256, device=dev, dtype=ACT)
        Tl = ops.lse_len(T)
        lse = torch.zeros(B, 4, Tl, device=dev)
        ops.attn_fwd(qkv, kv, B, T, O, lse)
        delta = torch.empty(B, 4, Tl, device=dev); dq_acc = torch.empty(B * T, 256, device=dev)
        dQKV = torch.full((B * T, 768), 3.0, device=dev, dtype=GRD)
        ops.attn_bwd(qkv, O, dO.view(B * T, 256), kv, B, T, lse, delta, dq_acc, dQKV)
        g = qf.grad.view(B * T, 768)
        res[f"dQ_B{B}_T{T}"] = _err(dQKV[:, :256], g[:, :256])
        res[f"dK_B{B}_T{T}"] = _err(dQKV[:, 256:512], g[:, 256:512])
        res[f"dV_B{B}_T{T}"] = _err(dQKV[:, 512:], g[:, 512:])
        # fused protocol (what the training step runs): the LayerNorm backward that produces dO also writes delta and
        # zeroes the dQ columns; the attention kernel reduce-adds its dQ tiles in fp16 in place
        gam = torch.ones(256, device=dev)
        xin = torch.randn(B * T, 256, device=dev).half()
        dy = torch.zeros(B * T, 256, device=dev, dtype=GRD)       # LN'(0) = 0  ->  dx = dres = dO
        dx = torch.empty(B * T, 256, device=dev, dtype=GRD)
        dgm = torch.zeros(256, device=dev); dbt = torch.zeros(256, device=dev)
        delta2 = torch.zeros(B, 4, Tl, device=dev)
        dQKV2 = torch.full((B * T, 768), 3.0, device=dev, dtype=GRD)
        ops.layernorm_bwd_attn(dy, xin, dO.view(B * T, 256), gam, dx, dgm, dbt, O, T, delta2, dQKV2)
        dref = (dO.view(B, T, 4, 64).float() * O.view(B, T, 4, 64).float()).sum(-1).permute(0, 2, 1)
        res[f"fused_delta_B{B}_T{T}"] = _err(delta2[:, :, :T], dref)
        res[f"fused_dx_B{B}_T{T}"] = _err(dx, dO.view(B * T, 256))
        zero_ok = bool((dQKV2[:, :256] == 0).all().item()) and bool((dQKV2[:, 256:] == 3.0).all().item())
        ops.attn_bwd(qkv, O, dx, kv, B, T, lse, delta2, None, dQKV2)
        res[f"fused_dQ_B{B}_T{T}"] = _err(dQKV2[:, :256], g[:, :256])
        res[f"fused_dK_B{B}_T{T}"] = _err(dQKV2[:, 256:512], g[:, 256:512])
        res[f"fused_dV_B{B}_T{T}"] = _err(dQKV2[:, 512:], g[:, 512:])
        res[f"fused_zero_B{B}_T{T}"] = {"rel_to_max": 0.0 if zero_ok else 1.0, "finite": True}
        # single-query form (CLS-only last layer): dO is zero except in ONE query row per sample; compare with autograd of the
        # reference attention under that dO (query row 4 is live in every sample with kv_len > 4; de-selected samples and
        # samples shorter than the row get all-zero gradients)
        q_row = 4 if T > 4 else 0
        dO1 = torch.zeros(B, T, 256, device=dev)
        row = torch.randn(B, 256, device=dev).to(GRD)
        dO1[:, q_row, :] = row.float() * (kv > q_row)[:, None]
        qf1 = qkv.float().requires_grad_(True)
        _attn_ref(qf1, kv, B, T).backward(dO1)
        g1 = qf1.grad.view(B * T, 768)
        dQKV3 = torch.full((B * T, 768), 3.0, device=dev, dtype=GRD)
        ops.attn_bwd_single_query(qkv, row.contiguous(), O.view(B, T, 256)[:, q_row, :].contiguous(), kv, B, T, q_row, lse,
                                  dQKV3)
        res[f"single_dQ_B{B}_T{T}"] = _err(dQKV3[:, :256], g1[:, :256])
        res[f"single_dK_B{B}_T{T}"] = _err(dQKV3[:, 256:512], g1[:, 256:512])
        res[f"single_dV_B{B}_T{T}"] = _err(dQKV3[:, 512:], g1[:, 512:])
    res["ok"] = all(v["rel_to_max"] < 5e-3 and v["finite"] for v in res.values() if isinstance(v, dict))
    return res


def run_case(name):
    import torch
    t0 = time.time()
    out = globals()["case_" + name]()
    torch.cuda.synchronize()
    out["seconds"] = round(time.time() - t0, 2)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--timeout", type=int, default=180)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kernel_check.json"))
    args = ap.parse_args()
    if args.case:
        print("RESULT " + json.dumps(run_case(args.case)))
        return
    summary = {}
    for name in CASES:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name], capture_output=True,
                               text=True, timeout=args.timeout)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                summary[name] = json.loads(line[-1][7:])
            else:
                summary[name] = {"ok": False, "rc": r.returncode, "stderr": r.stderr[-1500:], "stdout": r.stdout[-800:]}
        except subprocess.TimeoutExpired:
            summary[name] = {"ok": False, "timeout": True}
        print(name, "OK" if summary[name].get("ok") else "FAIL", flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps({k: v.get("ok") for k, v in summary.items()}))


if __name__ == "__main__":
    main()
