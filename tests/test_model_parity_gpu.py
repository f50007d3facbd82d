"""GPU parity of the B200 path against the reference (golden fixtures) and the pinned CPU oracle.
Tolerances are the north-star's 16-bit bars: logits within 2e-2 relative, parameter-gradient cosine >= 0.999."""
import numpy as np
import pytest
import torch

from golden_util import fixture_inputs, fixture_names, fp16_representable, gpu_fixture_names, load_fixture

pytestmark = pytest.mark.gpu
LOGIT_RTOL_BF16 = 2e-2
GRAD_COS_MIN = 0.999       # north-star: every parameter gradient, cosine >= 0.999
E2E_MARGIN = 1.5e-2        # end-to-end leg: allowed distance below the 16-bit storage-plan emulation (see the test)


def build_model(cfg, sd, B, dropout=0.0, input_types="vslt_img_txt", precision="fp16"):
    from medical_tri_modal_pilot_b200.config import make_args
    from builder.models import get_model
    args = make_args(transformer_num_layers=cfg.n_layers, multiimages=cfg.multiimages, mbt_only_vslt=cfg.vsltonly,
                     input_types=input_types, imgtxt_time=1, dropout=dropout, batch_size=B, img_pretrain="No")
    args.device = torch.device("cuda")
    args.precision = precision
    model = get_model(args)(args)
    res = model.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all(k.startswith("img_encoder.") for k in res.missing_keys)
    return model.to(args.device)


def run_model(model, batch, dev="cuda"):
    b = {k: v.to(dev) for k, v in batch.items()}
    out, o2, o3 = model(b["x"], None, None, None, None, b["age"], b["gen"], b["input_lengths"], b["txts"],
                        b["txt_lengths"], b["img_feats"], b["missing"], None, b["img_time"], b["txt_time"], "train",
                        None, None)
    assert o2 is None and o3 is None and out.shape == (b["x"].shape[0], 1)
    return out, b


def _oracle_grads(sd, batch, cfg, d_cls=None, autocast=None):
    """Oracle parameter gradients. d_cls=None: end-to-end BCE loss (also returns dL/dCLS); else the gradient of
    <CLS output, d_cls> (isolates the fusion encoder from the BatchNorm-on-a-tiny-batch head)."""
    import contextlib
    from oracle import tri_mbt_oracle as O
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point
              and "running" not in k and "positional_encoding" not in k}
    full = dict(sd)
    full.update(leaves)
    cm = torch.autocast("cpu", dtype=autocast) if autocast is not None else contextlib.nullcontext()
    with cm:
        logits, aux = O.forward(full, batch, cfg, return_aux=True)
    if d_cls is None:
        aux["vslt_out"].retain_grad()
        O.loss_fn(logits.float(), batch["y"]).backward()
        d = aux["vslt_out"].grad[:, 0].detach().clone().float()
    else:
        (aux["vslt_out"][:, 0].float() * d_cls).sum().backward()
        d = None
    return {k: v.grad for k, v in leaves.items() if v.grad is not None}, d, logits.detach().float()


def _cosines(got, ref, floor=1e-6):
    rows = {}
    for k, r in ref.items():
        if k not in got or got[k] is None:
            continue
        a = got[k].detach().double().cpu().flatten().numpy()
        r = r.double().flatten().numpy()
        if np.linalg.norm(r) < floor:
            continue
        rows[k] = float(a @ r / (np.linalg.norm(a) * np.linalg.norm(r) + 1e-30))
    ga = np.concatenate([got[k].detach().double().cpu().flatten().numpy() for k in rows])
    gr = np.concatenate([ref[k].double().flatten().numpy() for k in rows])
    return rows, float(ga @ gr / (np.linalg.norm(ga) * np.linalg.norm(gr)))


def _report(name, leg, rows, glob, extra=None):
    """Per-tensor cosines of one leg -> gpurun_out/parity_<fixture>.json (evidence for DESIGN.md 2) + one summary line."""
    import json, os
    vals = np.array(list(rows.values()))
    worst = sorted(rows.items(), key=lambda kv: kv[1])[:8]
    print(f"[parity {name} {leg}] global {glob:.5f} median {np.median(vals):.5f} min {vals.min():.4f} "
          f"n<0.999 {int((vals < 0.999).sum())}/{len(vals)} worst {worst[:3]}")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, f"parity_{name}.json")
        d = json.load(open(path)) if os.path.exists(path) else {}
        d[leg] = {"global": glob, "median": float(np.median(vals)), "min": float(vals.min()),
                  "below_0.999": {k: v for k, v in rows.items() if v < 0.999}, **(extra or {})}
        json.dump(d, open(path, "w"), indent=1)
    except OSError:
        pass


@pytest.mark.parametrize("name", gpu_fixture_names())
def test_logits_and_grads(name):
    """Every golden fixture, including the bench depth (6 layers) and the bench shape (B=64, TIE-len 1000).

    Logits vs the reference's own fp32 output: north-star 16-bit bar 2e-2 relative. Gradients vs the pinned fp32 oracle at
    IDENTICAL (fp16-representable) weights, two legs:

    (1) the fused path under a GENERIC upstream gradient (seeded Gaussian dL/dCLS): north-star bar, EVERY live tensor
        cosine >= 0.999.
    (2) end to end through the head and the BCE loss. Here the upstream gradient is special: the head's train-mode
        BatchNorm1d removes the batch mean in its backward, so sum_b dL/dCLS_b = 0 and every late-layer parameter
        gradient is a difference of nearly equal per-sample terms -- ill-conditioned with respect to ANY perturbation of
        the forward values. Yardstick: oracle/precision_sim.py, the fp32 oracle with fp16 rounding at exactly the tensors
        the kernels store in 16 bits ("all16"): the B200 path must stay in the band that storage plan produces, and the
        per-tensor values are written to gpurun_out/parity_<fixture>.json and summarised in DESIGN.md 2. The reference's
        own fp16 autocast is no better (test_not_worse_than_reference_fp16_autocast); the fp32 mode meets 0.999 end to
        end on every tensor (test_fp32_mode_*)."""
    from oracle import precision_sim as PS
    fx = load_fixture(name)
    sd, batch, cfg = fixture_inputs(fx)
    sd = fp16_representable(sd)          # identical weights on both sides (see golden_util.fp16_representable)
    B = batch["x"].shape[0]
    model = build_model(cfg, sd, B)
    model.train()
    out, b = run_model(model, batch)
    ref = torch.from_numpy(fx["logits"])  # the reference's own fp32 logits (at the unrounded weights)
    rel = ((out.detach().cpu() - ref).abs().max() / ref.abs().max()).item()
    print(f"[parity {name}] logits rel err {rel:.2e}")
    assert rel < LOGIT_RTOL_BF16, f"logits rel err {rel}"
    loss = torch.nn.BCEWithLogitsLoss()(out.squeeze(), b["y"])
    loss.backward()
    assert abs(loss.item() - float(fx["loss"])) < 2e-2
    named = dict(model.named_parameters())
    live = sorted(k for k, p in named.items() if p.grad is not None and not k.startswith("img_encoder."))
    g_e2e = {k: named[k].grad.detach().clone() for k in live}
    g_ref, _, _ = _oracle_grads(sd, batch, cfg)
    assert live == sorted(g_ref), sorted(set(live) ^ set(g_ref))

    # (1) generic upstream gradient through the fused path alone
    gen = torch.Generator().manual_seed(1234)
    R = torch.randn(B, 256, generator=gen) * 0.02
    model.zero_grad(set_to_none=True)
    cls = model._fused(b["x"], b["input_lengths"], b["txts"], b["txt_lengths"], model.encode_images(b["img_feats"], None),
                       b["img_time"], b["txt_time"], b["missing"])
    cls.backward(R.cuda())
    g_R, _, _ = _oracle_grads(sd, batch, cfg, d_cls=R)
    rows, glob = _cosines({k: p.grad for k, p in named.items() if p.grad is not None}, g_R)
    _report(name, "generic_upstream", rows, glob)
    worst = min(rows.items(), key=lambda kv: kv[1])
    assert worst[1] >= GRAD_COS_MIN, ("fused path, generic upstream gradient", worst)
    assert glob >= 0.9995, glob

    # (2) end to end (BatchNorm-centred upstream gradient), against the storage-plan emulation
    rows_e, glob_e = _cosines(g_e2e, g_ref, floor=1e-4)
    g_sim, _, _ = PS.grads(sd, batch, cfg, PS.PLANS["all16"])
    rows_s, glob_s = _cosines(g_sim, g_ref, floor=1e-4)
    _report(name, "end_to_end", rows_e, glob_e, {"emulation_global": glob_s,
                                                 "emulation_median": float(np.median(list(rows_s.values()))),
                                                 "emulation_min": float(min(rows_s.values()))})
    # The emulation and the kernels perturb the forward values by the same AMOUNT (checked above through the logits and in
    # leg (1)) but not by the same VALUES, and under the centred upstream gradient the outcome is chaotic in the values:
    # the emulation itself moves between 0.979 and 0.992 (minimum tensor) when its rounding sites are varied
    # (python -m oracle.precision_sim). So this leg asserts a band, not a match: within E2E_MARGIN of the emulation, and
    # absolute floors; the 0.999 end-to-end bar is met by the fp32 mode (test_fp32_mode_*).
    assert glob_e >= min(0.999, glob_s - E2E_MARGIN) and glob_e >= 0.98, ("end-to-end global", glob_e, "emulation", glob_s)
    med_e, med_s = np.median(list(rows_e.values())), np.median(list(rows_s.values()))
    assert med_e >= min(0.999, med_s - 2 * E2E_MARGIN) and med_e >= 0.97, ("end-to-end median", med_e, "emulation", med_s)
    for k in live:
        nr = g_ref[k].norm().item()
        if nr < 1e-4:               # mathematically-zero gradients (see test_oracle_golden): only bound the magnitude
            assert g_e2e[k].norm().item() < 5e-3, k
        else:
            assert abs(g_e2e[k].norm().item() / nr - 1) < 0.08, (k, g_e2e[k].norm().item(), nr)


@pytest.mark.parametrize("name", gpu_fixture_names())
def test_fp32_mode_logits_and_grads(name):
    """North-star FP32 parity mode (`args.precision = "fp32"`: fp32 storage, bf16x3-split tcgen05 GEMMs, fp32 attention):
    logits within 1e-3 relative of the REFERENCE's logits, loss, and EVERY parameter gradient of the END-TO-END loss
    (through the head's BatchNorm, the ill-conditioned case of the 16-bit plan) cosine >= 0.999 against the oracle -- at the
    reference's own unrounded weights -- plus the reference's stored gradient probes (norm, 32 samples, a random projection
    per tensor, written by tools/make_golden.py from the reference's autograd)."""
    from golden_util import grad_probe
    fx = load_fixture(name)
    sd, batch, cfg = fixture_inputs(fx)
    B = batch["x"].shape[0]
    model = build_model(cfg, sd, B, precision="fp32").train()
    out, b = run_model(model, batch)
    ref = torch.from_numpy(fx["logits"])
    rel = ((out.detach().cpu() - ref).abs().max() / ref.abs().max()).item()
    print(f"[parity {name} fp32-mode] logits rel err {rel:.2e}")
    assert rel < 1e-3, rel
    loss = torch.nn.BCEWithLogitsLoss()(out.squeeze(), b["y"])
    loss.backward()
    assert abs(loss.item() - float(fx["loss"])) < 1e-4
    named = dict(model.named_parameters())
    live = sorted(k for k, p in named.items() if p.grad is not None and not k.startswith("img_encoder."))
    assert live == sorted(str(n) for n in fx["grad_names"])
    g_ref, _, _ = _oracle_grads(sd, batch, cfg)
    rows, glob = _cosines({k: named[k].grad for k in live}, g_ref, floor=1e-4)
    _report(name, "fp32_mode_end_to_end", rows, glob)
    worst = min(rows.items(), key=lambda kv: kv[1])
    assert worst[1] >= GRAD_COS_MIN and glob >= 0.99999, (worst, glob)
    for k in live:                                           # the reference's own gradient probes
        rn = float(fx[f"g/{k}/norm"])
        nrm, samples, proj = grad_probe(k, named[k].grad.detach().cpu().numpy())
        if rn < 1e-4:
            assert nrm < 1e-3, (k, nrm, rn)
            continue
        assert abs(nrm - rn) <= 5e-3 * rn, (k, nrm, rn)
        assert np.allclose(samples, fx[f"g/{k}/samples"], rtol=2e-2, atol=5e-3 * rn), k


def test_not_worse_than_reference_fp16_autocast():
    """The reference trains under torch.cuda.amp.autocast() (fp16, trainer.py:126). Its algorithm (the pinned oracle) run
    under fp16 autocast deviates from its fp32 self MORE than the B200 path does -- per-tensor minimum and median."""
    fx = load_fixture("tri_nl3_multi_B16_L150")
    sd, batch, cfg = fixture_inputs(fx)
    sd = fp16_representable(sd)
    B = batch["x"].shape[0]
    _, d_cls, logits32 = _oracle_grads(sd, batch, cfg)
    g32, _, _ = _oracle_grads(sd, batch, cfg, d_cls=d_cls)
    g16, _, logits16 = _oracle_grads(sd, batch, cfg, d_cls=d_cls, autocast=torch.float16)
    rows16, glob16 = _cosines(g16, g32)
    model = build_model(cfg, sd, B).train()
    b = {k: v.cuda() for k, v in batch.items()}
    cls = model._fused(b["x"], b["input_lengths"], b["txts"], b["txt_lengths"], model.encode_images(b["img_feats"], None),
                       b["img_time"], b["txt_time"], b["missing"])
    cls.backward(d_cls.cuda())
    rows, glob = _cosines({k: p.grad for k, p in model.named_parameters() if p.grad is not None}, g32)
    print(f"fp16-autocast oracle: min {min(rows16.values()):.4f} median {np.median(list(rows16.values())):.5f} "
          f"global {glob16:.5f} | B200 path: min {min(rows.values()):.4f} median {np.median(list(rows.values())):.5f} "
          f"global {glob:.5f}")
    assert min(rows.values()) >= min(rows16.values())
    assert np.median(list(rows.values())) >= np.median(list(rows16.values())) - 1e-4
    assert glob >= glob16 - 2e-3


def test_input_types_map_to_tri_missing_code():
    """--input-types vslt / vslt_txt / vslt_img == tri model with missing code 3 / {2,3} / {1,3} (the
    2-modal codes {0,1} are the trainer's remap, reference trainer.py:99-105)."""
    from oracle import tri_mbt_oracle as O
    fx = load_fixture(fixture_names()[0])
    sd, batch, cfg = fixture_inputs(fx)
    B = batch["x"].shape[0]
    two = (torch.arange(B) % 2).to(torch.long)                       # per-sample "second modality missing" flag
    for it, tri in (("vslt", torch.full((B,), 3)), ("vslt_txt", 2 + two), ("vslt_img", 1 + 2 * two)):
        model = build_model(cfg, sd, B, input_types=it).train()
        b1 = dict(batch)
        b1["missing"] = two
        out, _ = run_model(model, b1)
        b2 = dict(batch)
        b2["missing"] = tri.to(torch.long)
        ref = O.forward(sd, b2, cfg)
        rel = ((out.detach().cpu() - ref).abs().max() / ref.abs().max()).item()
        assert rel < LOGIT_RTOL_BF16, (it, rel)


def test_padding_and_missing_streams_are_dead():
    """Perturbing padded vslt rows and the data of missing modalities changes the logits by exactly 0."""
    fx = load_fixture(fixture_names()[0])
    sd, batch, cfg = fixture_inputs(fx)
    B = batch["x"].shape[0]
    model = build_model(cfg, sd, B).train()
    out1, _ = run_model(model, batch)
    b2 = {k: v.clone() for k, v in batch.items()}
    L = b2["x"].shape[1]
    pad = torch.arange(L)[None, :] >= b2["input_lengths"][:, None]
    b2["x"][pad] = torch.tensor([-3.0, 0.7, 5.0])
    miss = b2["missing"]
    b2["txts"][(miss == 1) | (miss == 3)] = 1.5
    f = b2["img_feats"].view(B, -1, 49, 768)
    f[(miss == 2) | (miss == 3)] = -2.0
    out2, _ = run_model(model, b2)
    assert torch.equal(out1, out2)


def test_eval_mode_and_state_dict_roundtrip():
    fx = load_fixture(fixture_names()[0])
    sd, batch, cfg = fixture_inputs(fx)
    B = batch["x"].shape[0]
    model = build_model(cfg, sd, B).eval()
    with torch.no_grad():
        out, _ = run_model(model, batch)
    assert torch.isfinite(out).all()
    sd2 = model.state_dict()
    for k, v in sd.items():
        assert torch.equal(sd2[k].cpu(), v), k


@pytest.mark.parametrize("name,nl_multi", [("tri_nl3_multi_B16_L150", None), ("tri_nl6_multi_B16_L260", None)])
def test_fused_gradient_dropout_equals_the_separate_pass(name, nl_multi):
    """With dropout on, the gradient entering a layer's FFN2 is dropout(dX) under that layer's output mask. The LayerNorm
    backward of the layer above writes that copy in the same pass (dx_drop) and the bottleneck exchange refreshes rows 0..3
    of it; the older form ran one `dropout_apply` kernel per layer and stream. Two backward passes from the SAME forward
    (same masks), one per form: the parameter gradients must agree to fp16 rounding of one intermediate tensor."""
    fx = load_fixture(name)
    sd, batch, cfg = fixture_inputs(fx)
    sd = fp16_representable(sd)
    B = batch["x"].shape[0]
    model = build_model(cfg, sd, B, dropout=0.1)
    model.train()
    b = {k: v.to("cuda") for k, v in batch.items()}
    fp = model._fused
    cls = fp(b["x"], b["input_lengths"], b["txts"], b["txt_lengths"], model.encode_images(b["img_feats"], None),
             b["img_time"], b["txt_time"], b["missing"])
    gen = torch.Generator().manual_seed(4321)
    R = (torch.randn(B, 256, generator=gen) * 0.02).cuda()
    grads = {}
    for fused in (True, False):
        fp.fuse_grad_dropout = fused
        model.zero_grad(set_to_none=True)
        fp.backward(R.contiguous())
        torch.cuda.synchronize()
        grads[fused] = fp.flat_g[: fp.live_end()].clone()
    fp.fuse_grad_dropout = True
    a, c = grads[True].double(), grads[False].double()
    assert torch.isfinite(a).all() and a.abs().sum() > 0
    cos = (a @ c / (a.norm() * c.norm())).item()
    rel = ((a - c).norm() / c.norm()).item()
    print(f"[fused grad dropout {name}] cosine {cos:.8f} rel diff {rel:.2e}")
    assert cos > 0.99999 and rel < 3e-3
