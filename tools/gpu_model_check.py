"""Quick diagnostic (not a test): fused-path-level agreement with the oracle on the golden fixtures.
 (1) CLS output of the fusion encoder vs oracle, (2) fused-parameter gradients when the ORACLE's dL/dCLS is injected
 (isolates the kernels from the tiny-batch BatchNorm in the head), (3) end-to-end logits."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from golden_util import fixture_inputs, fixture_names, fp16_representable, load_fixture
from test_model_parity_gpu import build_model, run_model
from oracle import tri_mbt_oracle as O

names = sys.argv[1:] or fixture_names()
for name in names:
    fx = load_fixture(name)
    sd, batch, cfg = fixture_inputs(fx)
    if os.environ.get("RAW_WEIGHTS") != "1":
        sd = fp16_representable(sd)
    B = batch["x"].shape[0]
    if torch.cuda.is_available():
        model = build_model(cfg, sd, B).train()
        if os.environ.get("GRAD_SCALE"):
            model._fused.grad_scale = float(os.environ["GRAD_SCALE"])
    # oracle with grads wrt cls
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k and "positional" not in k}
    full = dict(sd); full.update(leaves)
    logits, aux = O.forward(full, batch, cfg, return_aux=True)
    aux["vslt_out"].retain_grad()
    cls_ref = aux["vslt_out"][:, 0].detach()
    loss = O.loss_fn(logits, batch["y"]); loss.backward()
    d_cls = aux["vslt_out"].grad[:, 0].clone()
    if not torch.cuda.is_available():
        print("dry run ok", d_cls.shape, d_cls.norm().item()); continue
    b = {k: v.cuda() for k, v in batch.items()}
    cls = model._fused(b["x"], b["input_lengths"], b["txts"], b["txt_lengths"], model.encode_images(b["img_feats"], None),
                       b["img_time"], b["txt_time"], b["missing"])
    e = (cls.detach().cpu() - cls_ref.detach())
    print(f"{name}: CLS max_abs {e.abs().max():.4f} (max |ref| {cls_ref.abs().max():.3f}) rmse_rel {e.pow(2).mean().sqrt() / cls_ref.pow(2).mean().sqrt():.5f}")
    cls.backward(d_cls.cuda())
    named = dict(model.named_parameters())
    rows = []
    for k, lv in leaves.items():
        if lv.grad is None or named[k].grad is None:
            continue
        a = named[k].grad.detach().double().cpu().flatten().numpy(); r = lv.grad.double().flatten().numpy()
        if np.linalg.norm(r) < 1e-6: continue
        rows.append((float(a @ r / (np.linalg.norm(a) * np.linalg.norm(r) + 1e-30)), k, np.linalg.norm(a), np.linalg.norm(r)))
    rows.sort()
    for c, k, na, nr in rows[:int(os.environ.get("NROWS", "8"))]:
        print(f"   cos {c:.5f} |g| {na:.3e} ref {nr:.3e} {k}")
    print(f"   injected-dCLS grads: min cos {rows[0][0]:.5f} median {np.median([r[0] for r in rows]):.5f} n={len(rows)}")
    model.zero_grad(set_to_none=True)
    out, _ = run_model(model, batch)
    ref = torch.from_numpy(fx["logits"])
    print("   end-to-end logits rel", ((out.detach().cpu() - ref).abs().max() / ref.abs().max()).item())
