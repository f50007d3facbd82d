"""Per-op device-time attribution of one training step at the bench workload (single-stream mode so that the
per-op CUDA-event intervals do not overlap). Every `ops.*` wrapper is bracketed by events; the label carries the op
name, the phase (swin / fwd / bwd / opt) and the GEMM shape.

    TMP_B200_SINGLE_STREAM=1 python tools/step_breakdown.py [--tie-len 1000] [--out gpurun_out/breakdown.json]
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import sys

os.environ.setdefault("TMP_B200_SINGLE_STREAM", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from builder.models import get_model  # noqa: E402
from medical_tri_modal_pilot_b200 import ops, synth, trainer  # noqa: E402
from medical_tri_modal_pilot_b200.config import make_args  # noqa: E402
from medical_tri_modal_pilot_b200.optim import FlatAdamW  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tie-len", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "breakdown.json"))
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    args = make_args(transformer_num_layers=6, multiimages=1, mbt_only_vslt=1, input_types="vslt_img_txt",
                     imgtxt_time=1, dropout=0.1, batch_size=a.batch, img_pretrain="No", TIE_len=a.tie_len)
    args.device = dev
    torch.manual_seed(0)
    model = get_model(args)(args).to(dev).train()
    opt = FlatAdamW(model, lr=1e-4, weight_decay=1e-6)
    crit = torch.nn.BCEWithLogitsLoss()
    host = synth.make_batch(a.batch, a.tie_len, n_img=3, seed=1000, full_length=True, missing_mode="none",
                            with_pixels=True, feats=False)
    miss = host["missing"]
    host["missing3"] = torch.stack([torch.zeros_like(miss), (miss >= 2).long(), (miss % 2).long()], 1).float()
    host["static"] = torch.stack([host["gen"], host["age"]], 1)
    r = {k: v.to(dev) for k, v in host.items()}
    prepared = trainer.prepare_batch(args, dev, r["x"], r["static"], r["input_lengths"], r["y"], r["img"], r["txts"],
                                     r["txt_lengths"], (r["img_time"], r["txt_time"]), r["missing3"])
    step = lambda i: trainer.train_step(args, model, opt, crit, prepared, None, i, None)
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    # host issue time vs device time (is the step launch-bound?)
    import time
    w0 = time.perf_counter()
    for i in range(10):
        step(i)
    w1 = time.perf_counter()
    torch.cuda.synchronize()
    w2 = time.perf_counter()
    print(f"host issue {1e2 * (w1 - w0):.2f} ms/step, issue+drain {1e2 * (w2 - w0):.2f} ms/step (10 steps)")

    records = []
    phase = {"p": "fwd"}

    def wrap(name):
        orig = getattr(ops, name)

        def f(*args_, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            label = name
            if name == "gemm":
                A, Bw = args_[0], args_[1]
                K = A.shape[-1]
                label = f"gemm M={A.numel() // K} N={Bw.shape[0]} K={K}"
            elif name == "gemm_wgrad":
                dY, X = args_[0], args_[1]
                label = f"wgrad M={dY.numel() // dY.shape[-1]} N={dY.shape[-1]} K={X.shape[-1]}"
            elif name in ("attn_fwd", "attn_bwd"):
                label = f"{name} T={args_[3] if name == 'attn_fwd' else args_[5]}" + (" q_rows=5" if kw.get("q_rows") else "")
            elif name in ("layernorm_fwd", "layernorm_bwd", "layernorm_bwd_attn", "colsum", "dropout_apply"):
                t = args_[0]
                label = f"{name} rows={t.numel() // t.shape[-1]} N={t.shape[-1]}"
            e0.record()
            out = orig(*args_, **kw)
            e1.record()
            records.append((phase["p"], label, e0, e1))
            return out
        setattr(ops, name, f)

    for n in ("gemm", "gemm_wgrad", "colsum", "attn_fwd", "attn_bwd", "layernorm_fwd", "layernorm_bwd", "layernorm_bwd_attn",
              "stream_prologue_fwd", "stream_prologue_bwd", "bottleneck_mix_fwd", "bottleneck_mix_bwd", "dropout_apply",
              "cast_weights", "adamw_step", "build_lengths", "swin_patch_embed_ln", "swin_ln_window", "swin_window_attn",
              "swin_unwindow_add_ln", "swin_merge_ln"):
        wrap(n)
    # phases: swin = inside encode_images, bwd = inside FusedPath.backward
    enc = model.encode_images
    fb = model._fused.backward

    def enc2(*x, **k):
        phase["p"] = "swin"
        try:
            return enc(*x, **k)
        finally:
            phase["p"] = "fwd"

    def fb2(*x, **k):
        phase["p"] = "bwd"
        try:
            return fb(*x, **k)
        finally:
            phase["p"] = "opt"
    model.encode_images = enc2
    model._fused.backward = fb2

    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    phase["p"] = "fwd"
    step(3)
    t1.record()
    torch.cuda.synchronize()
    total = t0.elapsed_time(t1)
    agg = collections.OrderedDict()
    for ph, label, e0, e1 in records:
        k = (ph, label)
        d = agg.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += e0.elapsed_time(e1)
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    covered = sum(v[1] for _, v in rows)
    print(f"step (single stream, with event overhead) {total:.3f} ms; inside ops.* {covered:.3f} ms")
    byphase = collections.Counter()
    for (ph, label), (n, ms) in rows:
        byphase[ph] += ms
    print("by phase:", {k: round(v, 3) for k, v in byphase.items()})
    for (ph, label), (n, ms) in rows:
        print(f"{ph:5s} {label:44s} n={n:3d} total {ms:8.3f} ms  avg {ms / n * 1e3:8.1f} us")
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump({"total_ms": total, "ops_ms": covered, "by_phase": dict(byphase),
                   "rows": [{"phase": ph, "op": label, "n": n, "ms": ms} for (ph, label), (n, ms) in rows]}, f, indent=1)


if __name__ == "__main__":
    main()
