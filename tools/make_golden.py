"""Generate tests/golden/*.npz by running the UNMODIFIED reference model imported from /root/reference (read-only,
only available in the build container). Commit the outputs; the GPU box and the CPU test-suite replay them.

    python tools/make_golden.py [case ...]

Import recipe = SURVEY.md Appendix A: stub `monai` (imported, unused), random-init Swin (no network), argv set before
`control.config` is imported. Inputs and weights are NOT stored: they are regenerated bit-identically from
oracle/synth.py and oracle/weights.py (numpy PCG64), so a fixture holds only what the reference computed.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)

CASES = {
    # name: (n_layers, multiimages, B, L, batch seed, weight seed, missing_mode)
    # batches are >= 16 so that the head's BatchNorm1d (batch statistics) is well conditioned, as at the real B=64
    "tri_nl2_multi_B32_L40": (2, 1, 32, 40, 11, 1, "mixed"),
    "tri_nl2_single_B24_L33": (2, 0, 24, 33, 12, 2, "mixed"),
    "tri_nl3_multi_B16_L150": (3, 1, 16, 150, 13, 3, "none"),
    # the bench depth (6 layers, last layer vslt-only) with three 128-key attention tiles per sample
    "tri_nl6_multi_B16_L260": (6, 1, 16, 260, 14, 4, "mixed"),
    # the bench configuration itself (BASELINE config 3/4 shape: B=64, 6 layers, TIE-len 1000, 3 images), ragged lengths
    # and every missing code; the embedding / gather outputs are stored at 1024 seeded (sample, position) rows
    "tri_nl6_multi_B64_L1000": (6, 1, 64, 1000, 15, 5, "mixed"),
}
EMB_FULL_MAX = 4 << 20      # store the whole [B,L,256] embedding below this many elements, else seeded rows


def emb_rows(B, L):
    """(b, l) index pairs at which a large fixture stores the embedding (same generator in the tests)."""
    g = np.random.Generator(np.random.PCG64(B * 100003 + L))
    return g.integers(0, B, 1024), g.integers(0, L, 1024)



def import_reference():
    # The repo ships its own `builder/` drop-in shim (a regular package), which would shadow the reference's
    # namespace package `builder/` whatever the sys.path order: keep the repo root off sys.path while importing.
    sys.path[:] = [REF] + [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
    sys.argv = ["x", "--model", "tri_mbt_vsltcls", "--input-types", "vslt_img_txt", "--vslt-type", "TIE",
                "--imgtxt-time", "1", "--mbt-only-vslt", "1", "--multiimages", "1", "--transformer-num-layers", "2",
                "--batch-size", "4", "--dropout", "0", "--img-pretrain", "No",
                "--modality-inclusion", "train-missing_test-missing"]
    for n in ("monai", "monai.networks", "monai.networks.blocks", "monai.networks.blocks.patchembedding"):
        sys.modules[n] = types.ModuleType(n)
    sys.modules["monai.networks.blocks.patchembedding"].PatchEmbeddingBlock = object
    import torch
    import control.config as C
    args = C.args
    args.device = torch.device("cpu")
    mod = importlib.import_module("builder.models.8_missing_models.tri_mbt_vsltcls")
    orig = mod.swin_t_m
    mod.swin_t_m = lambda weights=None, **k: orig(weights=None, **k)
    enc_mod = importlib.import_module("builder.models.src.transformer.mbt_encoder")
    assert mod.__file__.startswith(REF), mod.__file__
    sys.path.append(ROOT)
    return args, mod, enc_mod


def main():
    import torch
    from oracle import synth, weights
    from oracle import tri_mbt_oracle as O

    only = [a for a in sys.argv[1:] if a in CASES]     # python tools/make_golden.py [case names]; before argv is replaced
    os.chdir("/tmp")
    args, mod, enc_mod = import_reference()
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)

    class FeatStub(torch.nn.Module):
        """Stands in for the frozen Swin-T: returns the synthetic [N,7,7,768] feature map for the given pixels."""
        def __init__(self, feats):
            super().__init__()
            self.feats = feats
        def forward(self, img):
            return self.feats.reshape(-1, 7, 7, 768)

    for name, (nl, multi, B, L, bseed, wseed, mmode) in CASES.items():
        if only and name not in only:
            continue
        torch.manual_seed(0)
        args.transformer_num_layers = nl
        args.multiimages = multi
        args.batch_size = B
        n_img = 3 if multi else 1
        model = mod.TRI_MBT_VSLTCLS(args)
        sd = weights.make_state_dict(nl, wseed)
        ref_sd = model.state_dict()
        non_swin = {k: v for k, v in ref_sd.items() if not k.startswith("img_encoder.")}
        assert set(non_swin) == set(sd), (sorted(set(non_swin) ^ set(sd)))
        for k in sd:
            assert tuple(non_swin[k].shape) == tuple(sd[k].shape), (k, non_swin[k].shape, sd[k].shape)
        assert torch.allclose(ref_sd["fusion_transformer.positional_encoding.pe"],
                              sd["fusion_transformer.positional_encoding.pe"], atol=0, rtol=0)
        model.load_state_dict(sd, strict=False)
        batch = synth.make_batch(B, L, n_img=n_img, seed=bseed, missing_mode=mmode)
        model.img_encoder = FeatStub(batch["img_feats"])
        model.train()

        captured = {}
        real_mask_fn = enc_mod.get_attn_pad_mask
        def spy(padded_input, input_lengths, expand_length):
            m = real_mask_fn(padded_input, input_lengths, expand_length)
            captured.setdefault("masks", []).append(m.clone())
            return m
        enc_mod.get_attn_pad_mask = spy
        def pre_hook(module, a, kw):
            captured["vslt_embedding"] = kw["enc_outputs"][0].detach().clone()
        hdl = model.fusion_transformer.register_forward_pre_hook(pre_hook, with_kwargs=True)

        img = torch.zeros(B, 3, 1, 224, 224) if multi else torch.zeros(B, 1, 224, 224)
        def run():
            return model(batch["x"], None, None, None, None, batch["age"], batch["gen"], batch["input_lengths"].clone(),
                         batch["txts"], batch["txt_lengths"].clone(), img, batch["missing"], None,
                         batch["img_time"].clone(), batch["txt_time"].clone(), "train", None, None)
        model.zero_grad()
        out, _, _ = run()
        loss = torch.nn.BCEWithLogitsLoss()(out.squeeze(), batch["y"])
        loss.backward()
        hdl.remove()
        enc_mod.get_attn_pad_mask = real_mask_fn

        fx = {"logits": out.detach().numpy(), "loss": np.float64(loss.item())}
        sub = B * L * 256 > EMB_FULL_MAX
        take = (lambda e: e.numpy()[emb_rows(B, L)]) if sub else (lambda e: e.numpy())
        fx["vslt_embedding"] = take(captured["vslt_embedding"])
        # masks: self masks first (one per masked stream), then the fused-layer masks in call order v,(i),t
        masks = captured["masks"]
        n_masked = 3 if multi else 2
        assert len(masks) == 2 * n_masked, len(masks)
        fused = masks[n_masked:]
        for k, m in enumerate(fused):
            assert (m == m[:, :1, :]).all()
            fx[f"fused_mask_{k}"] = np.packbits(m.numpy())
            fx[f"fused_mask_{k}_shape"] = np.array(m.shape)
            fx[f"fused_kvlen_{k}"] = (~m[:, 0, :]).sum(-1).numpy().astype(np.int32)
        # gradients: norm, seeded samples and a seeded random projection per live tensor
        gnames = sorted(k for k, p in model.named_parameters() if p.grad is not None)
        fx["grad_names"] = np.array(gnames)
        for k in gnames:
            gr = dict(model.named_parameters())[k].grad.detach().double().flatten().numpy()
            rng = np.random.Generator(np.random.PCG64(abs(hash(k)) % (2 ** 31) if False else len(k) * 7919 + gr.size))
            idx = rng.integers(0, gr.size, 32)
            proj = rng.standard_normal(gr.size)
            fx[f"g/{k}/norm"] = np.float64(np.linalg.norm(gr))
            fx[f"g/{k}/samples"] = gr[idx]
            fx[f"g/{k}/proj"] = np.float64(gr @ proj)
            if gr.size <= 4096:
                fx[f"g/{k}/full"] = gr.astype(np.float32)
        # gather-only run (both LN-ReLU branches silenced): vslt_embedding == ie_feat(feat) bit-exactly
        with torch.no_grad():
            for p in ("ie_vslt", "ie_time"):
                getattr(model, p)[1].weight.zero_(); getattr(model, p)[1].bias.zero_()
        hdl = model.fusion_transformer.register_forward_pre_hook(pre_hook, with_kwargs=True)
        run()
        hdl.remove()
        fx["gather_embedding"] = take(captured["vslt_embedding"])
        fx["emb_subsampled"] = np.array(int(sub))
        fx["config"] = np.array([nl, multi, B, L, bseed, wseed])
        fx["missing_mode"] = np.array(mmode)

        # cross-check the restatement right here (tests/test_oracle_golden.py repeats it from the stored file)
        cfg = O.OracleConfig(n_layers=nl, multiimages=multi)
        lo, ls, grads = O.train_step_grads(sd, batch, cfg)
        print(f"{name}: |logit diff| {np.abs(lo.numpy() - fx['logits']).max():.3e}  loss diff "
              f"{abs(ls.item() - fx['loss']):.3e}  live grads ref={len(gnames)} oracle={len(grads)}")
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **fx)
        print("   wrote", name, f"{os.path.getsize(os.path.join(outdir, name + '.npz')) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
