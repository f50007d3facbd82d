"""eval / inference forward and checkpoint I/O.
  * eval-mode parity against the oracle with non-trivial BatchNorm running statistics (reference trainer.py:192-240:
    model.eval() forward, no dropout, BatchNorm1d running stats), 16-bit and fp32 modes;
  * a checkpoint WRITTEN BY THE REFERENCE ITSELF (its own model class from oracle/_ref, the logger.py:166-177 dict layout)
    loads into the B200 model -- image encoder weights included -- and the eval logits on real pixels match the reference's
    own CPU logits; the B200 model's state_dict round-trips through torch.save / load bit-exactly."""
import io
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from golden_util import fixture_inputs, fixture_names, load_fixture

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("precision,tol", [("fp16", 2e-2), ("fp32", 1e-3)])
def test_eval_mode_matches_oracle_with_running_stats(precision, tol):
    from oracle import tri_mbt_oracle as O
    from test_model_parity_gpu import build_model, run_model
    fx = load_fixture([n for n in fixture_names() if "B64" not in n][0])
    sd, batch, cfg = fixture_inputs(fx)
    g = torch.Generator().manual_seed(5)
    sd = dict(sd)
    sd["fc_list.1.running_mean"] = torch.randn(256, generator=g) * 0.3
    sd["fc_list.1.running_var"] = torch.rand(256, generator=g) + 0.5
    B = batch["x"].shape[0]
    model = build_model(cfg, sd, B, dropout=0.3, precision=precision).eval()      # dropout must be inert in eval mode
    with torch.no_grad():
        out, _ = run_model(model, batch)
        out2, _ = run_model(model, batch)
    assert torch.equal(out, out2)                                                  # deterministic: no dropout, no BN update
    cfg.training = False
    ref = O.forward(sd, batch, cfg)
    rel = ((out.cpu() - ref).abs().max() / ref.abs().max()).item()
    assert rel < tol, rel
    # running statistics untouched by an eval forward
    assert torch.equal(model.fc_list[1].running_mean.cpu(), sd["fc_list.1.running_mean"])
    # and a train-mode forward updates them like nn.BatchNorm1d does
    model.train()
    run_model(model, batch)
    assert not torch.equal(model.fc_list[1].running_mean.cpu(), sd["fc_list.1.running_mean"])


def test_reference_written_checkpoint_loads_and_matches(tmp_path):
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref not present")
    from builder.models import get_model
    from medical_tri_modal_pilot_b200 import synth
    from medical_tri_modal_pilot_b200.config import make_args
    NL, B, L = 2, 4, 50
    ckpt_path, logit_path = str(tmp_path / "best_fold0_seed0.pth"), str(tmp_path / "ref_logits.pt")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_reference_ckpt.py"), ckpt_path, logit_path,
                        str(NL), str(B), str(L)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    assert set(ckpt) == {"model", "optimizer", "best_step", "last_step", "score", "epoch"}
    args = make_args(transformer_num_layers=NL, multiimages=1, mbt_only_vslt=1, input_types="vslt_img_txt", imgtxt_time=1,
                     dropout=0.1, batch_size=B, img_pretrain="No")
    args.device = torch.device("cuda")
    model = get_model(args)(args)
    res = model.load_state_dict(ckpt["model"], strict=True)          # EVERY key of the reference checkpoint, Swin-T included
    assert not res.missing_keys and not res.unexpected_keys
    model = model.to(args.device).eval()
    hb = synth.make_batch(B, L, n_img=3, seed=77, missing_mode="mixed", with_pixels=True, feats=False)
    b = {k: v.cuda() for k, v in hb.items()}
    with torch.no_grad():
        out, _, _ = model(b["x"], None, None, None, None, b["age"], b["gen"], b["input_lengths"], b["txts"], b["txt_lengths"],
                          b["img"], b["missing"], None, b["img_time"], b["txt_time"], "test", None, None)
    ref = torch.load(logit_path)
    rel = ((out.cpu() - ref).abs().max() / ref.abs().max()).item()
    assert rel < 2e-2, rel                                            # fp16 path incl. the native Swin-T forward
    # our own state_dict: same keys, same values, survives torch.save / torch.load
    buf = io.BytesIO()
    torch.save({"model": model.state_dict()}, buf)
    buf.seek(0)
    sd2 = torch.load(buf, map_location="cpu", weights_only=False)["model"]
    assert set(sd2) == set(ckpt["model"])
    for k, v in ckpt["model"].items():
        assert torch.equal(sd2[k], v), k
