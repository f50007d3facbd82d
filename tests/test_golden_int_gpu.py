"""The CUDA integer path against the REFERENCE's own outputs (tests/golden/*.npz, written by tools/make_golden.py from the
imported reference): key-padding lengths and masks bit-exact (reference utils.py:79-125, mbt_encoder.py:703-714,748;
tri_mbt_vsltcls.py:226-237), the feature-id gather bit-exact (tri_mbt_vsltcls.py:187-188), the UMSE/TIE embedding within
2e-5 (fp32 output of the fused kernel; :183-189). All calls go through the C ABI (ops -> ctypes -> libtmp_b200.so)."""
import numpy as np
import pytest
import torch

from golden_util import fixture_embedding_view, fixture_inputs, fixture_names, load_fixture

pytestmark = pytest.mark.gpu


def _kv_len(batch, cfg, skip_missing=0):
    from medical_tri_modal_pilot_b200 import ops
    dev = "cuda"
    B, L = batch["x"].shape[:2]
    n_img = 3 if cfg.multiimages else 1
    T = [L + 5, 49 * n_img + 5, 133]
    kv = ops.build_lengths(batch["input_lengths"].to(dev), batch["txt_lengths"].to(dev),
                           batch["img_time"].float().reshape(B, n_img).contiguous().to(dev), n_img, cfg.multiimages,
                           batch["missing"].to(dev), skip_missing, T[0], T[1], T[2])
    return kv, T


@pytest.mark.parametrize("name", fixture_names())
def test_lengths_and_masks_equal_the_reference_masks(name):
    from medical_tri_modal_pilot_b200 import ops
    fx = load_fixture(name)
    sd, batch, cfg = fixture_inputs(fx)
    kv, T = _kv_len(batch, cfg)
    streams = [0, 1, 2] if cfg.multiimages else [0, 2]       # --multiimages 0: the img stream is called with mask=None
    for k, m in enumerate(streams):
        assert np.array_equal(kv[m].cpu().numpy(), fx[f"fused_kvlen_{k}"]), (name, m)
        shape = tuple(int(v) for v in fx[f"fused_mask_{k}_shape"])
        ref_mask = np.unpackbits(fx[f"fused_mask_{k}"])[: int(np.prod(shape))].reshape(shape).astype(bool)
        mine = ops.materialize_mask(kv[m].contiguous(), T[m]).cpu().numpy()
        assert mine.shape == shape and np.array_equal(mine, ref_mask), (name, m)
    if not cfg.multiimages:
        assert (kv[1] == T[1]).all()                         # unmasked == every key visible


@pytest.mark.parametrize("name", fixture_names())
def test_skip_missing_only_zeroes_deselected_streams(name):
    """skip_missing=1 (what the product path uses) differs from the reference lengths only where the `missing` code
    de-selects the stream (exact: SURVEY Appendix A)."""
    fx = load_fixture(name)
    sd, batch, cfg = fixture_inputs(fx)
    kv0, _ = _kv_len(batch, cfg, 0)
    kv1, _ = _kv_len(batch, cfg, 1)
    miss = batch["missing"].numpy()
    img_off = (miss == 2) | (miss == 3)
    txt_off = (miss == 1) | (miss == 3)
    kv0, kv1 = kv0.cpu().numpy(), kv1.cpu().numpy()
    assert np.array_equal(kv1[0], kv0[0])
    assert np.array_equal(kv1[1], np.where(img_off, 0, kv0[1]))
    assert np.array_equal(kv1[2], np.where(txt_off, 0, kv0[2]))


@pytest.mark.parametrize("name", fixture_names())
def test_embedding_and_gather_equal_the_reference(name):
    from medical_tri_modal_pilot_b200 import ops
    fx = load_fixture(name)
    sd, batch, cfg = fixture_inputs(fx)
    dev = "cuda"
    x = batch["x"].float().contiguous().to(dev)
    br = lambda p: [sd[f"{p}.0.weight"].reshape(256).to(dev), sd[f"{p}.0.bias"].to(dev), sd[f"{p}.1.weight"].to(dev),
                    sd[f"{p}.1.bias"].to(dev)]
    W = sd["ie_feat.weight"].to(dev)
    e = ops.umse_embed(x, br("ie_vslt"), br("ie_time"), W, torch.float32).cpu().numpy()
    assert np.abs(fixture_embedding_view(fx, e) - fx["vslt_embedding"]).max() < 2e-5
    z = lambda p: br(p)[:2] + [torch.zeros(256, device=dev), torch.zeros(256, device=dev)]
    g = ops.umse_embed(x, z("ie_vslt"), z("ie_time"), W, torch.float32).cpu().numpy()
    assert np.array_equal(fixture_embedding_view(fx, g), fx["gather_embedding"])
