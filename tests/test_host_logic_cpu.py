"""CPU tests of the host-side mirror of the reference interface: registry, flag surface, state_dict names, trainer
helpers, synthetic workload determinism, FLOP accounting."""
import pytest
import torch

from medical_tri_modal_pilot_b200 import synth
from medical_tri_modal_pilot_b200.config import build_parser, make_args
from oracle import tri_mbt_oracle as O
from oracle import weights


def _args(**kw):
    base = dict(transformer_num_layers=2, multiimages=1, mbt_only_vslt=1, input_types="vslt_img_txt", imgtxt_time=1,
                dropout=0.0, batch_size=4, img_pretrain="No")
    base.update(kw)
    return make_args(**base)


def test_registry_resolves_reference_name():
    from builder.models import get_model
    from medical_tri_modal_pilot_b200.model import TRI_MBT_VSLTCLS
    a = _args()
    assert a.model == "tri_mbt_vsltcls"
    assert get_model(a) is TRI_MBT_VSLTCLS
    a.model = "tri_mbt_v1"
    with pytest.raises(ModuleNotFoundError):
        get_model(a)


def test_flag_surface_matches_reference_names():
    p = build_parser()
    a = p.parse_args(["--input-types", "vslt_txt", "--modality-inclusion", "train-missing_test-missing",
                      "--transformer-num-layers", "6", "--TIE-len", "2000", "--multiimages", "1", "--mbt-only-vslt", "1",
                      "--vslt-type", "TIE", "--imgtxt-time", "1"])
    assert (a.input_types, a.TIE_len, a.multiimages, a.mbt_only_vslt, a.transformer_num_layers) == \
        ("vslt_txt", 2000, 1, 1, 6)
    with pytest.raises(SystemExit):
        p.parse_args(["--multiimages", "3"])          # reference choices=[0,1] (control/config.py:32)


@pytest.mark.parametrize("nl", [2, 6])
def test_state_dict_names_and_shapes_match_reference(nl):
    from builder.models import get_model
    a = _args(transformer_num_layers=nl)
    model = get_model(a)(a)
    sd = {k: v for k, v in model.state_dict().items() if not k.startswith("img_encoder.")}
    ref = weights.make_state_dict(nl)              # names/shapes asserted against the reference by tools/make_golden.py
    assert set(sd) == set(ref), set(sd) ^ set(ref)
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    if nl == 6:
        assert len(sd) == 296                        # SURVEY.md 5 [probe]


def test_unsupported_configurations_raise():
    from builder.models import get_model
    for kw in (dict(vslt_type="carryforward"), dict(transformer_dim=128), dict(mbt_fusion_startIdx=2),
               dict(img_model_type="vit"), dict(berttype="bert")):
        a = _args(**kw)
        with pytest.raises(NotImplementedError):
            get_model(a)(a)


def test_forward_on_cpu_raises_instead_of_falling_back():
    from builder.models import get_model
    a = _args()
    model = get_model(a)(a)
    b = synth.make_batch(4, 16, n_img=3, seed=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(b["x"], None, None, None, None, b["age"], b["gen"], b["input_lengths"], b["txts"], b["txt_lengths"],
              b["img_feats"], b["missing"], None, b["img_time"], b["txt_time"], "train", None, None)


def test_missing_to_num_matches_reference_unique_ranking():
    from medical_tri_modal_pilot_b200.trainer import missing_to_num
    g = torch.Generator().manual_seed(0)
    m = torch.zeros(64, 3)
    m[:, 1:] = torch.randint(0, 2, (64, 2), generator=g).float()
    assert torch.equal(missing_to_num(m), O.missing_to_num(m))


def test_tri_missing_code_remap():
    from builder.models import get_model
    two = torch.tensor([0, 1, 0, 1])
    for it, exp in (("vslt", [3, 3, 3, 3]), ("vslt_txt", [2, 3, 2, 3]), ("vslt_img", [1, 3, 1, 3]),
                    ("vslt_img_txt", [0, 1, 0, 1])):
        a = _args(input_types=it)
        m = get_model(a)(a)
        assert m.tri_missing_code(two, 4, torch.device("cpu")).tolist() == exp


def test_synthetic_batch_is_deterministic_and_well_formed():
    a = synth.make_batch(8, 50, n_img=3, seed=5, with_pixels=True)
    b = synth.make_batch(8, 50, n_img=3, seed=5, with_pixels=True)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    L = a["input_lengths"]
    pad = torch.arange(50)[None, :] >= L[:, None]
    assert a["x"][pad].abs().sum() == 0 and L.max() == 50
    assert set(a["missing"].tolist()) == {0, 1, 2, 3}
    tm = (a["missing"] == 1) | (a["missing"] == 3)
    assert (a["txt_lengths"][tm] == 0).all() and (a["txt_lengths"][~tm] > 0).all()
    im = (a["missing"] >= 2)
    assert (a["img_time"][im] == 10).all()
    assert a["img"].shape == (8, 3, 1, 224, 224)


def test_oracle_skip_missing_equivalence():
    """Zeroing kv_len of de-selected streams (the product's skip_missing) cannot change the result: the oracle gives
    identical logits when the data of missing modalities is replaced."""
    sd = weights.make_state_dict(2, 3)
    cfg = O.OracleConfig(n_layers=2, multiimages=1)
    b = synth.make_batch(8, 24, n_img=3, seed=2)
    out1 = O.forward(sd, b, cfg)
    b2 = {k: v.clone() for k, v in b.items()}
    miss = b2["missing"]
    b2["txts"][(miss == 1) | (miss == 3)] = 3.0
    b2["img_feats"].view(8, 3, 49, 768)[(miss >= 2)] = -1.0
    pad = torch.arange(24)[None, :] >= b2["input_lengths"][:, None]
    b2["x"][pad] = torch.tensor([-5.0, 0.3, 7.0])
    out2 = O.forward(sd, b2, cfg)
    assert torch.equal(out1, out2)


def test_fused_head_host_logic():
    """head.py: the parameter order handed to tmp_head_fwd / tmp_head_bwd names real parameters of the expected shapes, the
    scratch size covers both kernels' partial sums, and the three-launch head is never chosen off the GPU, in eval mode or
    when switched off (the stock modules serve those cases)."""
    from builder.models import get_model
    from medical_tri_modal_pilot_b200 import head, ops
    a = _args()
    model = get_model(a)(a).train()
    named = dict(model.named_parameters())
    shapes = {"layer_norms_after_concat.weight": (256,), "ie_demo.0.weight": (256, 2), "fc_list.0.weight": (256, 512),
              "fc_list.1.bias": (256,), "fc_list.3.weight": (1, 256), "fc_list.3.bias": (1,)}
    assert len(ops.HEAD_PARAM_ORDER) == 12 and all(n in named for n in ops.HEAD_PARAM_ORDER)
    for n, shp in shapes.items():
        assert tuple(named[n].shape) == shp, n
    assert [tuple(p.shape) for p in head._params(model)] == [tuple(named[n].shape) for n in ops.HEAD_PARAM_ORDER]
    for B in (2, 64, 4096):
        assert ops.head_scratch_floats(B) >= max(32 * B, 64 * 7 * 256)
    cls = torch.zeros(4, 256)
    assert not head.usable(model, cls)                       # CPU tensor: never
    model.fused_head = False
    assert not head.usable(model, cls)
    model.fused_head = True
    model.eval()
    assert not head.usable(model, cls)
