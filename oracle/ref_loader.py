"""ORACLE SUPPORT (test / baseline infrastructure, NOT product code): import the UNMODIFIED reference modules from
oracle/_ref (see oracle/build_ref.py) with the shims of SURVEY.md Appendix A:
  * `monai` is imported by tri_mbt_vsltcls.py:11 but unused with the swin encoder -> stub modules;
  * `swin_t_m` is always asked for ImageNet weights (tri_mbt_vsltcls.py:93,102): no network -> weights=None;
  * `control.config` runs argparse at import -> sys.argv is set first.
On CPU the reference trainer additionally needs `Tensor.cuda` / `torch.HalfTensor` shims (trainer.py:26-27,77-84).
The repo ships its own `builder/` drop-in package, which would shadow the reference's; `load()` therefore must run in a
process that has not imported the repo's `builder` (bench.py --impl reference*), or use `load_trainer_only()`.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

from . import build_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _argv(n_layers, batch_size, multiimages, dropout):
    return ["x", "--model", "tri_mbt_vsltcls", "--input-types", "vslt_img_txt", "--vslt-type", "TIE",
            "--imgtxt-time", "1", "--mbt-only-vslt", "1", "--multiimages", str(multiimages),
            "--transformer-num-layers", str(n_layers), "--batch-size", str(batch_size), "--dropout", str(dropout),
            "--img-pretrain", "No", "--modality-inclusion", "train-missing_test-missing"]


def _stub_monai():
    for n in ("monai", "monai.networks", "monai.networks.blocks", "monai.networks.blocks.patchembedding"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["monai.networks.blocks.patchembedding"].PatchEmbeddingBlock = object


def ref_root() -> str:
    if not build_ref.available():
        build_ref.build()
    if not build_ref.available():
        raise RuntimeError("oracle/_ref is missing: run `python -m oracle.build_ref` where /root/reference is mounted")
    return build_ref.REF_DST


def load(n_layers=6, batch_size=64, multiimages=1, dropout=0.1, device="cpu"):
    """Returns (args, model_module, trainer_module) of the reference itself. Whole-process import: the reference's
    `builder` / `control` packages take the names."""
    import torch
    if "builder" in sys.modules and not getattr(sys.modules["builder"], "__file__", None) is None:
        raise RuntimeError("the repo's `builder` shim is already imported in this process; the reference's namespace "
                           "package of the same name cannot be loaded next to it")
    ref = ref_root()
    sys.dont_write_bytecode = True
    sys.path[:] = [ref] + [p for p in sys.path if os.path.abspath(p or ".") not in (ROOT, ref)]
    saved_argv = sys.argv
    sys.argv = _argv(n_layers, batch_size, multiimages, dropout)
    _stub_monai()
    try:
        cfg = importlib.import_module("control.config")
    finally:
        sys.argv = saved_argv
    args = cfg.args
    args.device = torch.device(device)
    args.feature_means = torch.zeros(16)                     # trainer.py:46 reads it; unused by the TIE model
    mod = importlib.import_module("builder.models.8_missing_models.tri_mbt_vsltcls")
    orig = mod.swin_t_m
    mod.swin_t_m = lambda weights=None, **k: orig(weights=None, **k)
    trainer = importlib.import_module("builder.trainer")
    sys.path.append(ROOT)
    assert mod.__file__.startswith(ref), mod.__file__
    return args, mod, trainer


def cpu_trainer_shims():
    """The reference trainer hard-codes `.cuda()` and `torch.HalfTensor` (trainer.py:26-27,77,82,84): identity / fp32 on a
    CPU-only run."""
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.HalfTensor = torch.FloatTensor


def load_trainer_only(n_layers=6, batch_size=64, multiimages=1, dropout=0.1):
    """The reference's unmodified `missing_trainer` (builder/trainer/trainer.py) as a stand-alone module, next to the
    repo's own `builder` package: used to drive the B200 model through the reference's own train step. Returns
    (reference args Namespace, trainer module)."""
    ref = ref_root()
    sys.dont_write_bytecode = True
    if "control.config" not in sys.modules:
        spec = importlib.util.spec_from_file_location("control.config", os.path.join(ref, "control", "config.py"))
        cfg = importlib.util.module_from_spec(spec)
        pkg = types.ModuleType("control")
        pkg.__path__ = [os.path.join(ref, "control")]
        saved_argv = sys.argv
        sys.argv = _argv(n_layers, batch_size, multiimages, dropout)
        try:
            sys.modules["control"] = pkg
            sys.modules["control.config"] = cfg
            spec.loader.exec_module(cfg)
        finally:
            sys.argv = saved_argv
        pkg.config = cfg
    cfg = sys.modules["control.config"]
    spec = importlib.util.spec_from_file_location("_reference_trainer", os.path.join(ref, "builder", "trainer", "trainer.py"))
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)
    return cfg.args, tr
