"""Flag surface of the reference's control/config.py that the hot path reads (same names, defaults, choices;
reference control/config.py:18-122). `control.config` itself parses sys.argv at import time and is reused as-is
when the reference tree is on sys.path; this mirror exists so that the B200 path, the tests and the bench can build
the same `args` Namespace on a box where the reference tree is absent."""
from __future__ import annotations

import argparse


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(add_help=False)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--window-size", type=int, default=24)
    p.add_argument("--vslt-type", type=str, default="TIE", choices=["carryforward", "TIE", "QIE"])
    p.add_argument("--multiimages", type=int, default=0, choices=[0, 1])
    p.add_argument("--TIE-len", type=int, default=1000)
    p.add_argument("--input-types", type=str, default="vslt", choices=["vslt", "vslt_img", "vslt_txt", "vslt_img_txt"])
    p.add_argument("--modality-inclusion", type=str, default="train-full_test-full",
                   choices=["train-full_test-full", "train-missing_test-missing", "train-full_test-missing"])
    p.add_argument("--fullmodal-definition", type=str, default="txt1_img1", choices=["txt1_img1", "img1", "txt1"])
    p.add_argument("--imgtxt-time", type=int, default=0, choices=[0, 1])
    p.add_argument("--batch-size", type=int, default=32)
    p.add_argument("--dropout", type=float, default=0.1)
    p.add_argument("--lr-init", type=float, default=1e-3)
    p.add_argument("--weight_decay", "-wd", type=float, default=1e-6)
    p.add_argument("--berttype", type=str, default="biobert", choices=["biobert", "bert"])
    p.add_argument("--transformer-dim", type=int, default=256)
    p.add_argument("--transformer-num-layers", type=int, default=6)
    p.add_argument("--transformer-num-head", type=int, default=4)
    p.add_argument("--img-model-type", type=str, default="swin", choices=["resnet18", "resnet50", "swin", "vit", "maxvit"])
    p.add_argument("--img-pretrain", type=str, default="Yes", choices=["No", "Yes"])
    p.add_argument("--image-size", type=int, default=224, choices=[224, 512])
    p.add_argument("--residual-bottlenecks", type=int, default=0, choices=[0, 1])
    p.add_argument("--mbt-bottlenecks-n", type=int, default=4)
    p.add_argument("--mbt-fusion-startIdx", type=int, default=0)
    p.add_argument("--mbt-only-vslt", type=int, default=0)
    p.add_argument("--auxiliary-loss-type", type=str, default="None", choices=["None", "rmse", "tdecoder", "tdecoder_rmse"])
    p.add_argument("--vitalsign-labtest", type=list,
                   default=["HR", "RR", "BT", "SBP", "DBP", "Sat", "Hematocrit", "PLT", "WBC", "Bilirubin", "pH", "HCO3",
                            "Creatinine", "Lactate", "Potassium", "Sodium"])
    p.add_argument("--model", type=str, default="tri_mbt_vsltcls")
    return p


def make_args(argv=None, **overrides):
    """Namespace with the reference's defaults; `overrides` use the attribute names (e.g. transformer_num_layers=2)."""
    args = build_parser().parse_args(argv or [])
    for k, v in overrides.items():
        setattr(args, k, v)
    return args
