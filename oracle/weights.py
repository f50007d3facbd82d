"""ORACLE SUPPORT (test infrastructure): deterministic parameter generator for the non-Swin part of
`TRI_MBT_VSLTCLS` with the reference's state_dict names and shapes (probed from the reference by
tools/make_golden.py, which asserts that names and shapes agree). numpy PCG64 keyed by the parameter name, so the
same tensors can be regenerated on any box without shipping 20+ MB of weights."""
from __future__ import annotations

import hashlib

import numpy as np
import torch

from .tri_mbt_oracle import positional_encoding

D = 256


def param_shapes(n_layers: int) -> dict:
    s = {}
    for p in ("ie_vslt", "ie_time"):
        s[f"{p}.0.weight"] = (D, 1); s[f"{p}.0.bias"] = (D,); s[f"{p}.1.weight"] = (D,); s[f"{p}.1.bias"] = (D,)
    s["ie_feat.weight"] = (20, D)
    s["ie_demo.0.weight"] = (D, 2); s["ie_demo.0.bias"] = (D,); s["ie_demo.1.weight"] = (D,); s["ie_demo.1.bias"] = (D,)
    s["txt_embedding.weight"] = (D, 768); s["txt_embedding.bias"] = (D,)
    s["linear.weight"] = (D, 768); s["linear.bias"] = (D,)
    F = "fusion_transformer"
    s[f"{F}.bottlenecks"] = (1, 4, D)
    s[f"{F}.layer_norms_after_concat.weight"] = (D,); s[f"{F}.layer_norms_after_concat.bias"] = (D,)
    for m in range(3):
        s[f"{F}.cls_token_per_modality.{m}"] = (1, 1, D)
        s[f"{F}.layer_norms_in.{m}.weight"] = (D,); s[f"{F}.layer_norms_in.{m}.bias"] = (D,)
    for l in range(n_layers):
        for m in range(3):
            p = f"{F}.layer_stacks.{l}.{m}"
            for ln in ("attention_prenorm", "feed_forward_prenorm"):
                s[f"{p}.{ln}.gamma"] = (D,); s[f"{p}.{ln}.beta"] = (D,)
            for pr in ("query_proj", "key_proj", "value_proj"):
                s[f"{p}.self_attention.{pr}.linear.weight"] = (D, D); s[f"{p}.self_attention.{pr}.linear.bias"] = (D,)
            s[f"{p}.feed_forward.w_1.weight"] = (4 * D, D, 1); s[f"{p}.feed_forward.w_1.bias"] = (4 * D,)
            s[f"{p}.feed_forward.w_2.weight"] = (D, 4 * D, 1); s[f"{p}.feed_forward.w_2.bias"] = (D,)
    s["rmse_layer.weight"] = (1, 2 * D); s["rmse_layer.bias"] = (1,)
    s["layer_norms_after_concat.weight"] = (D,); s["layer_norms_after_concat.bias"] = (D,)
    s["fc_list.0.weight"] = (D, 2 * D); s["fc_list.0.bias"] = (D,)
    s["fc_list.1.weight"] = (D,); s["fc_list.1.bias"] = (D,)
    s["fc_list.3.weight"] = (1, D); s["fc_list.3.bias"] = (1,)
    s["activations.prelu.weight"] = (1,)
    return s


def _rng(name: str, seed: int) -> np.random.Generator:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return np.random.Generator(np.random.PCG64(int.from_bytes(h[:8], "little")))


def make_state_dict(n_layers: int, seed: int = 0) -> dict:
    """name -> fp32 torch tensor (CPU). Scales keep activations O(1) through the stack; every bias / LayerNorm
    parameter is non-trivial so that no epilogue term can be silently dropped."""
    sd = {}
    for name, shape in param_shapes(n_layers).items():
        g = _rng(name, seed)
        leaf = name.split(".")[-1]
        if name.endswith(("gamma",)) or (leaf == "weight" and len(shape) == 1):
            a = 1.0 + 0.1 * g.standard_normal(shape)
        elif leaf in ("beta", "bias"):
            a = 0.1 * g.standard_normal(shape)
        elif "cls_token" in name or name.endswith("bottlenecks") or name == "ie_feat.weight":
            a = g.standard_normal(shape)
        elif name.endswith(".0.weight") and shape[-1] in (1, 2):       # Linear(1|2, 256)
            a = g.standard_normal(shape)
        else:
            fan_in = shape[1]
            a = g.standard_normal(shape) / np.sqrt(fan_in)
        sd[name] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    sd["fc_list.1.running_mean"] = torch.zeros(D)
    sd["fc_list.1.running_var"] = torch.ones(D)
    sd["fc_list.1.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    sd["fusion_transformer.positional_encoding.pe"] = positional_encoding(D, 2500).unsqueeze(0)
    return sd
