"""Classifier head of TRI_MBT_VSLTCLS in training mode as three sm_100a launches (csrc/head.cu).

Reference tri_mbt_vsltcls.py:176-177 (demographic branch) and :248-255 (LayerNorm of the vslt CLS row, concat,
Linear(512,256) -> BatchNorm1d -> ReLU -> Linear(256,1)). The stock PyTorch modules stay the parameter / buffer owners
(state_dict names unchanged) and remain the path for eval mode, the `rmse` auxiliary output and anything else this kernel
does not cover (`usable`)."""
from __future__ import annotations

import os

import torch

from . import ops


def usable(model, cls_out) -> bool:
    """Training-mode batch statistics with running-statistics tracking, fp32 parameters on the GPU, 2 <= B <= 4096."""
    if os.environ.get("TMP_B200_FUSED_HEAD", "1") == "0" or not getattr(model, "fused_head", True):
        return False
    bn = model.fc_list[1]
    B = cls_out.shape[0]
    return bool(model.training and bn.training and bn.track_running_stats and bn.momentum is not None and bn.affine
                and cls_out.is_cuda and cls_out.dtype == torch.float32 and 2 <= B <= MAX_B
                and "rmse" not in getattr(model.args, "auxiliary_loss_type", "none")
                and model.fc_list[0].weight.dtype == torch.float32)


def _params(model):
    sd = dict(model.named_parameters())
    return [sd[n] for n in ops.HEAD_PARAM_ORDER]


MAX_B = 4096      # batch statistics of up to this many rows (tmp_head_fwd's documented range)


def _workspace(model, B, dev):
    """Partial-sum scratch + the two last-block counters. Sized once for the largest batch the kernels accept, so the
    buffer a captured step graph points at is never replaced by a later, larger batch."""
    ws = model.__dict__.get("_head_ws")
    need = ops.head_scratch_floats(MAX_B)
    if ws is None or ws[0].device != dev or ws[0].numel() < need:
        ws = (torch.empty(need, dtype=torch.float32, device=dev), torch.zeros(2, dtype=torch.int32, device=dev))
        model.__dict__["_head_ws"] = ws
    return ws


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, cls_out, age, gen, *params):
        B, dev = cls_out.shape[0], cls_out.device
        bn = model.fc_list[1]
        cls_c, age_c, gen_c = cls_out.contiguous(), age.float().contiguous(), gen.float().contiguous()
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        saved = [f(B, 512), f(B, 256), f(B, 256), f(B), f(B), f(B, 256), f(256)]
        logits = f(B)
        scratch, counter = _workspace(model, B, dev)
        pc = [p.detach().contiguous() for p in params]
        ops.head_fwd(cls_c, age_c, gen_c, pc, bn.running_mean, bn.running_var, bn.num_batches_tracked, bn.momentum, bn.eps,
                     saved, scratch, counter[:1], logits)
        ctx.model = model
        ctx.save_for_backward(age_c, gen_c, *saved, *pc)
        return logits.view(B, 1)

    @staticmethod
    def backward(ctx, dlogits):
        t = ctx.saved_tensors
        age_c, gen_c, saved, pc = t[0], t[1], list(t[2:9]), list(t[9:21])
        B, dev = age_c.shape[0], age_c.device
        grads = [torch.empty_like(p) for p in pc]
        dcls = torch.empty(B, 256, dtype=torch.float32, device=dev)
        DH = torch.empty(B, 256, dtype=torch.float32, device=dev)
        scratch, counter = _workspace(ctx.model, B, dev)
        ops.head_bwd(dlogits.reshape(B).float().contiguous(), age_c, gen_c, pc, saved, grads, dcls, DH, scratch, counter[1:])
        return (None, dcls, None, None, *grads)


def fused_head(model, cls_out, age, gen):
    """[B,256] vslt CLS rows (+ age, gender [B]) -> logits [B,1]; updates the BatchNorm running statistics like the module."""
    return _HeadFn.apply(model, cls_out, age, gen, *_params(model))
