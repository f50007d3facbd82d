"""FlatAdamW -- torch.optim.AdamW semantics (reference 2_train.py:110: lr, weight_decay, default betas/eps) with the
fused-path parameters updated by ONE sm_100a kernel over the flat fp32 buffers (runtime.FusedPath.flat_w / flat_g)
instead of ~260 per-tensor updates. Parameters that never receive a gradient (frozen Swin, `rmse_layer`, the last
layer's img/txt blocks under --mbt-only-vslt 1, ...) are skipped exactly like torch.optim.AdamW skips `grad is None`.
It is a torch.optim.Optimizer: LR schedulers (`CosineAnnealingWarmupRestarts`, 2_train.py:119) drive `param_groups`.
"""
from __future__ import annotations

import torch

from . import ops


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        params = [p for p in model.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.fp = model._fused
        dev = next(model.ie_vslt.parameters()).device
        self.fp._ensure_params(dev)
        self.n_live = self.fp.live_end()                       # flat_w[:n_live] receives gradients
        self.m = torch.zeros(self.n_live, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.n_live, dtype=torch.float32, device=dev)
        flat_ids = {id(p) for _, p in self.fp.layout}
        self.rest = [p for p in params if id(p) not in flat_ids]
        self._rest_opt = torch.optim.AdamW(self.rest, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.t = 0

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        self.t += 1
        fp = self.fp
        if fp.grads_fresh:
            ops.adamw_step(fp.flat_w[: self.n_live], fp.flat_g[: self.n_live], self.m, self.v, g["lr"], g["betas"][0],
                           g["betas"][1], g["eps"], g["weight_decay"], self.t)
            fp.grads_fresh = False
        for rg in self._rest_opt.param_groups:
            rg["lr"] = g["lr"]
        self._rest_opt.step()
        return None

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none=True)
