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
        # The step count and the learning rate live in device memory (`t_dev`, `lr_dev`): the whole optimizer step is
        # a fixed sequence of launches with no per-step host scalars, i.e. replayable inside a captured CUDA graph
        # (trainer.GraphedStep). The few head parameters use torch.optim.AdamW in its capturable form.
        # t_dev = {step() calls, calls skipped, last call whose gradient was non-finite}. Gradients of the 16-bit plan pass
        # through fp16 scratch with a static scale (runtime.GRAD_SCALE): an overflow there would otherwise poison w, m and
        # v for good. Every step checks the (all-reduced) flat gradient on the device; a flagged call changes nothing and
        # does not count (the semantics of torch.cuda.amp.GradScaler, without the host round trip).
        self.t_dev = torch.zeros(4, dtype=torch.int32, device=dev)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self._lr_host = float(lr)
        # fused=True: the ~10 head tensors in 2 launches instead of the 16 multi_tensor_apply launches of the foreach form
        # (0.14 ms at the serial end of every step in the graph-replay timeline, profiles/r2e_timeline_graph.json)
        self._rest_opt = torch.optim.AdamW(self.rest, lr=torch.tensor(float(lr), device=dev), betas=betas, eps=eps,
                                           weight_decay=weight_decay, capturable=True, fused=True)
        self.t = 0

    def sync_lr(self):
        """param_groups[0]['lr'] (driven by LR schedulers) -> the device words. A no-op while the value is unchanged;
        must run OUTSIDE a graph capture (GraphedStep calls it before every replay)."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self.lr_dev.fill_(lr)
            for rg in self._rest_opt.param_groups:
                rg["lr"].fill_(lr)
            self._lr_host = lr

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        self.t += 1
        self.t_dev[:1].add_(1)
        fp = self.fp
        if fp.grads_fresh:
            ops.grad_nonfinite(fp.flat_g[: self.n_live], self.t_dev)
            ops.adamw_step_dev(fp.flat_w[: self.n_live], fp.flat_g[: self.n_live], self.m, self.v, self.lr_dev,
                               g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self.t_dev)
            fp.grads_fresh = False
        self._rest_opt.step()
        return None

    def steps_taken(self):
        """(optimizer steps applied, calls skipped because of a non-finite gradient) -- reads the device words (a sync)."""
        calls, skipped = (int(v) for v in self.t_dev[:2].tolist())
        return calls - skipped, skipped

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none=True)

    # -- checkpointing: reference builder/utils/logger.py:167 saves optimizer.state_dict() into every checkpoint --------
    def state_dict(self):
        """torch.optim.Optimizer.state_dict() plus the flat moments, the step count and the head optimizer's state (the
        base class only knows `self.state`, which this optimizer does not use)."""
        sd = super().state_dict()
        sd["flat"] = {"m": self.m.detach().clone(), "v": self.v.detach().clone(), "t": self.steps_taken()[0],
                      "n_live": int(self.n_live), "rest": self._rest_opt.state_dict()}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        flat = state_dict.pop("flat", None)
        super().load_state_dict(state_dict)
        if flat is None:
            return
        if int(flat["n_live"]) != self.n_live:
            raise ValueError(f"FlatAdamW: checkpoint holds {flat['n_live']} live parameters, model has {self.n_live}")
        with torch.no_grad():
            self.m.copy_(flat["m"])
            self.v.copy_(flat["v"])
            self.t = int(flat["t"])
            self.t_dev.zero_()
            self.t_dev[:1].fill_(self.t)
        self._rest_opt.load_state_dict(flat["rest"])
        lr = float(self.param_groups[0]["lr"])
        self.lr_dev.fill_(lr)
        for rg in self._rest_opt.param_groups:          # capturable AdamW keeps lr as a device tensor
            if torch.is_tensor(rg["lr"]):
                rg["lr"] = rg["lr"].to(self.lr_dev.device)
                rg["lr"].fill_(lr)
            else:
                rg["lr"] = torch.tensor(lr, device=self.lr_dev.device)
        self._lr_host = lr
