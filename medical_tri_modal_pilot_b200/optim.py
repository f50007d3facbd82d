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
        # (trainer.GraphedStep). The few head parameters (classifier, demographic branch: stock PyTorch modules whose
        # gradients come from autograd) get the SAME kernel over a second, small flat buffer: torch.optim.AdamW in its
        # capturable form cost 16 foreach launches (0.14 ms) or one 78 us fused launch at the serial end of every step.
        # t_dev = {step() calls, calls skipped, last call whose gradient was non-finite}. Gradients of the 16-bit plan pass
        # through fp16 scratch with a static scale (runtime.GRAD_SCALE): an overflow there would otherwise poison w, m and
        # v for good. Every step checks the (all-reduced) flat gradient on the device; a flagged call changes nothing and
        # does not count (the semantics of torch.cuda.amp.GradScaler, without the host round trip).
        self.t_dev = torch.zeros(4, dtype=torch.int32, device=dev)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self._lr_host = float(lr)
        self._rest_live = None        # indices into self.rest of the parameters that receive gradients (set at the first step)
        self._rest_pending = {}       # index -> (m, v) restored by load_state_dict before the flat buffers exist
        self.rest_w = self.rest_g = self.rest_m = self.rest_v = None
        self._rest_views = []
        self.t = 0

    # -- head parameters: one flat fp32 buffer, parameters re-pointed at views of it ------------------------------------
    def _build_rest(self, live_idx):
        """(Re)build the flat buffers for the head parameters that receive gradients. Parameters whose gradient is None are
        skipped entirely, like torch.optim.AdamW does (no weight decay either). Allocates: must not happen inside a graph
        capture (GraphedStep's eager warm-up steps run first)."""
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("FlatAdamW: the set of head parameters with gradients changed inside a CUDA-graph capture")
        dev = self.t_dev.device
        old = {}
        if self._rest_live is not None:
            for i, (a, b) in zip(self._rest_live, self._rest_views):
                old[i] = (self.rest_m[a:b].clone(), self.rest_v[a:b].clone())
        old.update(self._rest_pending)
        self._rest_pending = {}
        views, off = [], 0
        for i in live_idx:
            n = self.rest[i].numel()
            views.append((off, off + n))
            off += (n + 3) // 4 * 4                      # every tensor starts on a 16-byte boundary
        total = max(off, 4)
        self.rest_w = torch.zeros(total, dtype=torch.float32, device=dev)
        self.rest_g = torch.zeros_like(self.rest_w)
        self.rest_m = torch.zeros_like(self.rest_w)
        self.rest_v = torch.zeros_like(self.rest_w)
        with torch.no_grad():
            for i, (a, b) in zip(live_idx, views):
                p = self.rest[i]
                self.rest_w[a:b].copy_(p.data.reshape(-1))
                p.data = self.rest_w[a:b].view_as(p)      # the kernel updates the parameter in place
                if i in old:
                    self.rest_m[a:b].copy_(old[i][0].reshape(-1))
                    self.rest_v[a:b].copy_(old[i][1].reshape(-1))
        self._rest_live, self._rest_views = list(live_idx), views
        self._rest_gviews = [self.rest_g[a:b].view_as(self.rest[i]) for i, (a, b) in zip(live_idx, views)]

    def _step_rest(self, g):
        live_idx = [i for i, p in enumerate(self.rest) if p.grad is not None]
        if live_idx != self._rest_live:
            self._build_rest(live_idx)
        if not live_idx:
            return
        with torch.no_grad():             # one multi-tensor copy: autograd's per-parameter gradients -> the flat buffer
            torch._foreach_copy_(self._rest_gviews, [self.rest[i].grad for i in self._rest_live])
        ops.adamw_step_dev(self.rest_w, self.rest_g, self.rest_m, self.rest_v, self.lr_dev, g["betas"][0], g["betas"][1],
                           g["eps"], g["weight_decay"], self.t_dev, count_skip=False)

    def sync_lr(self):
        """param_groups[0]['lr'] (driven by LR schedulers) -> the device words. A no-op while the value is unchanged;
        must run OUTSIDE a graph capture (GraphedStep calls it before every replay)."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self.lr_dev.fill_(lr)
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
        self._step_rest(g)
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
        rest = {}
        if self._rest_live is not None:
            for i, (a, b) in zip(self._rest_live, self._rest_views):
                rest[i] = (self.rest_m[a:b].detach().clone(), self.rest_v[a:b].detach().clone())
        for i, mv in self._rest_pending.items():
            rest.setdefault(i, mv)
        sd["flat"] = {"m": self.m.detach().clone(), "v": self.v.detach().clone(), "t": self.steps_taken()[0],
                      "n_live": int(self.n_live), "rest_mv": rest, "rest_numel": [p.numel() for p in self.rest]}
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
        if "rest_mv" in flat:
            if list(flat.get("rest_numel", [])) != [p.numel() for p in self.rest]:
                raise ValueError("FlatAdamW: the checkpoint's head parameters do not match this model")
            dev = self.t_dev.device
            self._rest_pending = {int(i): (m.to(dev).reshape(-1), v.to(dev).reshape(-1)) for i, (m, v) in flat["rest_mv"].items()}
            if self._rest_live is not None:             # buffers already exist: apply now
                self._build_rest(self._rest_live)
        lr = float(self.param_groups[0]["lr"])
        self.lr_dev.fill_(lr)
        self._lr_host = lr
