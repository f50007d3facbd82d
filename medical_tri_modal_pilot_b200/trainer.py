"""Training / evaluation step with the reference's trainer contract (reference builder/trainer/trainer.py:20-241,
builder/trainer/__init__.py:14-47) plus the one thing the reference does not have: data-parallel gradient exchange.

`missing_trainer` / `get_trainer` take exactly the reference's arguments (host or device tensors as `2_train.py:141-200`
passes them) and return `(model, loss.item())`. Differences, all deliberate and documented in DESIGN.md:
  * no hidden device->host round trips: the `missing` -> `missing_num` code (trainer.py:68-84, a `torch.unique` over
    rows) is computed arithmetically on device (`2*img_missing + txt_missing`, bit-identical on 0/1 rows);
  * `train_x` is NOT truncated to `max(input_lengths)` (trainer.py:41-42 needs a host sync); rows past each sample's
    length are dead in the fused path (kv_len masks them), so the result is identical;
  * `torch.cuda.amp.autocast()` is not entered: the fused path has its own fp16 tensor-core precision plan and the
    tiny classifier head stays fp32.

`GradSync` is the data-parallel part: one process per GPU, the flat fp32 gradient buffer of the fused
path is all-reduced range by range (one range per encoder layer, launched from the backward as soon as that layer's
weight gradients are complete) on a dedicated communication stream, so NCCL over NVLink overlaps the rest of the
backward; the few head parameters go in one extra flat bucket.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def missing_to_num(missing: torch.Tensor) -> torch.Tensor:
    """[B,3] rows (0, img_missing, txt_missing) -> code 0..3 (reference trainer.py:68-84: rank among the 4 canonical
    rows after torch.unique(sorted) == 2*img_missing + txt_missing)."""
    m = missing.to(torch.long)
    return 2 * m[:, 1] + m[:, 2]


DDP_RESERVE_SMS_DEFAULT = 0


def ddp_reserved_sms() -> int:
    """SMs left to NCCL in data-parallel runs (env TMP_B200_DDP_RESERVE_SMS, default DDP_RESERVE_SMS_DEFAULT)."""
    import os
    return int(os.environ.get("TMP_B200_DDP_RESERVE_SMS", str(DDP_RESERVE_SMS_DEFAULT)))


def ddp_setup_env() -> int:
    """Call BEFORE `init_process_group`: caps the CTAs NCCL may use per collective (env NCCL_MAX_CTAS, read when the
    communicator is created) at the number of SMs `GradSync` keeps free of compute CTAs. Returns that number."""
    import os
    r = ddp_reserved_sms()
    if r > 0:
        os.environ.setdefault("NCCL_MAX_CTAS", str(r))
        os.environ.setdefault("NCCL_MIN_CTAS", str(min(r, 4)))
    return r


def ddp_pg_options():
    """`pg_options` for `init_process_group("nccl", ...)`: NCCL's kernels on high-priority streams. The compute kernels are
    persistent grids that hold every SM; an all-reduce launched next to them gets its CTAs scheduled one by one as SMs
    drain, and it occupies those SMs for as long as its LAST channel is still waiting for one. High priority lets the block
    scheduler place NCCL's CTAs first (env TMP_B200_DDP_HIGH_PRIO=0 turns it off for A/B runs)."""
    import os
    if os.environ.get("TMP_B200_DDP_HIGH_PRIO", "1") == "0":
        return None
    opts = dist.ProcessGroupNCCL.Options()
    opts.is_high_priority_stream = True
    return opts


class GradSync:
    """Bucketed gradient all-reduce overlapped with the fused backward (NCCL, or gloo in the CPU tests).

    Attach with `GradSync(model, process_group)`; `missing_trainer` calls `finish()` between backward and
    optimizer.step(). Averaging: every range is pre-divided by world_size on the compute stream, then summed.

    `reserve_sms` (default: `ddp_reserved_sms()`): SMs kept free of compute CTAs for NCCL's kernels. The compute kernels
    are persistent / one-wave grids with equal work per CTA, so a kernel launched while NCCL holds a few SMs ran its last
    CTAs as a second wave (twice its time) -- the bulk of the 0.45 ms/step that data parallelism cost from N = 2 on.
    Pair it with `ddp_setup_env()` before `init_process_group` so that NCCL does not take more CTAs than that."""

    def __init__(self, model, group=None, broadcast_from: int | None = 0, overlap: bool = True,
                 reserve_sms: int | None = None):
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group)
        self.fp = model._fused
        self.overlap = overlap
        self.comm_stream = None
        self.head_params = [p for n, p in model.named_parameters()
                            if p.requires_grad and not n.startswith("img_encoder.") and not self._is_fused(n)]
        self._head_done = set()      # ids of the head parameters whose gradient went out early (head_early)
        self.n_collectives = 0
        self.fp.comm_hook = self._on_range_ready
        self.fp.head_hook = self._head_early
        self.fp.grad_post_scale = 1.0 / self.world      # averaging rides on the un-scaling pass of every range
        self.reserve_sms = ddp_reserved_sms() if reserve_sms is None else int(reserve_sms)
        if self.world > 1 and self.reserve_sms > 0 and next(model.parameters()).is_cuda:
            from . import ops
            ops.set_reserved_sms(self.reserve_sms)
        object.__setattr__(model, "grad_sync", self)      # found by train_step; not a submodule / state_dict entry
        if broadcast_from is not None:
            self.broadcast_parameters(broadcast_from)

    def _is_fused(self, name):
        from .runtime import _is_fused_param
        return _is_fused_param(name)

    def broadcast_parameters(self, src=0):
        """Identical initial weights on every rank. Buffers (BatchNorm statistics) included."""
        with torch.no_grad():
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                dist.broadcast(t.data, src, group=self.group)

    def _comm(self, dev):
        """The communication stream, made to wait for everything issued so far on the current stream."""
        if self.comm_stream is None:
            import os
            hi = os.environ.get("TMP_B200_DDP_HIGH_PRIO", "1") != "0"
            self.comm_stream = torch.cuda.Stream(device=dev, priority=-1 if hi else 0)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.comm_stream.wait_event(ev)
        return self.comm_stream

    # called by FusedPath.backward on the compute stream once flat_g[a:b] is final (unscaled and pre-divided by world_size)
    def _on_range_ready(self, a: int, b: int):
        g = self.fp.flat_g[a:b]
        if self.world == 1:
            return
        if g.is_cuda and self.overlap:
            with torch.cuda.stream(self._comm(g.device)):
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        self.n_collectives += 1

    def _head_bucket(self, params):
        """Average the gradients of `params` over the ranks: one flat bucket, three launches around the collective."""
        grads = [p.grad for p in params]
        flat = torch.cat([g.reshape(-1) for g in grads]).mul_(1.0 / self.world)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        self.n_collectives += 1
        with torch.no_grad():
            torch._foreach_copy_(grads, [c.view_as(g) for c, g in zip(flat.split([g.numel() for g in grads]), grads)])

    # called by FusedPath.backward when dL/dCLS arrives, i.e. after autograd has run the classifier head's backward: the
    # head gradients that exist by then are averaged on the communication stream while the fused backward runs
    def _head_early(self):
        self._head_done = set()
        if self.world == 1:
            return
        ready = [p for p in self.head_params if p.grad is not None]
        if not ready:
            return
        if ready[0].grad.is_cuda:
            if not self.overlap:
                return                                   # A/B mode: everything goes out in finish(), on the compute stream
            with torch.cuda.stream(self._comm(ready[0].grad.device)):
                self._head_bucket(ready)
        else:
            self._head_bucket(ready)                     # CPU tensors (gloo tests): synchronous
        self._head_done = {id(p) for p in ready}

    def close(self):
        """Detach from the model and drop every captured step graph that references the communicator. A process group
        whose communicator is still referenced by a live CUDA graph cannot be torn down (`destroy_process_group()` blocked
        in ncclCommDestroy in the first 2-GPU runs); call this first, then destroy the group."""
        import gc
        self.fp.comm_hook = None
        self.fp.head_hook = None
        self.fp.grad_post_scale = 1.0
        cache = self.model.__dict__.pop("_graphed_steps", None)
        if cache:
            for gs in cache.values():
                gs.graph = None
                gs.loss = None
            cache.clear()
        if getattr(self.model, "grad_sync", None) is self:
            object.__setattr__(self.model, "grad_sync", None)
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def finish(self):
        """Head parameters not covered by `_head_early` + join the communication stream. Call after loss.backward(),
        before optimizer.step()."""
        if self.world == 1:
            return
        live = [p for p in self.head_params if p.grad is not None and id(p) not in self._head_done]
        self._head_done = set()
        if live:
            self._head_bucket(live)
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)


def prepare_batch(args, device, train_x, static_x, input_lengths, train_y, x_img, x_txt, txt_lengths, imgtxt_time,
                  missing):
    """Host->device moves and casts of reference 2_train.py:143-169 + trainer.py:25-105 (no-ops for tensors that are
    already resident). Returns the dict `train_step` consumes."""
    nb = dict(device=device, non_blocking=True)
    img_time, txt_time = imgtxt_time
    missing_num = missing_to_num(missing.to(**nb))
    if args.input_types == "vslt_txt":                           # trainer.py:99-105 remap (the model maps it back)
        missing_num = torch.where(missing_num >= 2, missing_num - 2, missing_num)
    elif args.input_types == "vslt_img":
        missing_num = torch.where(missing_num == 3, torch.ones_like(missing_num), torch.zeros_like(missing_num))
    static_x = static_x.to(**nb)
    return dict(
        x=train_x.to(**nb), y=train_y.to(dtype=torch.float32, **nb),
        age=static_x[:, 1].float(), gender=static_x[:, 0].float(),           # trainer.py:92-95
        input_lengths=input_lengths.to(**nb), txts=x_txt.to(**nb), txt_lengths=txt_lengths.to(**nb),
        img=x_img.to(**nb), missing_num=missing_num,
        img_time=img_time.to(dtype=torch.float32, **nb),                     # reference: HalfTensor (:26-27)
        txt_time=txt_time.to(dtype=torch.float32, **nb))


def forward_loss(args, model, criterion, b, flow_type):
    mean = getattr(args, "feature_means", None)
    output, _, _ = model(b["x"], None, None, None, mean, b["age"], b["gender"], b["input_lengths"], b["txts"],
                         b["txt_lengths"], b["img"], b["missing_num"], None, b["img_time"], b["txt_time"], flow_type,
                         None, None)
    output = output.squeeze()
    return output, criterion(output, b["y"])


def train_step(args, model, optimizer, criterion, b, scheduler=None, iteration=0, logger=None):
    """One optimisation step on a prepared (device-resident) batch; returns the loss as a DEVICE tensor (no sync).
    reference trainer.py:124-191."""
    optimizer.zero_grad()
    _, loss = forward_loss(args, model, criterion, b, "train")
    loss.backward()
    sync = getattr(model, "grad_sync", None)
    if sync is not None:
        sync.finish()
    optimizer.step()
    if scheduler is not None:
        scheduler.step(iteration)
        if logger is not None:
            logger.log_lr(scheduler.get_lr()[0], iteration)
    return loss.detach()


_RAW_KEYS = ("train_x", "static_x", "input_lengths", "train_y", "x_img", "x_txt", "txt_lengths", "img_time", "txt_time",
             "missing")


class GraphedStep:
    """The whole optimisation step (casts of `prepare_batch`, frozen image encoder, fused encoder forward + backward on
    three CUDA streams, classifier head + its autograd, gradient all-reduce, AdamW) captured ONCE as a CUDA graph and
    replayed per iteration. The eager step issues ~450 launches and costs ~14 ms of host time against ~15 ms of device
    time at the bench workload (tools/step_breakdown.py): it is launch-bound as soon as the kernels get faster.

    Everything that varies between steps is data in HBM: the batch (static input buffers, refilled by `load`), the
    dropout step counter (`FusedPath.step_dev`), AdamW's step count and learning rate (`FlatAdamW.t_dev / lr_dev`).
    Calls 1..WARMUP with a given input signature run eagerly on the capture stream (lazy initialisation: workspaces,
    cuBLAS handles, autograd buffers), the next call captures, later calls replay."""

    WARMUP = 2

    def __init__(self, args, model, optimizer, criterion, raw):
        self.args, self.model, self.optimizer, self.criterion = args, model, optimizer, criterion
        self.key = self.signature(raw, model, optimizer)
        dev = next(model.ie_vslt.parameters()).device
        self.device = dev
        self.static = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in raw.items()}
        self.stream = torch.cuda.Stream(device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        # "external" events: captured as event-wait NODES that refer to the real event (recorded by load() on the copy stream
        # before every replay), not as capture-internal dependencies
        from .swin_feed import SwinFeed
        x_img = raw["x_img"]
        n_images = x_img.numel() // (224 * 224) if x_img.shape[-1] == 224 else 0
        k = SwinFeed.n_chunks(n_images) if n_images else 1
        ev = lambda: torch.cuda.Event(external=True)
        self.ready = {"small": ev(), "txt": ev(), "img": [ev() for _ in range(k)]}
        self._loaded = False
        self.graph = None
        self.loss = None
        self.calls = 0
        self.launches_per_replay = 0

    @staticmethod
    def signature(raw, model, optimizer):
        return (tuple((k, tuple(v.shape), v.dtype) for k, v in raw.items()), id(optimizer), bool(model.training))

    def load(self, raw):
        """Host (pinned) or device tensors of one batch -> the static input buffers, as a STAGED asynchronous upload on a
        copy stream: the small tensors (1 MB: the vslt lane starts at once), pixels of image chunk 0, the remaining image
        chunks, the text embeddings -- each group followed by an event that the captured step waits for exactly where it
        first touches that data (FusedPath.forward / SwinFeed). The image encoder of chunk c therefore overlaps the upload
        of everything behind it, and the vslt lane overlaps all of it: `e2e` is within 1.5 % of the device-resident step
        (10.70 vs 10.56 ms; pixels first: 10.75-10.87; env TMP_B200_UPLOAD_SMALL_FIRST=0 restores that order).
        Reference: 2_train.py:143-169 issues the same copies up front, in the compute stream."""
        cs = self.copy_stream
        cs.wait_stream(torch.cuda.current_stream())       # the previous step may still be reading the static buffers
        img_src, img_dst = raw["x_img"], self.static["x_img"]
        n_flat = img_dst.numel()
        k = len(self.ready["img"])
        bounds = [n_flat * c // k for c in range(k + 1)]

        def put(key):
            v = raw[key]
            if v.data_ptr() != self.static[key].data_ptr():
                self.static[key].copy_(v, non_blocking=True)

        def put_img(c):
            if img_src.data_ptr() == img_dst.data_ptr():
                return
            if img_src.is_contiguous() and img_src.dtype == img_dst.dtype:
                img_dst.view(-1)[bounds[c]:bounds[c + 1]].copy_(img_src.view(-1)[bounds[c]:bounds[c + 1]], non_blocking=True)
            elif c == 0:
                img_dst.copy_(img_src, non_blocking=True)

        import os
        small_first = os.environ.get("TMP_B200_UPLOAD_SMALL_FIRST", "1") != "0"
        with torch.cuda.stream(cs):
            if small_first:
                for key in raw:
                    if key not in ("x_img", "x_txt"):
                        put(key)
                self.ready["small"].record(cs)
            put_img(0)
            self.ready["img"][0].record(cs)
            if not small_first:
                for key in raw:
                    if key not in ("x_img", "x_txt"):
                        put(key)
                self.ready["small"].record(cs)
            for c in range(1, k):
                put_img(c)
                self.ready["img"][c].record(cs)
            put("x_txt")
            self.ready["txt"].record(cs)
        self._loaded = True

    def _eager(self):
        s = self.static
        fp = self.model._fused
        fp.input_ready = self.ready
        try:
            torch.cuda.current_stream().wait_event(self.ready["small"])
            b = prepare_batch(self.args, self.device, s["train_x"], s["static_x"], s["input_lengths"], s["train_y"], s["x_img"],
                              s["x_txt"], s["txt_lengths"], (s["img_time"], s["txt_time"]), s["missing"])
            return train_step(self.args, self.model, self.optimizer, self.criterion, b)
        finally:
            fp.input_ready = None

    def step(self, scheduler=None, iteration=0, logger=None):
        """One optimisation step on the batch currently in the static buffers; returns the loss (device tensor)."""
        from . import _lib
        self.calls += 1
        if not self._loaded:
            raise RuntimeError("GraphedStep.step() before load(): the static input buffers are empty")
        if hasattr(self.optimizer, "sync_lr"):
            self.optimizer.sync_lr()
        if self.graph is None:
            cur = torch.cuda.current_stream()
            self.stream.wait_stream(cur)
            if self.calls <= self.WARMUP:
                with torch.cuda.stream(self.stream):
                    self.loss = self._eager()
                cur.wait_stream(self.stream)
            else:
                self.optimizer.zero_grad(set_to_none=True)
                torch.cuda.synchronize()
                n0 = _lib.launch_count
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.stream):
                    self.loss = self._eager()
                self.launches_per_replay = _lib.launch_count - n0
                _lib.launch_count = n0
                self.graph = g
                # the graph holds raw pointers into the fused-path / image-encoder workspaces of THIS shape: keep those
                # objects alive for as long as the graph exists (the per-shape caches are LRUs and may drop them)
                swin = self.model.__dict__.get("_swin_native")
                self._pinned = (self.model._fused.pin(), swin[1].pin() if swin else None)
        if self.graph is not None:
            self.graph.replay()
            _lib.launch_count += self.launches_per_replay
        if scheduler is not None:
            scheduler.step(iteration)
            if logger is not None:
                logger.log_lr(scheduler.get_lr()[0], iteration)
        return self.loss


def graphs_enabled(args, optimizer, scaler=None) -> bool:
    """CUDA-graph replay of the train step: on by default (`args.cuda_graph` / env TMP_B200_GRAPH=0 turn it off) when
    the optimizer keeps its per-step scalars on the device (optim.FlatAdamW)."""
    import os
    flag = getattr(args, "cuda_graph", None)
    if flag is None:
        flag = os.environ.get("TMP_B200_GRAPH", "1") != "0"
    if getattr(getattr(optimizer, "fp", None), "precision", "fp16") != "fp16":
        return False                       # the fp32 parity mode runs eagerly
    return bool(flag) and scaler is None and hasattr(optimizer, "sync_lr")


def graphed_step(args, model, optimizer, criterion, raw) -> GraphedStep:
    """The model's cached GraphedStep for this input signature (one per signature; a new signature = new capture)."""
    cache = model.__dict__.setdefault("_graphed_steps", {})
    key = GraphedStep.signature(raw, model, optimizer)
    gs = cache.get(key)
    if gs is None:
        if len(cache) >= 4:                       # e.g. a ragged last batch every epoch: keep the cache bounded
            cache.pop(next(iter(cache)))
        gs = cache[key] = GraphedStep(args, model, optimizer, criterion, raw)
    return gs


def missing_trainer(args, iteration, train_x, static_x, input_lengths, train_y, model, logger, device, scheduler=None,
                    optimizer=None, criterion=None, scaler=None, flow_type=None, output_lengths=None, seq_lengths=None,
                    x_img=None, x_txt=None, txt_lengths=None, imgtxt_time=None, missing=None, reports_tokens=None,
                    reports_lengths=None, criterion_aux=None):
    """Same signature and return value as reference builder/trainer/trainer.py:20-241 (TIE branch)."""
    if getattr(args, "vslt_type", "TIE") != "TIE":
        raise NotImplementedError("B200 trainer implements --vslt-type TIE")
    if flow_type == "train" and graphs_enabled(args, optimizer, scaler) and torch.device(device).type == "cuda":
        raw = dict(zip(_RAW_KEYS, (train_x, static_x, input_lengths, train_y, x_img, x_txt, txt_lengths, imgtxt_time[0],
                                   imgtxt_time[1], missing)))
        gs = graphed_step(args, model, optimizer, criterion, raw)
        gs.load(raw)
        return model, gs.step(scheduler, iteration, logger).item()
    b = prepare_batch(args, device, train_x, static_x, input_lengths, train_y, x_img, x_txt, txt_lengths, imgtxt_time,
                      missing)
    if flow_type == "train":
        loss = train_step(args, model, optimizer, criterion, b, scheduler, iteration, logger)
    else:
        with torch.no_grad():
            output, loss = forward_loss(args, model, criterion, b, flow_type)
            output = torch.sigmoid(output)
        if logger is not None:
            logger.evaluator.add_batch(b["y"], output)
    return model, loss.item()


def get_trainer(args, iteration, x, static, input_lengths, y, output_lengths, model, logger, device, scheduler,
                optimizer, criterion, x_txt=None, x_img=None, txt_lengths=None, seq_lengths=None, imgtxt_time=None,
                scaler=None, missing=None, flow_type=None, reports_tokens=None, reports_lengths=None,
                criterion_aux=None):
    """reference builder/trainer/__init__.py:14-47"""
    return missing_trainer(args, iteration, x, static, input_lengths, y, model, logger, device, scheduler, optimizer,
                           criterion, scaler, flow_type, output_lengths, seq_lengths=seq_lengths, x_img=x_img,
                           x_txt=x_txt, txt_lengths=txt_lengths, imgtxt_time=imgtxt_time, missing=missing,
                           reports_tokens=reports_tokens, reports_lengths=reports_lengths, criterion_aux=criterion_aux)
