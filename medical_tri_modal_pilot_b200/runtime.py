"""FusedPath -- host-side orchestration of the sm_100a kernels for the UMSE embedding + MBT fusion encoder of
`tri_mbt_vsltcls` (reference tri_mbt_vsltcls.py:183-240 and mbt_encoder.py:696-784), forward and backward.

Memory plan (all device buffers are allocated once and reused; nothing is allocated inside a step):
  * every fused-path parameter lives in ONE flat fp32 buffer (`flat_w`, q/k/v of a block adjacent so QKV is a single
    [768,256] GEMM operand); the nn.Parameters of the model are views into it, so optimizers / state_dict see the
    reference's names and shapes. Gradients live in a second flat buffer with the same offsets (`flat_g`): the
    wgrad / reduction kernels accumulate straight into it and `param.grad` are views -> the DDP allreduce walks
    contiguous ranges of `flat_g` without any flatten/copy.
  * fp16 copies of the parameters (`flat_w16`) and transposed copies of the GEMM weights (`flat_wT16`, dgrad
    operands) are refreshed by two kernel launches per step.
  * activations: stream m keeps X[l] (layer inputs, l = 0..n_layers) plus per-layer xn, qkv, O, h, hn, a, lse as
    padded [B, T_m, *] fp16 tensors (gradient scratch is fp16 too, scaled by GRAD_SCALE) (T_v = 5+L, T_i = 5+49*n_img, T_t = 133: 4 bottleneck rows, CLS, tokens).
    Key-padding is never materialised: kv_len[3,B] int32 stays on device.
"""
from __future__ import annotations

import contextlib
import os

import numpy as np
import torch

from . import ops

D = 256
FF = 1024
ACT = ops.ACT    # fp16: forward activations and 16-bit weight copies
GRD = ops.GRD    # fp16: gradient tensors, multiplied by GRAD_SCALE
# Static power-of-two scale carried by every 16-bit gradient tensor (fp16 has 5 exponent bits; dS ~ P*dP/8 would sit
# in the subnormal range unscaled). Backward kernels are linear in the incoming gradient, so the scale is applied
# once to dL/dCLS and removed once from the flat fp32 gradient buffer. The reference itself runs fp16 autocast with
# NO scaling at all (trainer.py:126,185-188: GradScaler commented out).
GRAD_SCALE = 4096.0


def _is_fused_param(name: str) -> bool:
    if name.startswith(("ie_vslt.", "ie_time.", "ie_feat.", "txt_embedding.", "linear.")):
        return True
    if name.startswith("fusion_transformer.") and "layer_norms_after_concat" not in name:
        return True
    return False


class _Block:
    """Flat-buffer views of one (layer, modality) encoder block."""
    __slots__ = ("ln1_g", "ln1_b", "wqkv", "bqkv", "ln2_g", "ln2_b", "w1", "b1", "w2", "b2")


class FusedPath:
    def __init__(self, model):
        self.model = model
        self.device = None
        self.step = 0
        self._ws_cache = {}       # (B, L, n_img) -> workspace dict (small LRU; captured graphs pin theirs, see pin())
        self._ctx_token = 0
        self.comm_hook = None     # optional callable(a, b): flat_g[a:b] is final (trainer.GradSync all-reduces it)
        self.head_hook = None     # optional callable(): dL/dCLS has arrived, the classifier head's backward has run
        self.grad_post_scale = 1.0   # extra factor folded into the un-scaling pass (GradSync: 1 / world_size)
        self.skip_missing = True
        self.fuse_grad_dropout = os.environ.get("TMP_B200_FUSE_GRAD_DROPOUT", "1") != "0"   # A/B switch
        self.cls_attn_bwd_fast = os.environ.get("TMP_B200_CLS_ATTN_BWD_GENERIC") is None    # A/B switch
        self.grads_fresh = False  # set by backward(), cleared by optim.FlatAdamW.step()
        self.debug_trace = None   # dict -> backward() records per-layer input gradients (debugging aid)
        # img / txt modality streams on side CUDA streams (TMP_B200_SINGLE_STREAM=1 serialises them: debugging aid)
        self.multi_stream = os.environ.get("TMP_B200_SINGLE_STREAM", "0") != "1"
        self.grad_scale = GRAD_SCALE
        # "fp16": the tensor-core plan above. "fp32": north-star's FP32 parity mode -- every activation / gradient tensor in
        # fp32, GEMM operands split into bf16x3 for the same tcgen05 kernels, attention on the CUDA cores (csrc/precise.cu):
        # logits within 1e-3 of the reference and every parameter gradient cosine >= 0.999 end to end. Chosen by
        # `args.precision` / env TMP_B200_PRECISION (model.py); ~10x slower, for validation not throughput.
        self.precision = "fp16"
        self.input_ready = None   # set by trainer.GraphedStep around a step whose inputs are still being uploaded
        # last fused layer under --mbt-only-vslt 1: only the vslt CLS row of its output reaches the classifier
        # (mbt_encoder.py:757-763, tri_mbt_vsltcls.py:248), so the rows the reference computes and then drops are not
        # computed here: attention for the first query tile only, LayerNorm2 / FFN on the B CLS rows, and the mirror image
        # in the backward (exact: those rows have no consumer and receive no gradient). False = the reference's full layer.
        self.cls_only = os.environ.get("TMP_B200_FULL_LAST_LAYER", "0") != "1"
        self.seed_base = None     # dropout seed = seed_base + step_dev (device int32 counter, see forward)
        self.step_dev = None

    # ------------------------------------------------------------------------------------------------------------
    # parameter flattening
    # ------------------------------------------------------------------------------------------------------------
    def _layout(self):
        """Ordered list of (state_dict name, param) such that tensors consumed together are adjacent."""
        m = self.model
        named = dict(m.named_parameters())
        order = []
        for p in ("ie_vslt", "ie_time"):
            order += [f"{p}.0.weight", f"{p}.0.bias", f"{p}.1.weight", f"{p}.1.bias"]
        order += ["ie_feat.weight"]
        F = "fusion_transformer"
        order += [f"{F}.cls_token_per_modality.{k}" for k in range(3)]
        order += [f"{F}.bottlenecks"]
        for k in range(3):
            order += [f"{F}.layer_norms_in.{k}.weight", f"{F}.layer_norms_in.{k}.bias"]
        order += ["txt_embedding.weight", "txt_embedding.bias", "linear.weight", "linear.bias"]
        for l in range(m.num_layers):
            for k in range(3):
                p = f"{F}.layer_stacks.{l}.{k}"
                order += [f"{p}.attention_prenorm.gamma", f"{p}.attention_prenorm.beta"]
                order += [f"{p}.self_attention.{q}_proj.linear.weight" for q in ("query", "key", "value")]
                order += [f"{p}.self_attention.{q}_proj.linear.bias" for q in ("query", "key", "value")]
                order += [f"{p}.feed_forward_prenorm.gamma", f"{p}.feed_forward_prenorm.beta"]
                order += [f"{p}.feed_forward.w_1.weight", f"{p}.feed_forward.w_1.bias",
                          f"{p}.feed_forward.w_2.weight", f"{p}.feed_forward.w_2.bias"]
        fused = {n for n in named if _is_fused_param(n)}
        assert fused == set(order), sorted(fused ^ set(order))
        return [(n, named[n]) for n in order]

    def _ensure_params(self, device):
        m = self.model
        first = next(iter(m.ie_vslt.parameters()))
        if self.device == device and getattr(self, "_sig", None) == first.data_ptr():
            return
        if device.type != "cuda":
            raise RuntimeError("FusedPath needs a CUDA device")
        layout = self._layout()
        offs, total = {}, 0
        for n, p in layout:
            offs[n] = total
            total += p.numel()
        assert all(o % 256 == 0 for o in offs.values())
        flat_w = torch.empty(total, dtype=torch.float32, device=device)
        flat_g = torch.zeros(total, dtype=torch.float32, device=device)
        self.gviews = {}
        with torch.no_grad():
            for n, p in layout:
                v = flat_w[offs[n]: offs[n] + p.numel()].view(p.shape)
                v.copy_(p.data.to(device))
                p.data = v
                self.gviews[n] = flat_g[offs[n]: offs[n] + p.numel()].view(p.shape)
        self.layout, self.offs, self.total = layout, offs, total
        self.flat_w, self.flat_g = flat_w, flat_g
        self.flat_w16 = torch.empty(total, dtype=ACT, device=device)
        W = lambda n, *shape: flat_w[offs[n]:].as_strided(shape, _contig_strides(shape))
        G = lambda n, *shape: flat_g[offs[n]:].as_strided(shape, _contig_strides(shape))
        H16 = lambda n, *shape: self.flat_w16[offs[n]:].as_strided(shape, _contig_strides(shape))
        self.W, self.G, self.H16 = W, G, H16
        # transposed fp16 copies of the GEMM weights used by dgrad
        F = "fusion_transformer"
        t_total = m.num_layers * 3 * (768 * 256 + 1024 * 256 * 2)
        self.flat_wT16 = torch.empty(t_total, dtype=ACT, device=device)
        descs = np.zeros(1 + m.num_layers * 9, dtype=np.dtype([("src", "<u8"), ("dst", "<u8"), ("dst_t", "<u8"),
                                                               ("R", "<i4"), ("C", "<i4")]))
        descs[0] = (flat_w.data_ptr(), self.flat_w16.data_ptr(), 0, total // 256, 256)
        self.wT = {}
        t_off, k = 0, 1
        for l in range(m.num_layers):
            for s in range(3):
                p = f"{F}.layer_stacks.{l}.{s}"
                for key, name, R, C in (("qkv", f"{p}.self_attention.query_proj.linear.weight", 768, 256),
                                        ("w1", f"{p}.feed_forward.w_1.weight", 1024, 256),
                                        ("w2", f"{p}.feed_forward.w_2.weight", 256, 1024)):
                    dst_t = self.flat_wT16[t_off: t_off + R * C].view(C, R)
                    self.wT[(l, s, key)] = dst_t
                    descs[k] = (flat_w.data_ptr() + 4 * offs[name], 0, dst_t.data_ptr(), R, C)
                    t_off += R * C
                    k += 1
        self.n_desc = k
        self.cast_descs = torch.from_numpy(descs.view(np.uint8).copy()).to(device)
        self.blocks = {}
        for l in range(m.num_layers):
            for s in range(3):
                self.blocks[(l, s)] = self._block_views(l, s)
        self.device = device
        self._sig = first.data_ptr()
        self._trigger = torch.zeros(1, device=device, requires_grad=True)

    def _block_views(self, l, s):
        p = f"fusion_transformer.layer_stacks.{l}.{s}"
        names = dict(ln1_g=f"{p}.attention_prenorm.gamma", ln1_b=f"{p}.attention_prenorm.beta",
                     wqkv=f"{p}.self_attention.query_proj.linear.weight",
                     bqkv=f"{p}.self_attention.query_proj.linear.bias",
                     ln2_g=f"{p}.feed_forward_prenorm.gamma", ln2_b=f"{p}.feed_forward_prenorm.beta",
                     w1=f"{p}.feed_forward.w_1.weight", b1=f"{p}.feed_forward.w_1.bias",
                     w2=f"{p}.feed_forward.w_2.weight", b2=f"{p}.feed_forward.w_2.bias")
        shapes = dict(ln1_g=(D,), ln1_b=(D,), wqkv=(768, D), bqkv=(768,), ln2_g=(D,), ln2_b=(D,), w1=(FF, D),
                      b1=(FF,), w2=(D, FF), b2=(D,))
        w, g, h = _Block(), _Block(), _Block()
        for k, n in names.items():
            setattr(w, k, self.W(n, *shapes[k]))
            setattr(g, k, self.G(n, *shapes[k]))
            setattr(h, k, self.H16(n, *shapes[k]))
        return w, g, h

    # ------------------------------------------------------------------------------------------------------------
    # workspace
    # ------------------------------------------------------------------------------------------------------------
    def _ensure_workspace(self, B, L, n_img):
        """Activations / gradient scratch for one input shape. A captured CUDA graph (trainer.GraphedStep) bakes the raw
        device pointers of these buffers, so a workspace is never freed behind a graph's back: workspaces live in a
        small LRU keyed by shape (eager calls with ever-changing shapes cannot accumulate more than 3), and every graph
        additionally pins the workspace object it captured (`pin()`), which keeps the memory alive for as long as the
        graph exists even after the LRU dropped it."""
        f32 = self.precision == "fp32"
        ACT = GRD = torch.float32 if f32 else ops.ACT     # noqa: N806 (shadow the module-level dtypes for this workspace)
        key = (B, L, n_img, self.precision)
        hit = self._ws_cache.pop(key, None)
        if hit is not None:
            self._ws_cache[key] = hit
            self.ws, self.proj, self.g_proj, self.T = hit["ws"], hit["proj"], hit["g_proj"], hit["T"]
            self._ws_cur = hit
            return
        dev = self.device
        NL = self.model.num_layers
        T = [5 + L, 5 + 49 * n_img, 5 + 128]
        self.T = T
        ws = []
        for s in range(3):
            M = B * T[s]
            Tl = ops.lse_len(T[s])
            nl_s = NL if (s == 0 or not self.model.vsltonly) else NL - 1      # img/txt skip the last layer
            e = lambda *shape, dt=ACT: torch.empty(*shape, dtype=dt, device=dev)
            g = lambda *shape: torch.empty(*shape, dtype=GRD, device=dev)
            st = {
                "M": M, "T": T[s], "Tl": Tl, "n_layers": nl_s,
                "X": [e(B, T[s], D) for _ in range(nl_s + 1)],
                "xn": [e(M, D) for _ in range(nl_s)], "qkv": [e(M, 768) for _ in range(nl_s)],
                "O": [e(M, D) for _ in range(nl_s)], "h": [e(M, D) for _ in range(nl_s)],
                "hn": [e(M, D) for _ in range(nl_s)], "a": [e(M, FF) for _ in range(nl_s)],
                # ReLU / dropout pattern of the FFN hidden activation, 1 bit per element (gate of the FFN2 input gradient)
                "am": [torch.empty(M, FF // 32, dtype=torch.int32, device=dev) for _ in range(nl_s)],
                # rows [T, Tl) of lse (like those of delta below) are never written by the kernels but ARE loaded by the
                # attention backward together with the last query tile, where they meet masked (-inf) scores: they must be
                # finite, so the buffers start as zeros (torch.empty handed back NaN patterns from earlier allocations:
                # sporadic NaN gradients in a long-lived process)
                "lse": [torch.zeros(B, 4, Tl, dtype=torch.float32, device=dev) for _ in range(nl_s)],
                # backward scratch (reused by every layer of the stream)
                "g_y": g(B, T[s], D), "g_x": g(B, T[s], D), "g_yd": g(B, T[s], D), "g_a": g(M, FF), "g_hn": g(M, D),
                "g_h": g(M, D), "g_qkv": g(M, 768), "g_xn": g(M, D),
                # delta rows in [T, Tl) are never written and must stay finite (they meet masked, exactly-zero P entries)
                "delta": torch.zeros(B, 4, Tl, dtype=torch.float32, device=dev),
            }
            if s == 0:
                # compact CLS-row buffers of the last fused layer (--mbt-only-vslt 1: only the vslt CLS row of its output
                # is consumed, so its LayerNorm2 / FFN run on B rows instead of B*T, see _layer_fwd_cls)
                st.update({"c_x": e(B, D), "c_o": e(B, D), "c_h": e(B, D), "c_hn": e(B, D), "c_a": e(B, FF), "c_y": e(B, D),
                           "c_am": torch.empty(B, FF // 32, dtype=torch.int32, device=dev),
                           "c_gy": g(B, D), "c_gyd": g(B, D), "c_ga": g(B, FF), "c_ghn": g(B, D), "c_gh": g(B, D)})
            ws.append(st)
        self.ws = ws
        self.proj = [None, torch.empty(B * 49 * n_img, D, dtype=ACT, device=dev),
                     torch.empty(B * 128, D, dtype=ACT, device=dev)]
        self.g_proj = [None, torch.empty(B * 49 * n_img, D, dtype=GRD, device=dev),
                       torch.empty(B * 128, D, dtype=GRD, device=dev)]
        self._ws_cur = self._ws_cache[key] = dict(ws=ws, proj=self.proj, g_proj=self.g_proj, T=T)
        while len(self._ws_cache) > 3:
            self._ws_cache.pop(next(iter(self._ws_cache)))

    def pin(self):
        """Everything a captured graph of the current step references by raw pointer (workspace, cast inputs, lengths)."""
        return (self._ws_cur, getattr(self, "ctx", None))

    # ------------------------------------------------------------------------------------------------------------
    def __call__(self, x, input_lengths, txts, txt_lengths, img, img_time, txt_time, missing):
        """img: raw pixels [B,(n_img,)1,224,224] (encoded here, on the img lane, by model.encode_images) or precomputed
        encoder features [B*n_img,49,768] (tests)."""
        self._ensure_params(x.device)
        return _FusedFn.apply(self._trigger, self, x, input_lengths, txts, txt_lengths, img, img_time, txt_time, missing)

    # ------------------------------------------------------------------------------------------------------------
    def _branch(self, prefix, grads=False):
        src = self.G if grads else self.W
        return [src(f"{prefix}.0.weight", D), src(f"{prefix}.0.bias", D), src(f"{prefix}.1.weight", D),
                src(f"{prefix}.1.bias", D)]

    def _prologue_args(self, s, ctx):
        m = self.model
        F = "fusion_transformer"
        B = ctx["B"]
        n = self.T[s] - 5
        pe = m.fusion_transformer.positional_encoding.pe[0] if s == 2 else None
        common = dict(tim4=self._branch("ie_time"), Wfeat=self.W("ie_feat.weight", 20, D),
                      cls=self.W(f"{F}.cls_token_per_modality.{s}", D), bottlenecks=self.W(f"{F}.bottlenecks", 4, D),
                      ln_g=self.W(f"{F}.layer_norms_in.{s}.weight", D), ln_b=self.W(f"{F}.layer_norms_in.{s}.bias", D),
                      pe=pe, drop_p=ctx["p"], seed=ctx["seed"], salt=1000 + s, seed_dev=ctx["seed_dev"])
        if s == 0:
            return dict(kind=0, B=B, n=n, x=ctx["x"], val4=self._branch("ie_vslt"), proj=None, times=None, n_slots=0,
                        feat_id=0, **common)
        times = ctx["img_time"] if s == 1 else ctx["txt_time"]
        return dict(kind=1, B=B, n=n, x=None, val4=None, proj=self.proj[s], times=times,
                    n_slots=(ctx["n_img"] if s == 1 else 1), feat_id=(18 if s == 1 else 19), **common)

    def forward(self, x, input_lengths, txts, txt_lengths, img, img_time, txt_time, missing, training):
        m = self.model
        B, L = x.shape[0], x.shape[1]
        # staged upload (trainer.GraphedStep.load): events that fire when the small tensors / the text embeddings / each
        # chunk of pixels have landed in their device buffers; None = everything is already resident
        ready = self.input_ready or {}
        cur = torch.cuda.current_stream()
        if "small" in ready:
            cur.wait_event(ready["small"])
        n_img = 3 if m.multiimages == 1 else 1
        self._ensure_workspace(B, L, n_img)
        NL = m.num_layers
        p = m.dropout if training else 0.0
        # dropout masks are a function of (seed_base + device step counter, salt, element): the counter lives in HBM and
        # is bumped by a device op, so a captured CUDA graph of the step draws fresh masks on every replay
        self.step += 1
        if self.seed_base is None:
            # independent dropout streams per data-parallel rank (the masks are a pure function of the seed)
            rank = 0
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                rank = torch.distributed.get_rank()
            self.seed_base = ((int(torch.initial_seed()) * 1000003) ^ (rank * 0x9E3779B1)) & 0x7FFFFFFF
            self.step_dev = torch.zeros(1, dtype=torch.int32, device=x.device)
        if training:
            self.step_dev.add_(1)
        ctx = dict(B=B, L=L, n_img=n_img, p=p, seed=self.seed_base, seed_dev=self.step_dev)
        ctx["img_time"] = img_time.float().reshape(B, n_img).contiguous()
        ctx["missing"] = missing.to(torch.long).contiguous()
        f32 = self.precision == "fp32"
        ctx["f32"] = f32
        adt = torch.float32 if f32 else ACT
        Wop = self.W if f32 else self.H16       # GEMM weight operand: fp32 master (split in ops.gemm) or the fp16 copy
        # The three modality streams of a layer are independent until the bottleneck exchange (mbt_encoder.py:744-776):
        # vslt runs on the caller's stream, img / txt on two side streams, joined at every exchange. The frozen image
        # encoder runs at the head of the img lane. It is the critical path of the forward pass (5 ms of kernels; the other
        # lanes wait for it at the first exchange) and needs nothing but the pixels, `missing` and `img_time`, so the lane is
        # forked BEFORE the per-step preparation of the other lanes (input casts, kv_len, the fp16 weight refresh: 0.18 ms
        # that used to run in front of it); `_ev_prep` marks the point from which the refreshed fp16 weights may be read.
        self._fork()
        with self._lane(1):
            img_feats = m.encode_images(img, missing=ctx["missing"] if self.skip_missing else None, ready=ready.get("img"),
                                        img_time=ctx["img_time"])
            ctx["img16"] = img_feats.reshape(B * 49 * n_img, 768).to(adt).contiguous()
        ctx["x"] = x.float().contiguous()
        ctx["txt_time"] = txt_time.float().contiguous()
        T = self.T
        ctx["kv_len"] = ops.build_lengths(input_lengths.to(torch.long).contiguous(), txt_lengths.to(torch.long).contiguous(),
                                          ctx["img_time"], n_img, m.multiimages, ctx["missing"], int(self.skip_missing),
                                          T[0], T[1], T[2])
        if not f32:
            # refresh fp16 (+ transposed) parameter copies
            ops.cast_weights(self.cast_descs, 1, self.total // 256, 256)              # flat fp32 -> fp16
            ops.cast_weights(self.cast_descs[32:], self.n_desc - 1, 1024, 1024)        # transposed GEMM weights
        self._prep_done()
        # 768 -> 256 projections of the text tokens and image patches (tri_mbt_vsltcls.py:200, 210-211)
        with self._lane(1):
            ops.gemm(ctx["img16"], Wop("linear.weight", D, 768), out=self.proj[1], bias=self.W("linear.bias", D))
            ops.stream_prologue_fwd(X0=self.ws[1]["X"][0], **self._prologue_args(1, ctx))
        with self._lane(2):
            if "txt" in ready:
                torch.cuda.current_stream().wait_event(ready["txt"])
            ctx["txts16"] = txts.reshape(B * 128, 768).to(adt).contiguous()
            ops.gemm(ctx["txts16"], Wop("txt_embedding.weight", D, 768), out=self.proj[2],
                     bias=self.W("txt_embedding.bias", D))
            ops.stream_prologue_fwd(X0=self.ws[2]["X"][0], **self._prologue_args(2, ctx))
        ops.stream_prologue_fwd(X0=self.ws[0]["X"][0], **self._prologue_args(0, ctx))
        for l in range(NL):
            last = m.vsltonly == 1 and l == NL - 1
            if not last:
                for s in (1, 2):
                    with self._lane(s):
                        self._layer_fwd(l, s, ctx)
            if last and self.cls_only and not f32:
                self._layer_fwd_cls(l, ctx)
            else:
                self._layer_fwd(l, 0, ctx)
            self._join()
            if l == NL - 1:
                break
            ops.bottleneck_mix_fwd(self.ws[0]["X"][l + 1], self.ws[1]["X"][l + 1], self.ws[2]["X"][l + 1], ctx["missing"])
            self._fork()
        self._ctx_token += 1
        ctx["token"] = self._ctx_token
        ctx["ws"] = self._ws_cur
        self.ctx = ctx
        return self.ws[0]["X"][NL][:, 4, :].float()

    # ------------------------------------------------------------------------------------------------------------
    # stream lanes: lane 0 = the caller's stream, lanes 1/2 = side streams for the img / txt modality streams
    # ------------------------------------------------------------------------------------------------------------
    def _lanes_init(self):
        if getattr(self, "_side", None) is None or self._side_dev != self.device:
            self._side = [torch.cuda.Stream(device=self.device) for _ in range(2)]
            self._ev_main = torch.cuda.Event()
            self._ev_prep = torch.cuda.Event()
            self._ev_side = [torch.cuda.Event() for _ in range(2)]
            self._side_dev = self.device

    def _lane(self, s):
        if s == 0 or not self.multi_stream:
            return contextlib.nullcontext()
        return torch.cuda.stream(self._side[s - 1])

    def _fork(self):
        """side lanes wait for everything issued so far on the caller's stream"""
        if not self.multi_stream:
            return
        self._lanes_init()
        self._ev_main.record()
        for st in self._side:
            st.wait_event(self._ev_main)

    def _prep_done(self):
        """side lanes wait for everything issued so far on the caller's stream (second fork point of the forward pass: the
        lanes were forked earlier and already hold work of their own)"""
        if not self.multi_stream:
            return
        self._ev_prep.record()
        for st in self._side:
            st.wait_event(self._ev_prep)

    def _join(self):
        """the caller's stream waits for everything issued so far on the side lanes"""
        if not self.multi_stream:
            return
        cur = torch.cuda.current_stream()
        for st, ev in zip(self._side, self._ev_side):
            ev.record(st)
            cur.wait_event(ev)

    def _layer_fwd(self, l, s, ctx):
        st = self.ws[s]
        w, _, h16 = self.blocks[(l, s)]
        if ctx["f32"]:
            h16 = w                       # fp32 mode: the fp32 masters are the GEMM operands (split into bf16x3 by ops.gemm)
        B, T, M = ctx["B"], st["T"], st["M"]
        p, seed, sd = ctx["p"], ctx["seed"], ctx["seed_dev"]
        x = st["X"][l]
        ops.layernorm_fwd(x, w.ln1_g, w.ln1_b, st["xn"][l])
        ops.gemm(st["xn"][l], h16.wqkv, out=st["qkv"][l], bias=w.bqkv)
        ops.attn_fwd(st["qkv"][l], ctx["kv_len"][s], B, T, st["O"][l], st["lse"][l])
        ops.layernorm_fwd(x, w.ln2_g, w.ln2_b, st["hn"][l], add=st["O"][l], sum_out=st["h"][l])
        ops.gemm(st["hn"][l], h16.w1, out=st["a"][l], bias=w.b1, relu=True, drop_p=p, seed=seed, salt=(l * 3 + s) * 4 + 1,
                 seed_dev=sd, mask_out=None if ctx["f32"] else st["am"][l])
        ops.gemm(st["a"][l], h16.w2, out=st["X"][l + 1].view(M, D), bias=w.b2, residual=st["h"][l], drop_p=p,
                 seed=seed, salt=(l * 3 + s) * 4 + 2, seed_dev=sd)

    def _layer_fwd_cls(self, l, ctx):
        """The last fused layer of the vslt stream when only its CLS row is consumed (see `cls_only`)."""
        st = self.ws[0]
        w, _, h16 = self.blocks[(l, 0)]
        B, T, M = ctx["B"], st["T"], st["M"]
        p, seed, sd = ctx["p"], ctx["seed"], ctx["seed_dev"]
        x = st["X"][l]
        ops.layernorm_fwd(x, w.ln1_g, w.ln1_b, st["xn"][l])                       # K and V need every row
        ops.gemm(st["xn"][l], h16.wqkv, out=st["qkv"][l], bias=w.bqkv)
        ops.attn_fwd(st["qkv"][l], ctx["kv_len"][0], B, T, st["O"][l], st["lse"][l], q_rows=5)   # first query tile only
        st["c_x"].copy_(x[:, 4, :])
        st["c_o"].copy_(st["O"][l].view(B, T, D)[:, 4, :])
        ops.layernorm_fwd(st["c_x"], w.ln2_g, w.ln2_b, st["c_hn"], add=st["c_o"], sum_out=st["c_h"])
        ops.gemm(st["c_hn"], h16.w1, out=st["c_a"], bias=w.b1, relu=True, drop_p=p, seed=seed, salt=(l * 3) * 4 + 1,
                 seed_dev=sd, mask_out=st["c_am"])
        ops.gemm(st["c_a"], h16.w2, out=st["c_y"], bias=w.b2, residual=st["c_h"], drop_p=p, seed=seed,
                 salt=(l * 3) * 4 + 2, seed_dev=sd)
        st["X"][l + 1][:, 4, :] = st["c_y"]

    def _layer_bwd_cls(self, l, d_cls, ctx):
        st = self.ws[0]
        w, g, h16 = self.blocks[(l, 0)]
        B, T, M = ctx["B"], st["T"], st["M"]
        p, seed = ctx["p"], ctx["seed"]
        st["c_gy"].copy_(d_cls * ctx["gscale"])
        if p > 0:
            ops.dropout_apply(st["c_gy"], st["c_gyd"], p, seed, (l * 3) * 4 + 2, seed_dev=ctx["seed_dev"])
            gyd = st["c_gyd"]
        else:
            gyd = st["c_gy"]
        scale = 1.0 / (1.0 - p) if p > 0 else 1.0
        ops.gemm(gyd, self.wT[(l, 0, "w2")], out=st["c_ga"], gate=st["c_am"], alpha=scale)
        ops.gemm_wgrad(gyd, st["c_a"], g.w2, dbias=g.b2)
        ops.gemm(st["c_ga"], self.wT[(l, 0, "w1")], out=st["c_ghn"])
        ops.gemm_wgrad(st["c_ga"], st["c_hn"], g.w1, dbias=g.b1)
        ops.layernorm_bwd(st["c_ghn"], st["c_h"], st["c_gy"], w.ln2_g, st["c_gh"], g.ln2_g, g.ln2_b)
        # the attention's dO / delta / dQ are zero everywhere but in the CLS rows
        st["g_h"].zero_()
        st["g_h"].view(B, T, D)[:, 4, :] = st["c_gh"]
        if self.cls_attn_bwd_fast:
            # one query row per sample: the single-query kernel (delta, zero rows and all of dQ|dK|dV in one pass)
            ops.attn_bwd_single_query(st["qkv"][l], st["c_gh"], st["c_o"], ctx["kv_len"][0], B, T, 4, st["lse"][l],
                                      st["g_qkv"])
        else:
            st["delta"].zero_()
            st["delta"][:, :, 4] = (st["c_gh"].float().view(B, 4, 64) * st["c_o"].float().view(B, 4, 64)).sum(-1)
            st["g_qkv"][:, :D].zero_()
            ops.attn_bwd(st["qkv"][l], st["O"][l], st["g_h"], ctx["kv_len"][0], B, T, st["lse"][l], st["delta"], None,
                         st["g_qkv"], q_rows=5)
        ops.gemm(st["g_qkv"], self.wT[(l, 0, "qkv")], out=st["g_xn"])
        ops.gemm_wgrad(st["g_qkv"], st["xn"][l], g.wqkv, dbias=g.bqkv)
        # LN1 (+ residual); the same pass writes dropout(dX[l]) for the layer below (its FFN2 output dropout mask): g_yd is
        # free by now, its readers (FFN2 dgrad / wgrad of this layer) ran earlier on this stream
        if p > 0 and l > 0 and self.fuse_grad_dropout:
            ops.layernorm_bwd(st["g_xn"], st["X"][l].view(M, D), st["g_h"], w.ln1_g, st["g_x"].view(M, D), g.ln1_g, g.ln1_b,
                              dx_drop=st["g_yd"].view(M, D), drop_p=p, seed=seed, salt=((l - 1) * 3) * 4 + 2,
                              seed_dev=ctx["seed_dev"])
            ctx["gyd_layer"][0] = l - 1
        else:
            ops.layernorm_bwd(st["g_xn"], st["X"][l].view(M, D), st["g_h"], w.ln1_g, st["g_x"].view(M, D), g.ln1_g, g.ln1_b)
        st["g_y"], st["g_x"] = st["g_x"], st["g_y"]

    # ------------------------------------------------------------------------------------------------------------
    def backward(self, d_cls, token=None):
        m = self.model
        ctx = self.ctx
        if token is not None and token != ctx["token"]:
            raise RuntimeError("FusedPath.backward: another forward ran between this backward's forward and now (e.g. an "
                               "eval probe inside a train step); the saved activations live in per-shape workspaces and "
                               "were overwritten. Run backward before the next forward.")
        if ctx["ws"] is not self._ws_cur:
            self._ensure_workspace(ctx["B"], ctx["L"], ctx["n_img"])
        if self.head_hook is not None:
            self.head_hook()
        # Gradient accumulation (a second backward without zero_grad in between): the kernels WRITE the flat gradient
        # buffer (and unscale it range by range), so the gradients still attached to the parameters are set aside and
        # added back at the end. After zero_grad(set_to_none=True) -- the torch default -- this costs nothing.
        n0, p0 = self.layout[0]
        acc_prev = None
        if p0.grad is not None and p0.grad.data_ptr() == self.gviews[n0].data_ptr():
            acc_prev = self.flat_g[: self.live_end()].clone()
        NL = m.num_layers
        B, p, seed = ctx["B"], ctx["p"], ctx["seed"]
        self.flat_g.zero_()
        gscale = 1.0 if ctx["f32"] else self.grad_scale      # fp32 gradients need no scale
        ctx["gscale"] = gscale
        ctx["gyd_layer"] = {}      # stream -> layer whose dropped input gradient (g_yd) is already in place
        cls_last = m.vsltonly == 1 and self.cls_only and not ctx["f32"]
        for s in range(3):
            if s == 0 and cls_last:
                continue                 # the CLS-only last layer takes dL/dCLS directly and overwrites g_y
            self.ws[s]["g_y"].zero_()
        if not cls_last:
            self.ws[0]["g_y"][:, 4, :] = (d_cls * gscale).to(self.ws[0]["g_y"].dtype)
        for l in range(NL - 1, -1, -1):
            last = m.vsltonly == 1 and l == NL - 1
            if not last:
                self._fork()
                for s in (1, 2):
                    with self._lane(s):
                        self._layer_bwd(l, s, ctx)
            if last and cls_last:
                self._layer_bwd_cls(l, d_cls, ctx)
            else:
                self._layer_bwd(l, 0, ctx)
            if not last:
                self._join()
            if self.debug_trace is not None:       # tools/gpu_grad_trace.py: dL/dX[l] per stream (scaled fp16)
                for s in ([0] if last else [0, 1, 2]):
                    self.debug_trace[("dX", l, s)] = self.ws[s]["g_y"].float().div(gscale).cpu()
            # after _layer_bwd, g_y of each processed stream holds dX[l] (gradient wrt the layer input)
            if l > 0:
                upper_has_it = 0 if last else 1
                # rows 0..3 change here: streams whose dropped gradient copy was written by this layer's LayerNorm backward
                # get those rows refreshed in the same launch
                fused = [ctx["gyd_layer"].get(s) == l - 1 for s in range(3)]
                ops.bottleneck_mix_bwd(self.ws[0]["g_y"], self.ws[1]["g_y"], self.ws[2]["g_y"], upper_has_it,
                                       ctx["missing"],
                                       dropped=tuple(self.ws[s]["g_yd"] if fused[s] else None for s in range(3)),
                                       drop_p=p if any(fused) else 0.0, seed=seed,
                                       salts=tuple(((l - 1) * 3 + s) * 4 + 2 for s in range(3)), seed_dev=ctx["seed_dev"])
            self._range_done(*self.grad_range_of_layer(l))
        # prologue + projections
        F = "fusion_transformer"
        self._fork()
        for s in (1, 2, 0):
            with self._lane(s):
                a = self._prologue_args(s, ctx)
                ops.stream_prologue_bwd(dX0=self.ws[s]["g_y"],
                                        g_val=self.G("ie_vslt.0.weight", 4, D) if s == 0 else None,
                                        g_tim=self.G("ie_time.0.weight", 4, D), g_feat=self.G("ie_feat.weight", 20, D),
                                        g_cls=self.G(f"{F}.cls_token_per_modality.{s}", D),
                                        g_bott=self.G(f"{F}.bottlenecks", 4, D),
                                        g_ln=self.G(f"{F}.layer_norms_in.{s}.weight", 2, D),
                                        dproj=self.g_proj[s], **a)
                if s == 1:
                    ops.gemm_wgrad(self.g_proj[1], ctx["img16"], self.G("linear.weight", D, 768),
                                   dbias=self.G("linear.bias", D))
                elif s == 2:
                    ops.gemm_wgrad(self.g_proj[2], ctx["txts16"], self.G("txt_embedding.weight", D, 768),
                                   dbias=self.G("txt_embedding.bias", D))
        self._join()
        self._range_done(*self.grad_range_of_layer(-1))
        if acc_prev is not None:
            self.flat_g[: self.live_end()].add_(acc_prev)
        self._publish_grads()

    def _range_done(self, a, b):
        """flat_g[a:b] is complete: remove the fp16 gradient scale and hand the range to the data-parallel hook
        (trainer.GradSync launches its all-reduce on the communication stream while the backward continues)."""
        f = self.grad_post_scale / self.ctx["gscale"]
        if f != 1.0:
            self.flat_g[a:b].mul_(f)        # one pass: fp16 gradient scale out, data-parallel averaging in
        if self.comm_hook is not None:
            self.comm_hook(a, b)

    def _layer_bwd(self, l, s, ctx):
        st = self.ws[s]
        w, g, h16 = self.blocks[(l, s)]
        B, T, M = ctx["B"], st["T"], st["M"]
        p, seed = ctx["p"], ctx["seed"]
        if ctx["f32"]:       # fp32 mode: transposed fp32 masters (dgrad operands), split into bf16x3 by ops.gemm
            wT = {k: getattr(w, k).t().contiguous() for k in ("w2", "w1")}
            wT["qkv"] = w.wqkv.t().contiguous()
        else:
            wT = {k: self.wT[(l, s, k)] for k in ("w2", "w1", "qkv")}
        gy = st["g_y"].view(M, D)
        if p > 0:
            # dropout(gy) with the mask of this layer's FFN2 output dropout: already written by the LayerNorm backward of
            # the layer above (dx_drop) + the bottleneck exchange (rows 0..3) when that layer ran through _layer_bwd
            if ctx["gyd_layer"].get(s) != l:
                ops.dropout_apply(gy, st["g_yd"].view(M, D), p, seed, (l * 3 + s) * 4 + 2, seed_dev=ctx["seed_dev"])
            gyd = st["g_yd"].view(M, D)
        else:
            gyd = gy
        scale = 1.0 / (1.0 - p) if p > 0 else 1.0
        # FFN2: y = h + drop2(a W2^T + b2)
        ops.gemm(gyd, wT["w2"], out=st["g_a"], gate=st["a"][l] if ctx["f32"] else st["am"][l], alpha=scale)
        ops.gemm_wgrad(gyd, st["a"][l], g.w2, dbias=g.b2)     # bias gradient = column sums, fused into the wgrad kernel
        # FFN1: a = drop1(relu(hn W1^T + b1))
        ops.gemm(st["g_a"], wT["w1"], out=st["g_hn"])
        ops.gemm_wgrad(st["g_a"], st["hn"][l], g.w1, dbias=g.b1)
        # LN2 (+ residual): h = x + O
        if ctx["f32"]:
            ops.layernorm_bwd(st["g_hn"], st["h"][l], gy, w.ln2_g, st["g_h"], g.ln2_g, g.ln2_b)
            ops.attn_bwd(st["qkv"][l], st["O"][l], st["g_h"], ctx["kv_len"][s], B, T, st["lse"][l], st["delta"], None,
                         st["g_qkv"])
        else:
            # g_h is the attention's dO: the same pass writes delta and zeroes the dQ columns (fused attn_bwd protocol)
            ops.layernorm_bwd_attn(st["g_hn"], st["h"][l], gy, w.ln2_g, st["g_h"], g.ln2_g, g.ln2_b, st["O"][l], T,
                                   st["delta"], st["g_qkv"])
            ops.attn_bwd(st["qkv"][l], st["O"][l], st["g_h"], ctx["kv_len"][s], B, T, st["lse"][l], st["delta"], None,
                         st["g_qkv"])
        ops.gemm(st["g_qkv"], wT["qkv"], out=st["g_xn"])
        ops.gemm_wgrad(st["g_qkv"], st["xn"][l], g.wqkv, dbias=g.bqkv)
        # LN1 (+ residual)
        if p > 0 and l > 0 and self.fuse_grad_dropout:     # + dropout(dX[l]) for the layer below, see _layer_bwd_cls
            ops.layernorm_bwd(st["g_xn"], st["X"][l].view(M, D), st["g_h"], w.ln1_g, st["g_x"].view(M, D), g.ln1_g, g.ln1_b,
                              dx_drop=st["g_yd"].view(M, D), drop_p=p, seed=seed, salt=((l - 1) * 3 + s) * 4 + 2,
                              seed_dev=ctx["seed_dev"])
            ctx["gyd_layer"][s] = l - 1
        else:
            ops.layernorm_bwd(st["g_xn"], st["X"][l].view(M, D), st["g_h"], w.ln1_g, st["g_x"].view(M, D), g.ln1_g, g.ln1_b)
        st["g_y"], st["g_x"] = st["g_x"], st["g_y"]

    def live_end(self):
        """flat_w[:live_end()] / flat_g[:live_end()] are the parameters that receive gradients. With --mbt-only-vslt 1
        the img/txt blocks of the last layer are never run (mbt_encoder.py:757-763): they are the tail of the flat
        layout and keep `grad = None` like in the reference (torch.optim.AdamW skips them)."""
        m = self.model
        if m.vsltonly == 1:
            return self.offs[f"fusion_transformer.layer_stacks.{m.num_layers - 1}.1.attention_prenorm.gamma"]
        return self.total

    def _publish_grads(self):
        """param.grad <- view of flat_g (accumulates if the caller kept older gradients)."""
        self.grads_fresh = True
        end = self.live_end()
        for n, prm in self.layout:
            if self.offs[n] >= end:
                continue
            gv = self.gviews[n]
            if prm.grad is None:
                prm.grad = gv
            elif prm.grad.data_ptr() != gv.data_ptr():
                prm.grad.add_(gv)

    # ranges of flat_g per backward stage (used by the DDP bucket plan)
    def grad_range_of_layer(self, l):
        F = "fusion_transformer"
        if l >= 0:
            a = self.offs[f"{F}.layer_stacks.{l}.0.attention_prenorm.gamma"]
            nxt = f"{F}.layer_stacks.{l + 1}.0.attention_prenorm.gamma"
            b = self.offs[nxt] if nxt in self.offs else self.live_end()
            return a, b
        return 0, self.offs[f"{F}.layer_stacks.0.0.attention_prenorm.gamma"]


def _contig_strides(shape):
    st, acc = [], 1
    for d in reversed(shape):
        st.append(acc)
        acc *= d
    return tuple(reversed(st))


class _FusedFn(torch.autograd.Function):
    """Autograd boundary: inputs are data tensors (no gradient) + a dummy trigger; the parameters' gradients are
    produced by the kernels directly into the flat gradient buffer (see FusedPath._publish_grads)."""

    @staticmethod
    def forward(ctx, trigger, fp, x, input_lengths, txts, txt_lengths, img_feats, img_time, txt_time, missing):
        ctx.fp = fp
        out = fp.forward(x, input_lengths, txts, txt_lengths, img_feats, img_time, txt_time, missing,
                         training=fp.model.training)
        ctx.token = fp.ctx["token"]
        return out

    @staticmethod
    def backward(ctx, d_cls):
        ctx.fp.backward(d_cls.contiguous(), token=ctx.token)
        return (torch.zeros(1, device=d_cls.device),) + (None,) * 9
