"""SwinFeed -- B200-native forward of the frozen Swin-T image encoder that feeds the fused path (SURVEY.md 8f rank 1;
reference tri_mbt_vsltcls.py:205-211 calls builder/models/src/swin_transformer.py under torch.no_grad).

The stock torchvision module stays the parameter / state_dict master (`model.img_encoder`); this class holds fp16,
GEMM-padded copies of its weights and runs the forward as 7 launches per block: LayerNorm+shift+window-partition,
QKV GEMM, window attention, proj GEMM, window-reverse+residual+LayerNorm, FC1 GEMM (+GELU), FC2 GEMM (+residual) --
matmuls on the tcgen05 GEMM, the rest in csrc/swin.cu. The stock eager forward spends ~90 % of its GPU time in
LayerNorm / copy / roll / bmm / softmax / GELU glue (profiles/r1b_launches.csv); here every activation is read and
written once per fused stage. Channel counts that are not multiples of 128 are zero-padded (96 -> 128, 192 -> 256).
"""
from __future__ import annotations

import torch

from . import ops

# (C, heads, H=W, depth, padded C)
STAGES = [(96, 3, 56, 2, 128), (192, 6, 28, 2, 256), (384, 12, 14, 6, 384), (768, 24, 7, 2, 768)]


def _pad2(w, rows, cols):
    out = torch.zeros(rows, cols, dtype=torch.float16, device=w.device)
    out[: w.shape[0], : w.shape[1]] = w.to(torch.float16)
    return out


def _pad1(b, n):
    out = torch.zeros(n, dtype=torch.float32, device=b.device)
    out[: b.numel()] = b.float()
    return out


def _up(n, m=128):
    return (n + m - 1) // m * m


class SwinFeed:
    def __init__(self, swin: torch.nn.Module):
        """swin: torchvision-style SwinTransformer (1-channel patch conv) already on the CUDA device."""
        f = swin.features
        dev = f[0][0].weight.device
        if dev.type != "cuda":
            raise RuntimeError("SwinFeed needs CUDA weights (no CPU fallback)")
        conv, ln0 = f[0][0], f[0][2]
        if tuple(conv.weight.shape) != (96, 1, 4, 4):
            raise NotImplementedError("SwinFeed implements the reference's 1-channel Swin-T patch embedding")
        self.embed = dict(Wt=conv.weight.detach().float().reshape(96, 16).t().contiguous(), b=conv.bias.detach().float(),
                          g=ln0.weight.detach().float().contiguous(), be=ln0.bias.detach().float().contiguous())
        self.blocks, self.merges = [], []
        for si, (C, heads, H, depth, Cp) in enumerate(STAGES):
            stage = f[1 + 2 * si]
            assert len(stage) == depth
            blks = []
            for j, blk in enumerate(stage):
                a = blk.attn
                idx = a.relative_position_index.view(-1)
                rel = a.relative_position_bias_table.detach().float()[idx].view(49, 49, heads).permute(2, 0, 1).contiguous()
                nq = _up(3 * C)
                blks.append(dict(
                    shift=0 if (j % 2 == 0 or H <= 7) else 3,
                    n1g=blk.norm1.weight.detach().float().contiguous(), n1b=blk.norm1.bias.detach().float().contiguous(),
                    n2g=blk.norm2.weight.detach().float().contiguous(), n2b=blk.norm2.bias.detach().float().contiguous(),
                    wqkv=_pad2(a.qkv.weight.detach(), nq, Cp), bqkv=_pad1(a.qkv.bias.detach(), nq),
                    wproj=_pad2(a.proj.weight.detach(), Cp, Cp), bproj=_pad1(a.proj.bias.detach(), Cp),
                    wfc1=_pad2(blk.mlp[0].weight.detach(), 4 * C, Cp), bfc1=blk.mlp[0].bias.detach().float().contiguous(),
                    wfc2=_pad2(blk.mlp[3].weight.detach(), Cp, 4 * C), bfc2=_pad1(blk.mlp[3].bias.detach(), Cp),
                    rel=rel))
            self.blocks.append(blks)
            if si < 3:
                mg = f[2 + 2 * si]
                Cn = STAGES[si + 1][4]
                self.merges.append(dict(g=mg.norm.weight.detach().float().contiguous(),
                                        b=mg.norm.bias.detach().float().contiguous(),
                                        w=_pad2(mg.reduction.weight.detach(), Cn, 4 * C)))
        self.norm = dict(g=swin.norm.weight.detach().float().contiguous(), b=swin.norm.bias.detach().float().contiguous())
        self.device = dev
        self._ws_cache = {}      # (n_total, n_chunk) -> (ws, out); see _workspace

    def _workspace(self, n_total, n_chunk):
        """Stage workspaces: stages 1-2 (the high-resolution stages, >= 10^5 token rows even for a third of the batch) are
        sized for one CHUNK of images, stages 3-4 and the feature output for the whole batch. Cached per batch size and
        NEVER freed while anything can still reference them: a captured CUDA graph (trainer.GraphedStep) bakes their raw
        device pointers and additionally pins the objects (`pin()`); the small LRU only bounds what eager calls with
        ever-changing shapes can accumulate."""
        key = (n_total, n_chunk)
        hit = self._ws_cache.pop(key, None)
        if hit is not None:
            self._ws_cache[key] = hit          # most recently used last
            self.ws, self.out = hit
            return self.ws
        dev = self.device
        ws = []
        for si, (C, heads, H, depth, Cp) in enumerate(STAGES):
            M = (n_chunk if si < 2 else n_total) * H * H
            h16 = lambda *s: torch.empty(*s, dtype=torch.float16, device=dev)
            ws.append(dict(M=M, x=h16(M, Cp), xw=h16(M, Cp), qkv=h16(M, _up(3 * C)),
                           ao=torch.zeros(M, Cp, dtype=torch.float16, device=dev),      # pad columns stay zero
                           y=h16(M, Cp), hn=h16(M, Cp), a=h16(M, 4 * C),
                           mg=h16(M // 4, 4 * C) if H > 7 else None))
        self.out = torch.empty(n_total * 49, 768, dtype=torch.float16, device=dev)
        self.ws = ws
        self._ws_cache[key] = (ws, self.out)
        while len(self._ws_cache) > 3:
            self._ws_cache.pop(next(iter(self._ws_cache)))
        return ws

    def pin(self):
        """The workspace objects of the most recent call (held by a captured graph so they outlive the LRU)."""
        return (self.ws, self.out)

    @staticmethod
    def n_chunks(n_img):
        """Large batches run the two high-resolution stages (patch embedding, stages 1-2: 56x56 and 28x28 tokens per image,
        60 % of the encoder's time) in 3 chunks of images: the host->device upload of chunk c+1 then overlaps the encoder
        of chunk c (trainer.GraphedStep stages the upload per chunk), and their workspaces shrink to a third. Stages 3-4
        (14x14 / 7x7 tokens) run on the whole batch: their kernels are too small to split (measured: +0.9 ms per step
        when all four stages were chunked)."""
        import os
        k = int(os.environ.get("TMP_B200_SWIN_CHUNKS", "3"))     # 1 = whole batch at once (A/B measurements)
        return k if (k > 1 and n_img >= 32 * k and n_img % k == 0) else 1

    @torch.no_grad()
    def __call__(self, img, ready=None, live=None):
        """img fp32 [N,1,224,224] (or [N,224,224]) in [0,1] -> features fp16 [N,49,768] = norm(features(img)).
        ready: optional list of CUDA events, one per chunk (`n_chunks(N)` of them): chunk c is not touched before
        ready[c] has fired (its pixels are still being uploaded).
        live: optional uint8 [N] device tensor; images with 0 are skipped by every kernel (their features have no
        consumer) and come out as zero rows."""
        n_img = img.numel() // (224 * 224)
        if img.dtype != torch.float32 or not img.is_contiguous():
            if ready is not None:                       # a cast reads every pixel: all chunks must have landed
                for ev in ready:
                    torch.cuda.current_stream().wait_event(ev)
                ready = None
            img = img.float().contiguous()
        img = img.reshape(n_img, 224, 224)
        k = self.n_chunks(n_img)
        nc = n_img // k
        ws = self._workspace(n_img, nc)
        rows3 = nc * 14 * 14                            # stage-3 input rows produced by one chunk
        for c in range(k):
            if ready is not None:
                torch.cuda.current_stream().wait_event(ready[c] if len(ready) == k else ready[-1])
            lv = None if live is None else live[c * nc:(c + 1) * nc]
            self._stages(img[c * nc:(c + 1) * nc], nc, ws, 0, 2, ws[2]["x"][c * rows3:(c + 1) * rows3], lv)
        self._stages(None, n_img, ws, 2, 4, None, live)
        ops.swin_ln_window(ws[3]["x"], self.norm["g"], self.norm["b"], n_img, 7, 768, 768, 0, self.out, live=live,
                           zero_dead=True)
        return self.out.view(n_img, 49, 768)

    def _stages(self, img, n_img, ws, s0, s1, x_next, live=None):
        """Stages [s0, s1) on n_img images. Stage 0 starts from the pixels; the merged output of stage s1-1 goes to `x_next`
        (a slice of the next stage's input) when s1 < 4."""
        if s0 == 0:
            e = self.embed
            ops.swin_patch_embed_ln(img, e["Wt"], e["b"], e["g"], e["be"], ws[0]["x"], STAGES[0][4], live=live)
        for si in range(s0, s1):
            C, heads, H, depth, Cp = STAGES[si]
            w = ws[si]
            rl = dict(row_live=live, rows_per_group=H * H) if live is not None else {}      # token rows are image-major
            for b in self.blocks[si]:
                sh = b["shift"]
                ops.swin_ln_window(w["x"], b["n1g"], b["n1b"], n_img, H, C, Cp, sh, w["xw"], live=live)
                ops.gemm(w["xw"], b["wqkv"], out=w["qkv"], bias=b["bqkv"], **rl)
                ops.swin_window_attn(w["qkv"], b["rel"], n_img, H, C, heads, sh, w["ao"], live=live)
                ops.gemm(w["ao"], b["wproj"], out=w["y"], bias=b["bproj"], **rl)
                ops.swin_unwindow_add_ln(w["y"], w["x"], b["n2g"], b["n2b"], n_img, H, C, Cp, sh, w["hn"], live=live)
                ops.gemm(w["hn"], b["wfc1"], out=w["a"], bias=b["bfc1"], relu=2, **rl)
                ops.gemm(w["a"], b["wfc2"], out=w["x"], bias=b["bfc2"], residual=w["x"], **rl)
            if si < 3:
                m = self.merges[si]
                ops.swin_merge_ln(w["x"], m["g"], m["b"], n_img, H, C, Cp, w["mg"], live=live)
                rm = dict(row_live=live, rows_per_group=(H // 2) * (H // 2)) if live is not None else {}
                ops.gemm(w["mg"], m["w"], out=(x_next if si == s1 - 1 and x_next is not None else ws[si + 1]["x"]), **rm)
