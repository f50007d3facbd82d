"""Stand-alone launches of every hot kernel at the bench workload's shapes (B=64, T_v=1005, ...), for
`ncu --set full` captures (two launches per kernel) and for CUDA-event timing tables (--time).

    ncu --set full --clock-control none --import-source on -k regex:'attn|gemm|layernorm|umse|prologue|colsum' \
        -o gpurun_out/prof python tools/profile_kernels.py
    python tools/profile_kernels.py --time        # prints a JSON table: kernel, shape, ms, TFLOP/s or GB/s
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from medical_tri_modal_pilot_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--sweep", action="store_true",
                    help="BASELINE config 5: attention alone, seq 256..4096, full / ragged / de-selected samples")
    ap.add_argument("--flush", action="store_true",
                    help="time every launch separately with an L2 flush (256 MB write) in between: cold-L2 figures for the "
                         "HBM-bound kernels, as they run inside the step")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kernel_times.json"))
    a = ap.parse_args()
    dev = "cuda"
    torch.manual_seed(0)
    B = a.B
    only = [s for s in a.only.split(",") if s]
    cases = []

    def add(name, fn, flops=None, bytes_=None, shape=""):
        if only and not any(s in name for s in only):
            return
        cases.append((name, fn, flops, bytes_, shape))

    if a.sweep:
        # kv_len patterns: "full" = every sample at T; "ragged" = U{T/4..T} (one sample at T); "half_missing" = the
        # modality is de-selected (kv_len 0) for every second sample, the rest ragged. Flops count live
        # (query, key) pairs only: 4*len_b^2*64 per head forward, x2.5 backward.
        gen = torch.Generator().manual_seed(0)
        for T in (256, 512, 1024, 2048, 4096):
            Bs = max(8, min(B, (64 * 1024) // T))
            M = Bs * T
            Tl = ops.lse_len(T)
            qkv = torch.randn(M, 768, device=dev).half()
            O = torch.empty(M, 256, device=dev, dtype=torch.float16)
            lse = torch.zeros(Bs, 4, Tl, device=dev)
            dO = (torch.randn(M, 256, device=dev) * 0.1).half()
            delta = torch.empty(Bs, 4, Tl, device=dev)
            dq = torch.empty(M, 256, device=dev)
            dqkv = torch.empty(M, 768, device=dev, dtype=torch.float16)
            rag = torch.randint(T // 4, T + 1, (Bs,), generator=gen); rag[0] = T
            half = rag.clone(); half[1::2] = 0
            for pat, lens in (("full", torch.full((Bs,), T)), ("ragged", rag), ("half_missing", half)):
                kv = lens.to(device=dev, dtype=torch.int32)
                f_fwd = float((lens.double() ** 2).sum()) * 4.0 * 64 * 4
                add(f"sweep_attn_fwd_S{T}_{pat}", lambda qkv=qkv, kv=kv, T=T, O=O, lse=lse, Bs=Bs:
                    ops.attn_fwd(qkv, kv, Bs, T, O, lse), f_fwd, None, f"B={Bs} H=4 S={T} d=64 kv_len={pat}")
                delta.zero_()
                add(f"sweep_attn_bwd_S{T}_{pat}", lambda qkv=qkv, kv=kv, T=T, O=O, lse=lse, dO=dO, delta=delta, dq=dq,
                    dqkv=dqkv, Bs=Bs: ops.attn_bwd(qkv, O, dO, kv, Bs, T, lse, delta, None, dqkv), 2.5 * f_fwd, None,
                    f"B={Bs} H=4 S={T} d=64 kv_len={pat} (fused protocol: delta / dQ zeroing by the LayerNorm backward)")
    for T in (() if a.sweep else (1005, 2005, 152, 133)):
        M = B * T
        Tl = ops.lse_len(T)
        qkv = (torch.randn(M, 768, device=dev)).half()
        kv = torch.full((B,), T, device=dev, dtype=torch.int32)
        O = torch.empty(M, 256, device=dev, dtype=torch.float16)
        lse = torch.zeros(B, 4, Tl, device=dev)
        dO = (torch.randn(M, 256, device=dev) * 0.1).half()
        delta = torch.empty(B, 4, Tl, device=dev)
        dq = torch.empty(M, 256, device=dev)
        dqkv = torch.empty(M, 768, device=dev, dtype=torch.float16)
        f_fwd = 4.0 * T * T * 64 * 4 * B
        add(f"attn_fwd_T{T}", lambda qkv=qkv, kv=kv, T=T, O=O, lse=lse: ops.attn_fwd(qkv, kv, B, T, O, lse), f_fwd, None,
            f"B={B} H=4 T={T} d=64")
        delta.zero_()
        add(f"attn_bwd_T{T}", lambda qkv=qkv, kv=kv, T=T, O=O, lse=lse, dO=dO, delta=delta, dq=dq, dqkv=dqkv:
            ops.attn_bwd(qkv, O, dO, kv, B, T, lse, delta, None, dqkv), 2.5 * f_fwd, None,
            f"B={B} H=4 T={T} d=64 (fused protocol, one launch)")
        add(f"attn_bwd_standalone_T{T}", lambda qkv=qkv, kv=kv, T=T, O=O, lse=lse, dO=dO, delta=delta, dq=dq, dqkv=dqkv:
            ops.attn_bwd(qkv, O, dO, kv, B, T, lse, delta, dq, dqkv), 2.5 * f_fwd, None,
            f"B={B} H=4 T={T} d=64 (stand-alone: delta + memset + main + dQ convert)")
        if T == 2005:
            continue
        # GEMMs of one encoder block at this stream length
        for (nm, N, K, relu) in (("qkv", 768, 256, False), ("ffn1", 1024, 256, True), ("ffn2", 256, 1024, False)):
            A = torch.randn(M, K, device=dev).half()
            W = (torch.randn(N, K, device=dev) / K ** 0.5).half()
            bias = torch.randn(N, device=dev)
            out = torch.empty(M, N, device=dev, dtype=torch.float16)
            res = torch.randn(M, N, device=dev).half() if nm == "ffn2" else None
            add(f"gemm_{nm}_T{T}", lambda A=A, W=W, out=out, bias=bias, relu=relu, res=res:
                ops.gemm(A, W, out=out, bias=bias, relu=relu, residual=res, drop_p=0.1, seed=1, salt=2),
                2.0 * M * N * K, None, f"M={M} N={N} K={K}")
            dY = torch.randn(M, N, device=dev).half()
            dW = torch.zeros(N, K, device=dev)
            add(f"wgrad_{nm}_T{T}", lambda dY=dY, A=A, dW=dW: ops.gemm_wgrad(dY, A, dW), 2.0 * M * N * K, None,
                f"M={M} N={N} K={K}")
            cs = torch.zeros(N, device=dev)
            add(f"wgradb_{nm}_T{T}", lambda dY=dY, A=A, dW=dW, cs=cs: ops.gemm_wgrad(dY, A, dW, dbias=cs), 2.0 * M * N * K,
                None, f"M={M} N={N} K={K} (+ fused bias gradient)")
            add(f"colsum_{nm}_T{T}", lambda dY=dY, cs=cs: ops.colsum(dY, cs), None, M * N * 2.0, f"M={M} N={N}")
        # dgrads of the same block: FFN2 dgrad is gated by the 1-bit ReLU/dropout mask FFN1 wrote in the forward pass
        gy = torch.randn(M, 256, device=dev).half()
        w2t = (torch.randn(1024, 256, device=dev) / 16).half()
        am = torch.randint(-2 ** 31, 2 ** 31 - 1, (M, 32), device=dev, dtype=torch.int32)
        ga = torch.empty(M, 1024, device=dev, dtype=torch.float16)
        add(f"dgrad_ffn2_gated_T{T}", lambda gy=gy, w2t=w2t, ga=ga, am=am: ops.gemm(gy, w2t, out=ga, gate=am, alpha=1.0 / 0.9),
            2.0 * M * 1024 * 256, None, f"M={M} N=1024 K=256, 1-bit gate")
        w1t = (torch.randn(256, 1024, device=dev) / 32).half()
        ghn = torch.empty(M, 256, device=dev, dtype=torch.float16)
        add(f"dgrad_ffn1_T{T}", lambda ga=ga, w1t=w1t, ghn=ghn: ops.gemm(ga, w1t, out=ghn), 2.0 * M * 1024 * 256, None,
            f"M={M} N=256 K=1024")
        gqkv = torch.randn(M, 768, device=dev).half()
        wqt = (torch.randn(256, 768, device=dev) / 28).half()
        add(f"dgrad_qkv_T{T}", lambda gqkv=gqkv, wqt=wqt, ghn=ghn: ops.gemm(gqkv, wqt, out=ghn), 2.0 * M * 768 * 256, None,
            f"M={M} N=256 K=768")
        x = torch.randn(M, 256, device=dev).half()
        y = torch.empty_like(x)
        g = torch.ones(256, device=dev); b = torch.zeros(256, device=dev)
        add(f"layernorm_fwd_T{T}", lambda x=x, y=y: ops.layernorm_fwd(x, g, b, y), None, M * 256 * 2 * 2.0, f"rows={M}")
        h = torch.empty_like(x)
        add(f"layernorm_fwd_add_T{T}", lambda x=x, y=y, h=h: ops.layernorm_fwd(x, g, b, y, add=x, sum_out=h), None,
            M * 256 * 2 * 4.0, f"rows={M}")
        dx = torch.empty_like(x); dg = torch.zeros(256, device=dev); db = torch.zeros(256, device=dev)
        add(f"layernorm_bwd_T{T}", lambda x=x, dx=dx: ops.layernorm_bwd(x, x, x, g, dx, dg, db), None,
            M * 256 * 2 * 4.0, f"rows={M}")
        # the LayerNorm backward in front of the attention backward: + O read, + dQ columns zeroed, + delta written
        add(f"layernorm_bwd_attn_T{T}", lambda x=x, dx=dx, O=O, delta=delta, dqkv=dqkv, T=T:
            ops.layernorm_bwd_attn(x, x, x, g, dx, dg, db, O, T, delta, dqkv), None, M * 256 * 2 * 6.0 + B * 4 * T * 4.0,
            f"rows={M} (x, dy, dres, O read; dx + zeroed dQ written)")
        # encoder prologue of this stream length as the step runs it: UMSE (T=1005) or projected rows (152 / 133) -> X0
        n = T - 5
        X0 = torch.empty(B, T, 256, device=dev, dtype=torch.float16)
        br = lambda: [torch.randn(256, device=dev), torch.randn(256, device=dev), torch.ones(256, device=dev),
                      torch.zeros(256, device=dev)]
        val4, tim4 = br(), br()
        Wfe = torch.randn(20, 256, device=dev); cls = torch.randn(256, device=dev); bott = torch.randn(4, 256, device=dev)
        common = dict(tim4=tim4, Wfeat=Wfe, cls=cls, bottlenecks=bott, ln_g=g, ln_b=b, pe=None, drop_p=0.1, seed=3, salt=7)
        if T == 1005:
            xv = torch.rand(B, n, 3, device=dev); xv[:, :, 2] = torch.randint(0, 18, (B, n), device=dev).float()
            pa = dict(kind=0, B=B, n=n, x=xv, val4=val4, proj=None, times=None, n_slots=0, feat_id=0, **common)
            byt = B * n * 12.0 + B * T * 512.0
        else:
            pr = torch.randn(B * n, 256, device=dev).half(); tm = -torch.rand(B, device=dev)
            pa = dict(kind=1, B=B, n=n, x=None, val4=None, proj=pr, times=tm, n_slots=1, feat_id=19, **common)
            byt = B * n * 512.0 + B * T * 512.0
        add(f"stream_prologue_fwd_T{T}", lambda pa=pa, X0=X0: ops.stream_prologue_fwd(X0=X0, **pa), None, byt, f"B={B} n={n}")
        gX = torch.randn(B, T, 256, device=dev).half()
        gacc = lambda *s_: torch.zeros(*s_, device=dev)
        gb = dict(g_val=gacc(4, 256) if T == 1005 else None, g_tim=gacc(4, 256), g_feat=gacc(20, 256), g_cls=gacc(256),
                  g_bott=gacc(4, 256), g_ln=gacc(2, 256),
                  dproj=None if T == 1005 else torch.empty(B * n, 256, device=dev, dtype=torch.float16))
        add(f"stream_prologue_bwd_T{T}", lambda pa=pa, gX=gX, gb=gb: ops.stream_prologue_bwd(dX0=gX, **pa, **gb), None,
            B * T * 512.0 + (B * n * 12.0 if T == 1005 else 2 * B * n * 512.0), f"B={B} n={n}")
    # UMSE embedding at >= 512k tokens (stable HBM figure), 12 B in + 512 B fp16 out per token
    n_tok = 0 if a.sweep else 1 << 20
    xt = torch.empty(n_tok, 3, device=dev)
    xt[:, 0] = -torch.rand(n_tok, device=dev) * 24; xt[:, 1] = torch.rand(n_tok, device=dev)
    xt[:, 2] = torch.randint(0, 18, (n_tok,), device=dev).float()
    mk = lambda: [torch.randn(256, device=dev), torch.randn(256, device=dev), torch.ones(256, device=dev),
                  torch.zeros(256, device=dev)]
    v4, t4 = mk(), mk()
    Wf = torch.randn(20, 256, device=dev)
    if not a.sweep:
      Bbig, nbig = 512, 1019                     # 524 288 rows of T = 1024: the step's prologue kernel at >= 512 k tokens
      xb = torch.rand(Bbig, nbig, 3, device=dev); xb[:, :, 2] = torch.randint(0, 18, (Bbig, nbig), device=dev).float()
      X0b = torch.empty(Bbig, nbig + 5, 256, device=dev, dtype=torch.float16)
      gln, bln = torch.ones(256, device=dev), torch.zeros(256, device=dev)
      pab = dict(kind=0, B=Bbig, n=nbig, x=xb, val4=v4, proj=None, times=None, n_slots=0, feat_id=0, tim4=t4, Wfeat=Wf,
                 cls=torch.randn(256, device=dev), bottlenecks=torch.randn(4, 256, device=dev), ln_g=gln, ln_b=bln, pe=None,
                 drop_p=0.1, seed=3, salt=7)
      add("stream_prologue_fwd_512k", lambda: ops.stream_prologue_fwd(X0=X0b, **pab), None,
          Bbig * nbig * 12.0 + Bbig * (nbig + 5) * 512.0, f"B={Bbig} n={nbig} ({Bbig * (nbig + 5)} rows)")
      add("umse_embed_fwd_1M", lambda: ops.umse_embed(xt, v4, t4, Wf, torch.float16), None, n_tok * 524.0,
        f"tokens={n_tok}")

    results = []
    flush_buf = torch.zeros(256 << 20, dtype=torch.uint8, device=dev) if a.flush else None
    for name, fn, flops, bytes_, shape in cases:
        fn()
        if not a.time:
            fn()
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if a.flush:
            tot = 0.0
            for _ in range(a.iters):
                flush_buf.view(torch.int32).sum()        # evict the 126 MB L2 with CLEAN lines (a write would leave 126 MB of
                                                         # dirty lines whose write-back the kernel under test then pays for)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            ms = tot / a.iters
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
        r = {"kernel": name, "shape": shape, "ms": round(ms, 4)}
        if flops:
            r["tflops"] = round(flops / ms / 1e9, 1)
        if bytes_:
            r["gbs"] = round(bytes_ / ms / 1e6, 1)
        results.append(r)
        print(json.dumps(r), flush=True)
    torch.cuda.synchronize()
    if a.time:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
