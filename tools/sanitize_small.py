"""Small-shape launches of the mbarrier / tcgen05 / TMA kernels for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
    compute-sanitizer --tool synccheck python tools/sanitize_small.py
Shapes are tiny (the tools slow kernels down 10-100x) but cover every role of every kernel: multi-tile K loops, M tails,
several work items per persistent CTA, partial key tiles, de-selected samples, fused bias gradient, fp16 dQ reduce."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from medical_tri_modal_pilot_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
which = set(sys.argv[1:]) or {"gemm", "wgrad", "attn", "rows"}
if "gemm" in which:
    for (M, N, K) in [(300, 256, 256), (520, 768, 256), (260, 256, 1024)]:
        A = torch.randn(M, K, device=dev).half(); W = (torch.randn(N, K, device=dev) / 16).half()
        out = torch.empty(M, N, device=dev, dtype=torch.float16)
        am = torch.empty(M, N // 32, device=dev, dtype=torch.int32)
        ops.gemm(A, W, out=out, bias=torch.randn(N, device=dev), relu=True, drop_p=0.1, seed=1, salt=2, mask_out=am)
        ops.gemm(A, W, out=out, gate=am, residual=out.clone())
if "wgrad" in which:
    for (M, N, K) in [(300, 256, 256), (1000, 1024, 256)]:
        dY = torch.randn(M, N, device=dev).half(); X = torch.randn(M, K, device=dev).half()
        ops.gemm_wgrad(dY, X, torch.zeros(N, K, device=dev), dbias=torch.zeros(N, device=dev))
if "attn" in which:
    for (B, T, lens) in [(3, 300, [300, 0, 150]), (2, 133, [133, 7])]:
        qkv = torch.randn(B * T, 768, device=dev).half()
        kv = torch.tensor(lens, device=dev, dtype=torch.int32)
        O = torch.zeros(B * T, 256, device=dev, dtype=torch.float16)
        Tl = ops.lse_len(T)
        lse = torch.zeros(B, 4, Tl, device=dev)
        ops.attn_fwd(qkv, kv, B, T, O, lse)
        dO = torch.randn(B * T, 256, device=dev).half()
        delta = torch.zeros(B, 4, Tl, device=dev)
        dqkv = torch.zeros(B * T, 768, device=dev, dtype=torch.float16)
        dq = torch.zeros(B * T, 256, device=dev)
        ops.attn_bwd(qkv, O, dO, kv, B, T, lse, delta, dq, dqkv)          # stand-alone protocol
        ops.layernorm_bwd_attn(dO, dO, dO, torch.ones(256, device=dev), torch.empty_like(dO), torch.zeros(256, device=dev),
                               torch.zeros(256, device=dev), O, T, delta, dqkv)
        ops.attn_bwd(qkv, O, dO, kv, B, T, lse, delta, None, dqkv)        # fused protocol
        ops.attn_fwd(qkv, kv, B, T, O, lse, q_rows=5)
        ops.attn_bwd(qkv, O, dO, kv, B, T, lse, delta, None, dqkv, q_rows=5)
if "rows" in which:
    x = torch.randn(777, 256, device=dev).half()
    y = torch.empty_like(x); h = torch.empty_like(x)
    g = torch.ones(256, device=dev); b = torch.zeros(256, device=dev)
    ops.layernorm_fwd(x, g, b, y, add=x, sum_out=h)
    ops.layernorm_bwd(x, x, x, g, y, torch.zeros(256, device=dev), torch.zeros(256, device=dev))
torch.cuda.synchronize()
print("sanitize_small: done", sorted(which))
