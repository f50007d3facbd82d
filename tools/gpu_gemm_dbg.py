"""Which GEMM shape / epilogue variant is wrong, and which output boxes (debug helper)."""
import os, sys
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medical_tri_modal_pilot_b200 import ops
dev = "cuda"
torch.manual_seed(4)
for (M, N, K) in [(128, 128, 64), (300, 256, 256), (4096, 768, 256)]:
    dt = torch.float16
    A = torch.randn(M, K, device=dev).to(dt); Bw = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
    ref = A.float() @ Bw.float().t()
    out = torch.full((M, N), 777.0, device=dev, dtype=dt)
    ops.gemm(A, Bw, out=out); torch.cuda.synchronize()
    o = out.float()
    unwritten = (o == 777.0)
    bad = (~unwritten) & ((o - ref).abs() > 0.05 * ref.abs().max())
    print(M, N, K, "unwritten frac %.3f wrong frac %.3f" % (unwritten.float().mean().item(), bad.float().mean().item()))
    # pattern inside the first 128x128 tile: per 16-byte chunk (8 cols) of each row
    t = (unwritten | bad)[:min(M, 64), :64]
    for r in range(0, min(M, 64), 1):
        print("%3d " % r + "".join("X" if t[r, c * 8:(c + 1) * 8].any().item() else "." for c in range(8)))
    if M >= 128:
        # where do the values of row 0..3, cols 0..63 of ref appear in out?
        for r in range(4):
            for c in (0, 8, 32):
                v = ref[r, c].item()
                hit = ((o[:128, :128] - v).abs() < 2e-3 * max(1.0, abs(v))).nonzero()
                print("ref[%d,%d]=%.3f found at %s" % (r, c, v, hit[:3].tolist()))
    break
