"""CPU check of the operand algebra of the fp32 parity mode (csrc/precise.cu `split3`, `kPatA`,
`kPatB`): an fp32 value is split into three bf16 terms hi + mid + lo (each the round-to-nearest bf16 of what is left),
both GEMM operands are laid out as six blocks along the reduction dimension so that ONE 16-bit tensor-core GEMM with
fp32 accumulation sums the six partial products of order <= 2^-16. The test restates that layout with torch's CPU
bfloat16 and shows it reproduces an fp64 product to fp32 accuracy, where a single bf16 (or TF32-like 10-bit) operand
rounding is 3-4 orders of magnitude worse -- the reason the mode is bf16x3 and not `kind::tf32`."""
import re
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _patterns():
    """kPatA / kPatB as the kernel source states them (the test follows the source, not a copy of the numbers)."""
    src = open(os.path.join(ROOT, "medical_tri_modal_pilot_b200", "csrc", "precise.cu")).read()
    pa = re.search(r"kPatA\[6\]\s*=\s*\{([^}]*)\}", src).group(1)
    pb = re.search(r"kPatB\[6\]\s*=\s*\{([^}]*)\}", src).group(1)
    return [int(v) for v in pa.split(",")], [int(v) for v in pb.split(",")]


def split3(x):
    hi = x.to(torch.bfloat16)
    r1 = x - hi.float()
    mid = r1.to(torch.bfloat16)
    r2 = r1 - mid.float()
    lo = r2.to(torch.bfloat16)
    return [hi, mid, lo]


def six(x, pat):
    t = split3(x)
    return torch.cat([t[k] for k in pat], dim=1)


def test_split_is_exact_to_24_bits():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4096, generator=g) * torch.logspace(-6, 6, 4096)
    hi, mid, lo = split3(x)
    back = hi.float().double() + mid.float().double() + lo.float().double()
    assert torch.all((back - x.double()).abs() <= x.double().abs() * 2.0 ** -23)


def test_patterns_cover_every_product_down_to_2_pow_minus_16():
    pa, pb = _patterns()
    pairs = sorted(zip(pa, pb))
    # (term of A, term of B) with order(A) + order(B) <= 2: hi.hi, hi.mid, mid.hi, hi.lo, lo.hi, mid.mid
    assert pairs == sorted([(0, 0), (0, 1), (1, 0), (0, 2), (2, 0), (1, 1)])


def test_six_block_gemm_reproduces_fp32_accuracy():
    pa, pb = _patterns()
    g = torch.Generator().manual_seed(1)
    M, N, K = 96, 80, 256
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g) * 0.05
    ref = A.double() @ B.double().t()
    A6, B6 = six(A, pa), six(B, pb)
    assert A6.shape == (M, 6 * K) and B6.shape == (N, 6 * K)
    # products of two bf16 values are exact in fp32; accumulate like the tensor core does (fp32, here in fp64 to isolate the
    # operand error from the accumulation order)
    got = A6.float().double() @ B6.float().double().t()
    scale = ref.abs().max()
    err3 = ((got - ref).abs().max() / scale).item()
    one = A.to(torch.bfloat16).float().double() @ B.to(torch.bfloat16).float().double().t()
    err1 = ((one - ref).abs().max() / scale).item()
    assert err3 < 3e-7, err3                 # fp32-level: 24 mantissa bits on both operands
    assert err1 > 1e3 * err3, (err1, err3)   # a single 16-bit rounding of the operands is >= 1000x worse
    # and with fp32 accumulation end to end
    got32 = A6.float() @ B6.float().t()
    assert ((got32.double() - ref).abs().max() / scale).item() < 2e-6


def test_row_stacked_layout_for_the_weight_gradient():
    """dW = dY^T X sums over tokens: the six blocks are stacked along the token dimension (stack_rows = 1)."""
    pa, pb = _patterns()
    g = torch.Generator().manual_seed(2)
    M, N, K = 512, 32, 48
    dY = torch.randn(M, N, generator=g) * 1e-3
    X = torch.randn(M, K, generator=g)
    ty, tx = split3(dY), split3(X)
    dY6 = torch.cat([ty[k] for k in pa], dim=0)      # [6M, N]
    X6 = torch.cat([tx[k] for k in pb], dim=0)       # [6M, K]
    got = dY6.float().double().t() @ X6.float().double()
    ref = dY.double().t() @ X.double()
    assert ((got - ref).abs().max() / ref.abs().max()).item() < 3e-7
