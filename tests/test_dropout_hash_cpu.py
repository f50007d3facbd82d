"""Statistics of the stateless dropout mask (csrc/tc05.cuh: dropout_key / dropout_bits / dropout_keep_lo / _hi), restated
in numpy. The CUDA kernels evaluate the same function of (seed, salt, element index) in the forward and the backward pass
(tests/test_kernels_gpu.py checks determinism and the kept fraction on the device); here: the kept fraction for the
reference's p = 0.1 (control/config.py --dropout), independence inside a pair, between neighbouring pairs, rows, steps
(seed + step counter, trainer.GraphedStep) and tensors (salt)."""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)
U = np.uint64


def _mix32(x):
    x = U(x)
    x ^= x >> U(16); x = (x * U(0x7feb352d)) & M32
    x ^= x >> U(15); x = (x * U(0x846ca68b)) & M32
    x ^= x >> U(16)
    return x


def _key(seed, salt):
    return (_mix32(seed) ^ ((U(salt) * U(0x85ebca6b)) & M32)) & M32


def keep_mask(seed, salt, n, thr16):
    idx = np.arange(n, dtype=np.uint64)
    h = ((idx >> U(1)) * U(0x9E3779B1) + _key(seed, salt)) & M32
    h ^= h >> U(16)
    lo = ((h * U(0x7feb352d)) & M32) >= U(thr16 << 16)
    hi = ((h * U(0x846ca68b)) & M32) >= U(thr16 << 16)
    return np.where(idx & U(1), hi, lo)


def test_keep_fraction_and_independence():
    p = 0.1
    thr = int(p * 65536 + 0.5)
    n = 1 << 21
    drop = ~keep_mask(3, 9, n, thr)
    sig = (p * (1 - p) / n) ** 0.5
    assert abs(drop.mean() - thr / 65536) < 5 * sig
    lo, hi = drop[0::2], drop[1::2]
    tol = 6 * (p * p / (n / 2)) ** 0.5 + 3e-4
    assert abs((lo & hi).mean() - p * p) < tol                      # the two elements of a pair
    assert abs((lo[:-1] & lo[1:]).mean() - p * p) < tol             # neighbouring pairs
    rows = drop.reshape(-1, 1024)
    assert abs((rows[:-1] & rows[1:]).mean() - p * p) < tol         # same column, neighbouring rows
    assert abs((drop & ~keep_mask(4, 9, n, thr)).mean() - p * p) < tol     # next step (seed + 1)
    assert abs((drop & ~keep_mask(3, 10, n, thr)).mean() - p * p) < tol    # another tensor (salt + 1)
    col = rows.mean(0)
    assert abs(col - p).max() < 6 * (p * (1 - p) / rows.shape[0]) ** 0.5   # no column is favoured


def test_threshold_extremes():
    n = 1 << 12
    assert keep_mask(1, 2, n, 0).all()                              # p = 0: everything kept
    assert keep_mask(1, 2, n, 65535).mean() < 1e-3                  # p -> 1: (almost) nothing kept
