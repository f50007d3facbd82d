"""Seeded synthetic batches of the sample_data shapes (shapes as produced by the reference
`Dataset.__getitem__`, builder/data/dataset_new.py:581-788): the workload of bench.py and the input of the parity
tests. numpy Generator only, so a batch is reproducible on any box."""
from __future__ import annotations

import numpy as np
import torch

# feature-id histogram of data/sample_data: ids 0,1,3,4,5 ~17 % each; 2,6 ~4 % each; 7..16 ~1 % each
_FEAT_P = np.array([17, 17, 4, 17, 17, 17, 4] + [1] * 10 + [0], dtype=np.float64)
_FEAT_P /= _FEAT_P.sum()


def make_batch(B: int, L: int, n_img: int = 3, seed: int = 0, full_length: bool = False, missing_mode: str = "mixed",
               with_pixels: bool = False, feats: bool = True) -> dict:
    """Returns CPU tensors.
    x[B,L,3] (time,value,feat) zero-padded past input_lengths; age, gen [B]; input_lengths [B] int64;
    txts[B,128,768] zero rows past txt_lengths; txt_lengths [B] int64 (0 when txt missing);
    img_time [B,n_img] (sentinel 10 = empty slot) or [B] when n_img == 1; txt_time [B];
    img_feats [B*n_img,49,768] (stand-in for the frozen Swin output) and/or img [B,n_img,1,224,224];
    missing [B] int64 codes (0 all, 1 txt missing, 2 img missing, 3 both); y [B]."""
    g = np.random.Generator(np.random.PCG64(seed))
    if missing_mode == "none":
        missing = np.zeros(B, dtype=np.int64)
    elif missing_mode == "mixed":
        missing = g.integers(0, 4, B)
        missing[: min(B, 4)] = np.arange(min(B, 4))        # every code present in every batch of >= 4
    elif missing_mode == "img_missing":
        missing = np.full(B, 2, dtype=np.int64)
    else:
        raise ValueError(missing_mode)
    lens = np.full(B, L, dtype=np.int64) if full_length else g.integers(max(1, L // 4), L + 1, B)
    lens[0] = L
    x = np.zeros((B, L, 3), dtype=np.float32)
    for b in range(B):
        n = int(lens[b])
        t = -g.uniform(0, 24, n)
        n_cf = min(18, n // 4)                              # carry-forward rows reaching back to -72 h
        if n_cf:
            t[:n_cf] = -g.uniform(24, 72, n_cf)
        x[b, :n, 0] = t
        x[b, :n, 1] = g.uniform(0, 1, n)
        x[b, :n, 2] = g.choice(18, size=n, p=_FEAT_P).astype(np.float32)
    age = g.uniform(0, 1, B).astype(np.float32)
    gen = g.integers(0, 2, B).astype(np.float32)
    txt_missing = (missing == 1) | (missing == 3)
    img_missing = (missing == 2) | (missing == 3)
    txt_len = g.integers(1, 127, B)
    txt_len[txt_missing] = 0
    if B > 4 and not txt_missing[4]:
        txt_len[4] = 126
    txts = np.zeros((B, 128, 768), dtype=np.float32)
    for b in range(B):
        txts[b, : txt_len[b]] = g.standard_normal((int(txt_len[b]), 768)).astype(np.float32)
    txt_time = -g.integers(3, 169, B).astype(np.float32)
    k_img = g.integers(1, n_img + 1, B)
    k_img[img_missing] = 0
    img_time = np.full((B, n_img), 10.0, dtype=np.float32)
    for b in range(B):
        img_time[b, : k_img[b]] = -g.uniform(0, 24, int(k_img[b]))
    out = {
        "x": torch.from_numpy(x), "age": torch.from_numpy(age), "gen": torch.from_numpy(gen),
        "input_lengths": torch.from_numpy(lens), "txts": torch.from_numpy(txts),
        "txt_lengths": torch.from_numpy(txt_len.astype(np.int64)),
        "img_time": torch.from_numpy(img_time if n_img > 1 else img_time[:, 0].copy()),
        "txt_time": torch.from_numpy(txt_time), "missing": torch.from_numpy(missing.astype(np.int64)),
        "y": torch.from_numpy((g.uniform(0, 1, B) < 0.2).astype(np.float32)),
    }
    if feats:
        f = g.standard_normal((B * n_img, 49, 768)).astype(np.float32)
        out["img_feats"] = torch.from_numpy(f)
    if with_pixels:
        img = g.uniform(0, 1, (B, n_img, 1, 224, 224)).astype(np.float32)
        for b in range(B):
            img[b, k_img[b]:] = 0
        out["img"] = torch.from_numpy(img if n_img > 1 else img[:, 0])
    return out
