"""Helpers shared by the golden-fixture tests (fixtures are produced by tools/make_golden.py from the reference)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixture_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def gpu_fixture_names():
    """Every fixture is part of the GPU parametrisation (the 6-layer bench depth and the B=64 / TIE-len 1000 bench shape
    included)."""
    return fixture_names()


def emb_rows(B, L):
    """(b, l) pairs at which the large fixtures store the embedding rows (same generator as tools/make_golden.py)."""
    g = np.random.Generator(np.random.PCG64(B * 100003 + L))
    return g.integers(0, B, 1024), g.integers(0, L, 1024)


def fixture_embedding_view(fx, e):
    """`e` [B,L,256] computed here -> the part the fixture holds (everything, or the seeded rows of a large fixture)."""
    if "emb_subsampled" in fx and int(fx["emb_subsampled"]):
        return e[emb_rows(e.shape[0], e.shape[1])]
    return e


def load_fixture(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def fixture_inputs(fx):
    """Regenerate (state_dict, batch, cfg) of a fixture from its seeds."""
    from oracle import synth, weights
    from oracle.tri_mbt_oracle import OracleConfig
    nl, multi, B, L, bseed, wseed = (int(v) for v in fx["config"])
    sd = weights.make_state_dict(nl, wseed)
    batch = synth.make_batch(B, L, n_img=3 if multi else 1, seed=bseed, missing_mode=str(fx["missing_mode"]))
    return sd, batch, OracleConfig(n_layers=nl, multiimages=multi)


def grad_probe(name, g):
    """Same seeded samples / projection as tools/make_golden.py."""
    g = np.asarray(g, dtype=np.float64).ravel()
    rng = np.random.Generator(np.random.PCG64(len(name) * 7919 + g.size))
    idx = rng.integers(0, g.size, 32)
    proj = rng.standard_normal(g.size)
    return np.linalg.norm(g), g[idx], float(g @ proj)


GEMM_WEIGHT_KEYS = ("_proj.linear.weight", "feed_forward.w_1.weight", "feed_forward.w_2.weight")


def fp16_representable(sd):
    """Copy of a state_dict whose tensor-core GEMM weights (QKV, FFN, the two 768->256 projections) are rounded to
    fp16-representable values. The B200 path feeds fp16 copies of exactly these tensors to tcgen05 (fp32 masters, fp32
    biases / LayerNorm parameters), so with such a state_dict the oracle (fp32 math) and the kernels see IDENTICAL
    weights. Why it matters: on the random-init fixtures the fp32 oracle's own gradients move by up to cos 0.975
    (layer_stacks.1.0.feed_forward_prenorm.beta) when only these weights are rounded to fp16 -- the sums over tokens
    cancel heavily, and a weight perturbation is coherent across tokens -- so a gradient comparison at non-identical
    weights measures that sensitivity, not the kernels."""
    out = {}
    for k, v in sd.items():
        if k.endswith(GEMM_WEIGHT_KEYS) or k in ("txt_embedding.weight", "linear.weight"):
            out[k] = v.half().float()
        else:
            out[k] = v
    return out
