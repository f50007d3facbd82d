"""CPU tests of bench.py's host logic (no GPU, no timing): the contract keys that do not need a device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_profiled_traffic_reads_the_committed_ncu_summary():
    """roofline.traffic = dram read + write bytes of one attn_bwd_kernel launch from profiles/*_ncu_full.csv."""
    t, src = bench.profiled_traffic("attn_bwd_kernel")
    assert src is not None and src.endswith("_ncu_full.csv")
    # algorithmic bytes at B=64, T=1005 in the fused protocol (delta and the zeroed dQ columns come from the LayerNorm
    # backward): Q,K,V,dO read once + dQ,dK,dV written = 64320 rows * (768 + 256 + 768) * 2 B; the profiled launch may sit a
    # few % below (lines still in L2) or above (lse / delta, partial sectors)
    algorithmic = 64320 * (768 + 256 + 768) * 2
    assert 0.9 * algorithmic <= t <= 1.25 * algorithmic, (t, algorithmic)
    assert bench.profiled_traffic("no_such_kernel") == (None, None)


def test_peaks_come_from_measured_file_or_fallback():
    peaks, src = bench.load_peaks()
    assert src in ("measured", "fallback")
    assert 3000 < peaks["hbm_gbs"] < 9000 and 800 < peaks["bf16_tflops_sustained"] <= peaks["bf16_tflops"] < 2500


def test_b200_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: the product arm of the bench refuses to run without a CUDA device."""
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_committed_bench_lines_carry_the_contract_keys():
    for name in ("r1i_bench.json", "r1i_bench_n2.json"):
        with open(os.path.join(ROOT, "profiles", name)) as f:
            d = json.loads(f.read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert k in d, (name, k)
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["gpu_launches"] > 0
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
        assert "workload" in d["config"]
