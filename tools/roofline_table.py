"""One table: every kernel of the step against the roofline that bounds it, from the committed CUDA-event timings
(`profiles/<tag>_kernel_times_coldl2.json`: L2 flushed between launches; `<tag>_kernel_times.json`: back to back) and the
driver-measured peaks (MEASURED_PEAKS.json, else the fallback of /opt/skills/guides/B200_PROFILING.md).

    python tools/roofline_table.py r2e > profiles/r2e_roofline_table.md

Tensor-bound kernels: algorithmic flops (2*M*N*K; attention 4 / 10 * B*H*T^2*64 forward / backward) / time against the BURST
bf16 peak (each kernel is timed alone). HBM-bound kernels: algorithmic bytes (table below, 16-bit activations of width 256
= 512 B per row) / time against the measured copy bandwidth. No GPU needed: it only reads committed files."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FALLBACK = {"hbm_gbs": 6500.0, "bf16_tflops": 1700.0}

# algorithmic bytes per row (or per element pair) of the HBM-bound kernels, as DESIGN.md section 4 states them
ROW_BYTES = {
    "layernorm_fwd": (1024, "x read, y written"),
    "layernorm_fwd_add": (1536, "x, residual read, y written"),
    "layernorm_bwd": (2048, "x, dy, dres read, dx written"),
    "layernorm_bwd_attn": (3072, "x, dy, dres, O read; dx + zeroed dQ written (delta rides along)"),
    "stream_prologue_fwd": (524, "12 B triple read, 512 B row written"),
    "stream_prologue_bwd": (524, "512 B gradient row + 12 B triple read"),
    "umse_embed_fwd": (524, "12 B triple read, 512 B row written"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "MEASURED_PEAKS.json"
    return FALLBACK["hbm_gbs"], FALLBACK["bf16_tflops"], "B200_PROFILING.md fallback"


def rows_of(shape):
    m = re.search(r"rows=(\d+)", shape) or re.search(r"\((\d+) rows\)", shape) or re.search(r"tokens=(\d+)", shape)
    if m:
        return int(m.group(1))
    m = re.search(r"B=(\d+) n=(\d+)", shape)
    if m:
        return int(m.group(1)) * (int(m.group(2)) + 5)      # + 4 bottleneck rows + CLS per sample
    return None


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2e"
    cold = json.load(open(os.path.join(ROOT, "profiles", f"{tag}_kernel_times_coldl2.json")))
    hot = {r["kernel"]: r for r in json.load(open(os.path.join(ROOT, "profiles", f"{tag}_kernel_times.json")))}
    hbm, tf, src = peaks()
    print(f"# Every kernel of the step against its roofline ({tag} build, B200, CUDA events)\n")
    print(f"Peaks ({src}): HBM copy {hbm:.0f} GB/s, bf16 burst {tf:.0f} TFLOP/s. `cold` = clean-line L2 flush between "
          f"launches, `hot` = back to back. Written by `tools/roofline_table.py {tag}` from `{tag}_kernel_times*.json`.\n")
    print("| kernel case | shape | cold µs | hot µs | bound | achieved (cold) | fraction of peak |")
    print("|---|---|---|---|---|---|---|")
    for r in cold:
        name, shape, ms = r["kernel"], r.get("shape", ""), r["ms"]
        hms = hot.get(name, {}).get("ms")
        stem = re.sub(r"_(T\d+|512k|1M)$", "", name)
        if r.get("tflops"):
            bound, ach, frac = "tensor", f"{r['tflops']:.0f} TFLOP/s", r["tflops"] / tf
        elif stem in ROW_BYTES and rows_of(shape):
            gb = ROW_BYTES[stem][0] * rows_of(shape) / (ms * 1e-3) / 1e9
            bound, ach, frac = "hbm", f"{gb:.0f} GB/s", gb / hbm
        elif stem.startswith("colsum"):
            m = re.search(r"M=(\d+) N=(\d+)", shape)
            gb = int(m.group(1)) * int(m.group(2)) * 2 / (ms * 1e-3) / 1e9
            bound, ach, frac = "hbm", f"{gb:.0f} GB/s", gb / hbm
        else:
            continue
        hs = f"{hms * 1e3:.1f}" if hms else ""
        print(f"| `{name}` | {shape[:60]} | {ms * 1e3:.1f} | {hs} | {bound} | {ach} | {frac:.2f} |")
    print("\nAlgorithmic bytes per row of the HBM-bound kernels:\n")
    for k, (b, what) in ROW_BYTES.items():
        print(f"* `{k}`: {b} B ({what})")
    print("\nNotes: the prologue kernels are listed against HBM because their algorithmic work is bytes, but they are "
          "instruction-issue bound (DESIGN.md section 4: 277 / 700 warp instructions per row). `colsum_*` is not launched by the "
          "16-bit step (the bias gradient rides in `wgradb_*`). Kernels changed AFTER this table's build: attention backward "
          "248 -> 219 µs hot-timed alone = 757 TFLOP/s = 0.45 (item order, `r2f_attn_sweep.json`, bench line `r2i_bench.json`), "
          "prologue backward 115 -> 83 µs (`r2i_ncu_full.csv`).")
    print("\nThe `_T152` / `_T133` cases are the image / text streams (9 728 / 8 512 rows): at 10-20 µs they are bound by launch "
          "and fill / drain latency, not by either roofline; they run on side streams next to the vslt kernels.")


if __name__ == "__main__":
    main()
