"""SASS opcode evidence for the tcgen05 / TMEM / TMA kernels: counts of the Blackwell-native mnemonics per kernel of
libtmp_b200.so (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UTMAREDG/UBLKCP;
legacy mma.sync would show as HMMA).   python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "medical_tri_modal_pilot_b200", "libtmp_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTCBAR", "UTCCP", "SYNCS", "HMMA.",
       "FFMA2", "MUFU.EX2"]


def cuda_tool(name):
    """Path of a CUDA binary utility (PATH first, then next to the toolkit's nvcc); None when the toolkit is absent."""
    import shutil
    return shutil.which(name) or next((p for p in (f"/usr/local/cuda/bin/{name}",) if os.path.exists(p)), None)


def collect(lib=LIB):
    """{kernel name (demangled, no parameter list): Counter of the mnemonics in OPS} for every kernel of the library."""
    out = subprocess.run([cuda_tool("cuobjdump"), "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
            counts[cur] = collections.Counter()
            continue
        if cur:
            for op in OPS:
                if re.search(r"\b" + re.escape(op), line):
                    counts[cur][op] += 1
    return counts


def main():
    counts = collect()
    print(f"{'kernel':58s} " + " ".join(f"{o.rstrip('.'):>8s}" for o in OPS))
    for k, c in counts.items():
        if any(c[o] for o in OPS[:10]):
            print(f"{k[:58]:58s} " + " ".join(f"{c[o]:8d}" for o in OPS))
    legacy = {k: c["HMMA."] for k, c in counts.items() if c["HMMA."]}
    print(f"\nlegacy HMMA (mma.sync / wmma) instructions: {legacy or 0}")
    print("  (window_attn_kernel is the 7x7-window attention of the frozen Swin-T image-encoder feed, SURVEY 8f rank 1: 49 tokens x"
          " d=32 per (window, head) -- below one 128-row tcgen05 tile; every kernel of the fusion encoder itself is tcgen05.)")


if __name__ == "__main__":
    main()
