"""Static resource summary of every kernel in csrc/ (registers, spills, static shared memory) from `ptxas -v`.

Runs without a GPU (nvcc cross-compiles sm_100a):  python tools/ptxas_summary.py > profiles/r2_ptxas_resources.txt
Same flags as medical_tri_modal_pilot_b200/build.py, plus -Xptxas -v.  Dynamic shared memory is set at launch and not
shown by ptxas; it is listed per kernel in DESIGN.md section 4.
"""
import os
import re
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from medical_tri_modal_pilot_b200.build import CSRC, NVCC_FLAGS, _nvcc  # noqa: E402

FUNC = re.compile(r"Compiling entry function '([^']+)' for 'sm_100a'")
STACK = re.compile(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads")
USED = re.compile(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?")


def demangle(names):
    r = subprocess.run(["cu++filt", *names], capture_output=True, text=True)
    out = r.stdout.strip().splitlines() if r.returncode == 0 else names
    return out if len(out) == len(names) else names


def one(src):
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [_nvcc(), *NVCC_FLAGS, "-Xptxas", "-v", "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", os.path.join(tmp, "o.o")]
        r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    rows, cur = [], None
    for line in r.stderr.splitlines():
        m = FUNC.search(line)
        if m:
            cur = {"name": m.group(1), "stack": 0, "sst": 0, "sld": 0, "regs": 0, "bars": 0, "smem": 0}
            rows.append(cur)
            continue
        if cur is None:
            continue
        m = STACK.search(line)
        if m:
            cur["stack"], cur["sst"], cur["sld"] = map(int, m.groups())
        m = USED.search(line)
        if m:
            cur["regs"] = int(m.group(1))
            cur["bars"] = int(m.group(2) or 0)
            cur["smem"] = int(m.group(3) or 0)
    return src, rows


def short(name, width=86):
    """'void <unnamed>::k<(int)0, (bool)1>(Params)' -> 'k<0, 1>' (parameter list = the last top-level parenthesis)."""
    name = re.sub(r"^void ", "", name).replace("<unnamed>::", "")
    depth = 0
    for i, c in enumerate(name):
        if c == "<":
            depth += 1
        elif c == ">":
            depth -= 1
        elif c == "(" and depth == 0:
            name = name[:i]
            break
    name = re.sub(r"\((?:int|bool|unsigned int)\)", "", name)
    return name if len(name) <= width else name[: width - 3] + "..."


def main():
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    with ThreadPoolExecutor(8) as ex:
        results = list(ex.map(one, srcs))
    print("ptxas -v, sm_100a, flags of build.py; regs = registers per thread, bars = named barriers, smem = STATIC shared bytes")
    print(f"{'kernel':86s} {'regs':>5s} {'bars':>4s} {'smem':>7s} {'stack':>6s} {'spill st/ld':>12s}")
    spilled = []
    for src, rows in results:
        if not rows:
            continue
        print(f"-- {src}")
        names = demangle([r["name"] for r in rows])
        for r, n in zip(rows, names):
            print(f"{short(n):86s} {r['regs']:5d} {r['bars']:4d} {r['smem']:7d} {r['stack']:6d} {r['sst']:5d}/{r['sld']:<5d}")
            if r["sst"] or r["sld"]:
                spilled.append((short(n, 60), r["sst"], r["sld"]))
    print()
    if spilled:
        print("kernels with register spills (bytes per thread):")
        for n, a, b in spilled:
            print(f"  {n}: {a} stored / {b} loaded")
    else:
        print("no kernel spills registers")


if __name__ == "__main__":
    main()
