"""Top stall-sample SASS instructions of one kernel from an .ncu-rep (needs `ncu --set full --import-source on`).
    python tools/ncu_hot.py gpurun_out/x.ncu-rep attn_bwd_kernel [top_n] [launch_skip]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                      f"regex:{kern}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(raw.splitlines())]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; idx = {k: i for i, k in enumerate(h)}
data = [r for r in rows[hi + 1:] if len(r) == len(h) and r[0] != "Address"]
S = lambda r, k: int(float((r[idx[k]] or "0").replace(",", "")))
tot = sum(S(r, "# Samples") for r in data)
print("total samples", tot, "instructions", len(data))
order = sorted(range(len(data)), key=lambda i: -S(data[i], "# Samples"))[:topn]
for i in sorted(order):
    r = data[i]; n = S(r, "# Samples")
    st = {k[6:]: S(r, k) for k in h if k.startswith("stall_") and "Not" not in k}
    main = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print("%5.1f%% #%4d %-72s %s" % (100.0 * n / max(tot, 1), i, r[idx["Source"]][:72], main))
# aggregate by opcode
import collections
agg = collections.Counter(); reasons = collections.Counter()
half = data[: len(data) // 2] if len(data) > 1 and data[0][idx["Source"]] == data[len(data) // 2][idx["Source"]] else data
for r in half:
    op = r[idx["Source"]].replace("@P0", "").replace("@!P0", "").split()[0] if r[idx["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[idx["Source"]].split()[1]
    agg[op] += S(r, "# Samples")
    for k in h:
        if k.startswith("stall_") and "Not" not in k:
            reasons[k[6:]] += S(r, k)
t = sum(agg.values())
print("by opcode:", [(k, round(100.0 * v / t, 1)) for k, v in agg.most_common(14)])
print("by reason:", [(k, round(100.0 * v / t, 1)) for k, v in reasons.most_common(10)])
