"""Turn gpurun_out/ artefacts into the small tracked summaries under profiles/ (the .ncu-rep files stay in scratch).

    python tools/summarize_profiles.py <tag> [--rep gpurun_out/prof.ncu-rep] [--launches gpurun_out/launches.csv]
                                             [--times gpurun_out/kernel_times.json] [--bench gpurun_out/bench.log]
"""
import argparse, collections, csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct"]


def us(row):
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    return v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--rep"); ap.add_argument("--launches"); ap.add_argument("--times"); ap.add_argument("--bench")
    a = ap.parse_args()
    out = os.path.join(ROOT, "profiles")
    os.makedirs(out, exist_ok=True)
    if a.rep:
        raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        with open(os.path.join(out, f"{a.tag}_ncu_full.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["id", "kernel"] + [f"{k} [{units[idx[k]]}]" for k in KEYS if k in idx])
            for r in rows[2:]:
                w.writerow([r[0], r[idx["Kernel Name"]][:70]] + [r[idx[k]] for k in KEYS if k in idx])
        print("wrote", f"profiles/{a.tag}_ncu_full.csv")
    if a.launches:
        lines = [l for l in open(a.launches) if not l.startswith("==")]
        rows = list(csv.DictReader(lines))
        agg = collections.defaultdict(lambda: [0, 0.0])
        for r in rows:
            k = r["Kernel Name"].split("(")[0][-70:]
            agg[k][0] += 1; agg[k][1] += us(r)
        tot = sum(v[1] for v in agg.values())
        with open(os.path.join(out, f"{a.tag}_launches.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["kernel", "launches", "total_us", "share_pct"])
            for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
                w.writerow([k, c, round(t, 1), round(100 * t / tot, 2)])
        print("wrote", f"profiles/{a.tag}_launches.csv", "total ms", tot / 1e3, "launches", len(rows))
    if a.times:
        with open(a.times) as f:
            data = json.load(f)
        with open(os.path.join(out, f"{a.tag}_kernel_times.json"), "w") as f:
            json.dump(data, f, indent=1)
    if a.bench:
        with open(a.bench) as f:
            lines = [l for l in f if l.startswith("{")]
        with open(os.path.join(out, f"{a.tag}_bench.json"), "w") as f:
            f.write(lines[-1])


if __name__ == "__main__":
    main()
