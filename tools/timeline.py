"""Kernel timeline of one training step at the bench workload (multi-stream, as the bench runs it), from the CUPTI
activity records torch.profiler collects: GPU busy time (union over streams), per-stream busy time, the largest idle
gaps with the kernels around them, and how much kernel time ran concurrently with another kernel.

    python tools/timeline.py [--graph] [--out gpurun_out/timeline.json]

Not a benchmark: the profiler adds launch overhead on the host; the eager step is host-issue bound under it, the graph
replay is not (one launch).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from builder.models import get_model  # noqa: E402
from medical_tri_modal_pilot_b200 import synth, trainer  # noqa: E402
from medical_tri_modal_pilot_b200.config import make_args  # noqa: E402
from medical_tri_modal_pilot_b200.optim import FlatAdamW  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tie-len", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--graph", action="store_true", help="profile a CUDA-graph replay instead of an eager step")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.json"))
    ap.add_argument("--steps", type=int, default=1, help="profile this many back-to-back steps (steady state)")
    ap.add_argument("--dump", action="store_true", help="also write every kernel record (name, stream, start us, us)")
    a = ap.parse_args()
    # under torchrun (WORLD_SIZE > 1): the data-parallel step (GradSync over NCCL); rank 0 writes its own timeline
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        trainer.ddp_setup_env()
        dist.init_process_group("nccl", device_id=dev, pg_options=trainer.ddp_pg_options())
    args = make_args(transformer_num_layers=6, multiimages=1, mbt_only_vslt=1, input_types="vslt_img_txt",
                     imgtxt_time=1, dropout=0.1, batch_size=a.batch, img_pretrain="No", TIE_len=a.tie_len)
    args.device = dev
    torch.manual_seed(0)
    model = get_model(args)(args).to(dev).train()
    if world > 1:
        trainer.GradSync(model)
    opt = FlatAdamW(model, lr=1e-4, weight_decay=1e-6)
    crit = torch.nn.BCEWithLogitsLoss()
    host = synth.make_batch(a.batch, a.tie_len, n_img=3, seed=1000 + rank, full_length=True, missing_mode="none",
                            with_pixels=True, feats=False)
    miss = host["missing"]
    host["missing3"] = torch.stack([torch.zeros_like(miss), (miss >= 2).long(), (miss % 2).long()], 1).float()
    host["static"] = torch.stack([host["gen"], host["age"]], 1)
    r = {k: v.to(dev) for k, v in host.items()}
    prepared = trainer.prepare_batch(args, dev, r["x"], r["static"], r["input_lengths"], r["y"], r["img"], r["txts"],
                                     r["txt_lengths"], (r["img_time"], r["txt_time"]), r["missing3"])
    if a.graph:
        raw = dict(zip(trainer._RAW_KEYS, (r["x"], r["static"], r["input_lengths"], r["y"], r["img"], r["txts"],
                                           r["txt_lengths"], r["img_time"], r["txt_time"], r["missing3"])))
        gs = trainer.graphed_step(args, model, opt, crit, raw)
        gs.load(raw)
        step = lambda i: gs.step(None, i)
    else:
        step = lambda i: trainer.train_step(args, model, opt, crit, prepared, None, i, None)
    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(a.steps):
            step(5 + i)
        torch.cuda.synchronize()
    if rank != 0:
        import torch.distributed as dist
        dist.barrier()
        model.grad_sync.close()
        os._exit(0)
    tmp = a.out + ".trace.json"
    prof.export_chrome_trace(tmp)
    ev = json.load(open(tmp))["traceEvents"]
    os.remove(tmp)
    ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    ks.sort(key=lambda e: e["ts"])
    t0 = ks[0]["ts"]
    t1 = max(e["ts"] + e["dur"] for e in ks)
    # union of busy intervals + concurrency
    pts = []
    for e in ks:
        pts.append((e["ts"], 1))
        pts.append((e["ts"] + e["dur"], -1))
    pts.sort()
    busy = conc = 0.0
    depth = 0
    last = pts[0][0]
    for t, d in pts:
        if depth >= 1:
            busy += t - last
        if depth >= 2:
            conc += t - last
        depth += d
        last = t
    streams = {}
    for e in ks:
        s = e["args"].get("stream", -1)
        streams.setdefault(s, [0.0, 0])
        streams[s][0] += e["dur"]
        streams[s][1] += 1
    # idle gaps
    gaps = []
    end = ks[0]["ts"] + ks[0]["dur"]
    prev = ks[0]
    for e in ks[1:]:
        if e["ts"] > end:
            gaps.append((e["ts"] - end, end - t0, prev["name"][:60], e["name"][:60]))
        if e["ts"] + e["dur"] > end:
            end = e["ts"] + e["dur"]
            prev = e
    gaps.sort(reverse=True)
    by_name = {}
    for e in ks:
        n = e["name"].replace("(anonymous namespace)::", "").replace("void ", "")
        n = n.split("(")[0][:70]
        by_name.setdefault(n, [0.0, 0])
        by_name[n][0] += e["dur"]
        by_name[n][1] += 1
    out = {
        "mode": "graph replay" if a.graph else "eager",
        "steps": a.steps, "span_ms": (t1 - t0) / 1e3, "busy_ms": busy / 1e3, "idle_ms": (t1 - t0 - busy) / 1e3,
        "concurrent_ms": conc / 1e3, "kernel_sum_ms": sum(e["dur"] for e in ks) / 1e3, "n_kernels": len(ks),
        "streams": {str(k): {"busy_ms": v[0] / 1e3, "n": v[1]} for k, v in streams.items()},
        "top_gaps_us": [{"gap": g[0], "at_ms": g[1] / 1e3, "after": g[2], "before": g[3]} for g in gaps[:25]],
        "n_gaps": len(gaps), "gaps_over_5us": sum(1 for g in gaps if g[0] > 5), "gap_sum_over_5us_ms": sum(g[0] for g in gaps if g[0] > 5) / 1e3,
        "top_kernels": sorted(([k, round(v[0] / 1e3, 3), v[1]] for k, v in by_name.items()), key=lambda x: -x[1])[:40],
    }
    # low-occupancy stretches: time during which only "small" kernels (grid < 64 blocks) are running -- the serial head /
    # tail of the step (classifier head, loss, optimizer glue), where most SMs idle although the GPU counts as busy
    def _blocks(e):
        g = e["args"].get("grid", [1 << 20])
        return g[0] * (g[1] if len(g) > 1 else 1) * (g[2] if len(g) > 2 else 1)

    def small(e):
        return _blocks(e) < 64
    pts2 = []
    for e in ks:
        pts2.append((e["ts"], 1, small(e)))
        pts2.append((e["ts"] + e["dur"], -1, small(e)))
    pts2.sort(key=lambda x: (x[0], x[1]))
    big = sm = 0
    last = pts2[0][0]
    small_only = 0.0
    runs = []
    run_start = None
    for t, d_, is_small in pts2:
        if big == 0 and sm > 0:
            small_only += t - last
            if run_start is None:
                run_start = last
        elif run_start is not None:
            runs.append((last - run_start, (run_start - t0) / 1e3))
            run_start = None
        if is_small:
            sm += d_
        else:
            big += d_
        last = t
    runs.sort(reverse=True)
    out["small_kernels_only_ms"] = small_only / 1e3
    out["small_only_runs"] = [{"us": r[0], "at_ms": r[1]} for r in runs[:12]]
    seq = []
    for e in ks:
        if small(e):
            seq.append([round((e["ts"] - t0) / 1e3, 3), round(e["dur"], 1), e["name"].replace("(anonymous namespace)::", "")[:80]])
    out["small_kernel_sequence"] = seq
    out["world"] = world
    if a.dump:
        out["kernels"] = [[e["name"].replace("(anonymous namespace)::", "").replace("void ", "")[:48], e["args"].get("stream", -1),
                           round(e["ts"] - t0, 1), round(e["dur"], 1), _blocks(e)] for e in ks]
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k not in ("top_gaps_us", "top_kernels", "small_kernel_sequence", "kernels")}))
    for g in out["top_gaps_us"][:15]:
        print(f"gap {g['gap']:7.1f} us at {g['at_ms']:7.3f} ms  after {g['after']}  before {g['before']}")
    for k in out["top_kernels"][:30]:
        print(f"{k[1]:8.3f} ms  n={k[2]:4d}  {k[0]}")
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        model.grad_sync.close()
        os._exit(0)


if __name__ == "__main__":
    main()
