#!/usr/bin/env python
"""bench.py -- training throughput (samples/s) of `tri_mbt_vsltcls` vslt_img_txt on N x B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # B200-native arm (this repo)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] # the reference algorithm on the host CPU cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU (NCCL), weak scaling

A "step" = one optimisation step (frozen Swin-T image encoder forward, fused UMSE/MBT encoder forward + backward,
classifier head, BCE loss, gradient all-reduce when N > 1, AdamW) on one synthetic batch of the sample_data shapes.
Rank 0 prints ONE JSON line. `value` is measured with the batch resident in HBM; `e2e` goes through the reference's
own user call (`builder.trainer.get_trainer`, reference 2_train.py:177-200) with pinned HOST tensors, so the
host->device copies and the `loss.item()` read-back are inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}   # B200_PROFILING.md


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--batch", type=int, default=64, help="per-GPU batch (--batch-size)")
    p.add_argument("--tie-len", type=int, default=1000)
    p.add_argument("--layers", type=int, default=6)
    p.add_argument("--multiimages", type=int, default=1)
    p.add_argument("--dropout", type=float, default=0.1)
    p.add_argument("--realistic", action="store_true", help="ragged lengths + mixed missing codes instead of full")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU budget of the cpu_baseline sample")
    p.add_argument("--optimizer", default="fused", choices=["fused", "torch"])
    p.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured CUDA graph")
    return p.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                d = json.load(f)
            out = dict(FALLBACK_PEAKS)
            for k in out:
                if isinstance(d.get(k), (int, float)):
                    out[k] = float(d[k])
            return out, "measured"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), "fallback"


def workload_name(a):
    return (f"tri_mbt_vsltcls --input-types vslt_img_txt --vslt-type TIE --imgtxt-time 1 --multiimages {a.multiimages} "
            f"--transformer-num-layers {a.layers} --TIE-len {a.tie_len} --mbt-only-vslt 1 --batch-size {a.batch} "
            f"--dropout {a.dropout}")


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w": statistics.median(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference algorithm (CPU oracle port + stock Swin-T) on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference(a, steps, warmup, budget_s=None):
    """Times the CPU restatement of the reference training step (oracle/, kind "port": the Python reference cannot
    travel to the GPU box) on a bounded sample of the same workload: per-step batch B_s <= --batch, same L / layers /
    images. Returns (samples_per_s, ms_per_step, info)."""
    import torch
    from medical_tri_modal_pilot_b200 import synth
    from medical_tri_modal_pilot_b200.model import build_swin_t_m
    from oracle import tri_mbt_oracle as O
    from oracle import weights

    threads = os.cpu_count() or 1
    try:
        threads = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(threads)
    n_img = 3 if a.multiimages else 1
    cfg = O.OracleConfig(n_layers=a.layers, multiimages=a.multiimages)
    sd = weights.make_state_dict(a.layers, seed=0)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point
              and "running" not in k and "positional_encoding" not in k}
    full = dict(sd); full.update(leaves)
    opt = torch.optim.AdamW(list(leaves.values()), lr=1e-4, weight_decay=1e-6)
    swin = build_swin_t_m().eval()

    def make(Bs, seed):
        return synth.make_batch(Bs, a.tie_len, n_img=n_img, seed=seed, full_length=not a.realistic,
                                missing_mode="mixed" if a.realistic else "none", with_pixels=True, feats=False)

    def one_step(batch):
        t0 = time.perf_counter()
        opt.zero_grad()
        with torch.no_grad():                                               # tri_mbt_vsltcls.py:205-209
            f = swin(batch["img"].reshape(-1, 1, 224, 224))
        batch = dict(batch); batch["img_feats"] = f.reshape(f.shape[0], 49, 768)
        loss = O.loss_fn(O.forward(full, batch, cfg), batch["y"])
        loss.backward()
        opt.step()
        return time.perf_counter() - t0

    # size the sample: probe with B_s = 2, then pick B_s so that (steps + warmup) steps fit the budget
    probe = one_step(make(2, 100))
    probe = min(probe, one_step(make(2, 101)))
    total = steps + warmup
    if budget_s is None:
        budget_s = 150.0
    per_sample = probe / 2.0
    Bs = int(max(2, min(a.batch, budget_s / max(total, 1) / max(per_sample, 1e-6))))
    batch = make(Bs, 7)
    for _ in range(warmup):
        one_step(batch)
    ts = [one_step(batch) for _ in range(steps)]
    dt = sum(ts) / len(ts)
    info = {"cores": threads, "kind": "port",
            "sample": f"{steps} steps of a B={Bs} slice of the workload batch (same L={a.tie_len}, {a.layers} layers, "
                      f"{n_img} images/sample, fp32 torch-CPU oracle + stock Swin-T + AdamW), {warmup} warm-up"}
    return Bs / dt, dt * 1e3, info


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, ms, info = cpu_reference(a, a.steps, a.warmup)
    line = {"impl": "reference", "metric": "train_samples_per_sec", "value": v, "unit": "samples/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "device": "host CPU"},
            "cpu_baseline": {"value": v, "unit": "samples/s", **info},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist
    from builder.models import get_model
    from builder.trainer import GradSync, get_trainer
    from medical_tri_modal_pilot_b200 import _lib, ops, synth, trainer
    from medical_tri_modal_pilot_b200.config import make_args
    from medical_tri_modal_pilot_b200.optim import FlatAdamW

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (B200 arm) needs a GPU; there is no CPU fallback")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    n_img = 3 if a.multiimages else 1
    args = make_args(transformer_num_layers=a.layers, multiimages=a.multiimages, mbt_only_vslt=1,
                     input_types="vslt_img_txt", imgtxt_time=1, dropout=a.dropout, batch_size=a.batch,
                     img_pretrain="No", modality_inclusion="train-missing_test-missing", TIE_len=a.tie_len)
    args.device = dev
    args.cuda_graph = not a.no_graph and a.optimizer == "fused"
    torch.manual_seed(0)
    model = get_model(args)(args).to(dev)
    model.train()
    if world > 1:
        GradSync(model)
    if a.optimizer == "fused":
        optimizer = FlatAdamW(model, lr=1e-4, weight_decay=1e-6)
    else:
        optimizer = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-6)     # reference 2_train.py:110
    criterion = torch.nn.BCEWithLogitsLoss()

    host = synth.make_batch(a.batch, a.tie_len, n_img=n_img, seed=1000 + rank, full_length=not a.realistic,
                            missing_mode="mixed" if a.realistic else "none", with_pixels=True, feats=False)
    miss = host["missing"]
    host["missing3"] = torch.stack([torch.zeros_like(miss), (miss >= 2).long(), (miss % 2).long()], 1).float()
    host["static"] = torch.stack([host["gen"], host["age"]], 1)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    h2d_keys = ["x", "static", "input_lengths", "y", "img", "txts", "txt_lengths", "img_time", "txt_time", "missing3"]
    h2d_bytes = sum(pinned[k].numel() * pinned[k].element_size() for k in h2d_keys)

    def call_trainer(src, it):
        return get_trainer(args, it, src["x"], src["static"], src["input_lengths"], src["y"], None, model, None, dev,
                           None, optimizer, criterion, x_txt=src["txts"], x_img=src["img"],
                           txt_lengths=src["txt_lengths"], imgtxt_time=(src["img_time"], src["txt_time"]),
                           missing=src["missing3"], flow_type="train")

    resident = {k: pinned[k].to(dev) for k in h2d_keys}
    prepared = trainer.prepare_batch(args, dev, resident["x"], resident["static"], resident["input_lengths"],
                                     resident["y"], resident["img"], resident["txts"], resident["txt_lengths"],
                                     (resident["img_time"], resident["txt_time"]), resident["missing3"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- value: batch resident in HBM, no per-step host sync ------------------------------------------------------
    if args.cuda_graph:
        # the same cached GraphedStep the user call (get_trainer) replays; here its static input buffers are loaded once
        raw = dict(zip(trainer._RAW_KEYS, (resident["x"], resident["static"], resident["input_lengths"], resident["y"],
                                           resident["img"], resident["txts"], resident["txt_lengths"],
                                           resident["img_time"], resident["txt_time"], resident["missing3"])))
        gs = trainer.graphed_step(args, model, optimizer, criterion, raw)
        gs.load(raw)
        step_dev = lambda i: gs.step(None, i)
    else:
        step_dev = lambda i: trainer.train_step(args, model, optimizer, criterion, prepared, None, i, None)
    for i in range(a.warmup):
        step_dev(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count
    ms_total = timed(step_dev, a.steps)
    launches = _lib.launch_count - n0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / a.steps
    value = world * a.batch / (ms_step * 1e-3)

    # ---- e2e: the user call with pinned host tensors ---------------------------------------------------------------
    e2e = None
    if not a.no_e2e:
        for i in range(max(2, a.warmup // 2)):
            call_trainer(pinned, i)
        ms_e = timed(lambda i: call_trainer(pinned, i), a.steps) / a.steps
        e2e = {"value": world * a.batch / (ms_e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e,
               "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": 4,
               "api": "builder.trainer.get_trainer(..., flow_type='train') -> (model, loss.item())"}

    # ---- roofline of the dominant kernel, timed live on its launching stream ---------------------------------------
    peaks, peak_src = load_peaks()
    roof = dominant_kernel_roofline(a, model, peaks, peak_src)

    if rank != 0:
        _finish(world, model)
        return
    line = {"metric": "train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16 (tcgen05 kind::f16 operands, fp32 accumulate/params/grads)",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "per_gpu_batch": a.batch, "global_batch": a.batch * world,
                       "parallelism": f"dp{world}", "lengths": "ragged+mixed-missing" if a.realistic else "full",
                       "optimizer": a.optimizer, "cuda_graph": bool(args.cuda_graph),
                       "l2": "per-step working set (activations ~5 GB at L=1000) >> 126 MB L2; no explicit flush"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof}
    if world == 1 and not a.no_cpu_baseline:
        v, ms, info = cpu_reference(a, steps=3, warmup=1, budget_s=a.cpu_seconds)
        line["cpu_baseline"] = {"value": v, "unit": "samples/s", **info}
    print(json.dumps(line), flush=True)
    _finish(world, model)


def _finish(world, model):
    """End of a multi-rank run. The NCCL communicator is referenced by the captured CUDA graphs of the train step, and
    `destroy_process_group()` then blocks in ncclCommDestroy (measured on 2 x B200: the JSON line was printed and the
    job never exited). All ranks meet at a barrier, drop the graphs, and leave without running the communicator's
    destructor; the result line is already flushed."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    dist.barrier()
    torch.cuda.synchronize()
    model.__dict__.pop("_graphed_steps", None)
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def profiled_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel_substr`, from the newest committed
    `ncu --set full` summary under profiles/ (tools/summarize_profiles.py; captured at the bench shapes by
    tools/profile_kernels.py). Returns (bytes, file) or (None, None)."""
    import csv
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_full.csv")), reverse=True):
        try:
            with open(path) as f:
                rows = list(csv.reader(f))
            hdr = rows[0]
            ir = next(i for i, h in enumerate(hdr) if h.startswith("dram__bytes_read.sum"))
            iw = next(i for i, h in enumerate(hdr) if h.startswith("dram__bytes_write.sum"))
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            ur = scale.get(hdr[ir].split("[")[-1].rstrip("]"), 1e6)
            uw = scale.get(hdr[iw].split("[")[-1].rstrip("]"), 1e6)
            hits = [r for r in rows[1:] if kernel_substr in r[1]]
            if hits:
                r = hits[-1]
                return float(r[ir]) * ur + float(r[iw]) * uw, os.path.basename(path)
        except Exception:
            continue
    return None, None


def dominant_kernel_roofline(a, model, peaks, peak_src, iters=10):
    """Times the kernel that takes the largest share of the step (attn_bwd_kernel on the vslt stream; share per
    profiles/) alone, on the current stream, with CUDA events; algorithmic FLOPs = 10*Sq*Sk*d per (sample, head)
    (5 GEMMs incl. the S recompute, SURVEY.md 8d ii)."""
    import torch
    fp = model._fused
    st = fp.ws[0]
    B, T, Tl = a.batch, st["T"], st["Tl"]
    kv = fp.ctx["kv_len"][0]
    from medical_tri_modal_pilot_b200 import ops
    l = 0
    run = lambda: ops.attn_bwd(st["qkv"][l], st["O"][l], st["g_h"], kv, B, T, st["lse"][l], st["delta"], st["dq_acc"],
                               st["g_qkv"])
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    lens = kv.float()
    flops = float((10.0 * lens * lens * 64 * 4).sum().item())
    ach = flops / (ms * 1e-3) / 1e12
    peak = peaks["bf16_tflops_sustained"]
    traffic, src = profiled_traffic("attn_bwd_kernel") if (a.tie_len == 1000 and a.batch == 64 and not a.realistic) else (None, None)
    return {"kernel": "attn_bwd_kernel (+delta, dQ convert; vslt stream, one layer)", "bound": "tensor",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "peak_source": peak_src,
            "ms_per_launch": ms, "flops_per_launch": flops, "traffic": traffic,
            "traffic_source": (f"profiles/{src}: dram read+write bytes of attn_bwd_kernel, one launch at these shapes"
                               if src else None)}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
