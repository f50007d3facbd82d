#!/usr/bin/env python
"""bench.py -- training throughput (samples/s) of `tri_mbt_vsltcls` vslt_img_txt on N x B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # B200-native arm (this repo)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] # the reference's own code on the host CPU cores
    python bench.py --impl reference-gpu-eager [--steps K] [--warmup W]  # the reference's own code on one B200, eager
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
    p.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu-eager"])
    p.add_argument("--batch", type=int, default=64, help="per-GPU batch (--batch-size)")
    p.add_argument("--tie-len", type=int, default=1000)
    p.add_argument("--layers", type=int, default=6)
    p.add_argument("--multiimages", type=int, default=1)
    p.add_argument("--dropout", type=float, default=0.1)
    p.add_argument("--realistic", action="store_true", help="ragged lengths + mixed missing codes instead of full")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-gpu-eager", action="store_true", help="skip the reference-on-B200 (PyTorch eager) comparator")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU budget of the cpu_baseline sample")
    p.add_argument("--optimizer", default="fused", choices=["fused", "torch"])
    p.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured CUDA graph")
    a = p.parse_args()
    a.cpu_seconds_set = any(x.startswith("--cpu-seconds") for x in sys.argv[1:])
    return a


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
# reference arms: the reference's OWN modules (oracle/_ref, see oracle/build_ref.py) stepping through the reference's OWN
# trainer (builder/trainer/trainer.py `missing_trainer`), on the host CPU cores (`--impl reference`, the cpu_baseline) or on
# one B200 in PyTorch eager under its fp16 autocast (`--impl reference-gpu-eager`, the bar the kernels are measured against)
# ---------------------------------------------------------------------------------------------------------------------
class _Sched:
    def step(self, it):
        pass

    def get_lr(self):
        return [1e-4]


class _Logger:
    def log_lr(self, lr, it):
        pass


def config_dict(a, world, graph=True):
    return {"workload": workload_name(a), "per_gpu_batch": a.batch, "global_batch": a.batch * world,
            "parallelism": f"dp{world}", "lengths": "ragged+mixed-missing" if a.realistic else "full",
            "optimizer": a.optimizer, "cuda_graph": bool(graph),
            "l2": "per-step working set (activations ~5 GB at L=1000) >> 126 MB L2; no explicit flush"}


def reference_runner(a, device, Bs):
    """(step_fn, info): one optimisation step of the UNMODIFIED reference (model + trainer from oracle/_ref) on a seeded
    batch of Bs samples of the workload (same L / layers / images / dropout). The reference bakes --batch-size into the
    model (tri_mbt_vsltcls.py:161-165, mbt_encoder.py:675), so the model is built for exactly Bs."""
    import torch
    from medical_tri_modal_pilot_b200 import synth
    from oracle import ref_loader
    on_gpu = torch.device(device).type == "cuda"
    if "control.config" in sys.modules:
        import control.config as C
        args = C.args
        mod = sys.modules["builder.models.8_missing_models.tri_mbt_vsltcls"]
        tr = sys.modules["builder.trainer"]
    else:
        args, mod, tr = ref_loader.load(a.layers, Bs, a.multiimages, a.dropout, device)
        if not on_gpu:
            ref_loader.cpu_trainer_shims()
        torch.autograd.set_detect_anomaly(False)     # the reference leaves it ON (2_train.py:31); off here, stated in DESIGN.md
    args.batch_size = Bs
    args.device = torch.device(device)
    torch.manual_seed(0)
    model = mod.TRI_MBT_VSLTCLS(args).to(device)
    model.train()                                    # 2_train.py:128 (this also leaves Swin's StochasticDepth on, as upstream)
    optimizer = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-6)       # 2_train.py:110
    criterion = torch.nn.BCEWithLogitsLoss()                                            # 2_train.py:76
    n_img = 3 if a.multiimages else 1
    hb = synth.make_batch(Bs, a.tie_len, n_img=n_img, seed=7, full_length=not a.realistic,
                          missing_mode="mixed" if a.realistic else "none", with_pixels=True, feats=False)
    miss = hb["missing"]
    missing3 = torch.stack([torch.zeros_like(miss), (miss >= 2).long(), (miss % 2).long()], 1).float()
    static = torch.stack([hb["gen"], hb["age"]], 1)
    dev = torch.device(device)
    # 2_train.py:143-169: everything but the times / missing rows is moved by the loop; x is cast to HalfTensor there
    x = hb["x"].type(torch.HalfTensor).to(dev) if on_gpu else hb["x"]
    mv = lambda t: t.to(dev)
    sched, logger = _Sched(), _Logger()

    def step(it=0):
        _, loss = tr.get_trainer(args, it, x, mv(static), mv(hb["input_lengths"]).clone(), mv(hb["y"]), None, model, logger,
                                 dev, sched, optimizer, criterion, x_txt=mv(hb["txts"]), x_img=mv(hb["img"]),
                                 txt_lengths=mv(hb["txt_lengths"]).clone(), imgtxt_time=(hb["img_time"], hb["txt_time"]),
                                 scaler=None, missing=missing3, flow_type="train", reports_tokens=None,
                                 reports_lengths=None, criterion_aux=(None, None))
        return loss

    return step


def host_threads():
    threads = os.cpu_count() or 1
    try:
        threads = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return threads


def cpu_reference(a, steps, warmup, budget_s=None):
    """Times the reference's own training step (oracle/_ref: unmodified model + trainer, fp32 -- torch.cuda.amp.autocast
    is inert on CPU tensors --, dropout on, anomaly mode off) on the host cores, on a bounded sample of the workload:
    per-step batch B_s <= --batch, same L / layers / images. Returns (samples_per_s, ms_per_step, info)."""
    import torch
    threads = host_threads()
    torch.set_num_threads(threads)
    n_img = 3 if a.multiimages else 1
    # size the sample: probe with B_s = 2, then pick B_s so that (steps + warmup) steps fit the budget
    probe_step = reference_runner(a, "cpu", 2)
    t0 = time.perf_counter(); probe_step(0); t1 = time.perf_counter(); probe_step(1); t2 = time.perf_counter()
    probe = min(t1 - t0, t2 - t1)
    total = steps + warmup
    if budget_s is None:
        budget_s = 150.0
    per_sample = probe / 2.0
    Bs = int(max(2, min(a.batch, budget_s / max(total, 1) / max(per_sample, 1e-6))))
    step = reference_runner(a, "cpu", Bs)
    for i in range(warmup):
        step(i)
    ts = []
    for i in range(steps):
        t0 = time.perf_counter()
        step(i)
        ts.append(time.perf_counter() - t0)
    dt = sum(ts) / len(ts)
    info = {"cores": threads, "kind": "reference",
            "sample": f"{steps} steps of a B={Bs} slice of the workload batch (same L={a.tie_len}, {a.layers} layers, "
                      f"{n_img} images/sample, dropout {a.dropout}): the reference's own TRI_MBT_VSLTCLS + Swin-T + "
                      f"missing_trainer + AdamW from oracle/_ref, fp32 on CPU, anomaly mode off, {warmup} warm-up"}
    return Bs / dt, dt * 1e3, info


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, ms, info = cpu_reference(a, a.steps, a.warmup, budget_s=a.cpu_seconds if a.cpu_seconds_set else None)
    line = {"impl": "reference", "metric": "train_samples_per_sec", "value": v, "unit": "samples/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(a, max(1, a.gpus), graph=not a.no_graph and a.optimizer == "fused"),
            "ran_on": "host CPU (rank 0 only; the other ranks idle)",
            "cpu_baseline": {"value": v, "unit": "samples/s", **info},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_reference_gpu_eager(a):
    """The unmodified reference modules on ONE B200 in PyTorch eager, stepped by the reference's own trainer under its
    fp16 autocast (trainer.py:126), anomaly mode off: the honest bar for the hand-written kernels."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    step = reference_runner(a, dev, a.batch)
    for i in range(max(1, a.warmup)):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    v = a.batch / (ms * 1e-3)
    line = {"impl": "reference-gpu-eager", "metric": "train_samples_per_sec", "value": v, "unit": "samples/s", "n_gpus": 1,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16 autocast (reference trainer.py:126), fp32 master weights", "data": "synthetic",
            "config": config_dict(a, 1, graph=False),
            "what": "unmodified reference TRI_MBT_VSLTCLS + Swin-T + missing_trainer + AdamW (oracle/_ref) on one B200, "
                    "PyTorch eager (cuBLAS / cuDNN / ATen kernels), device-resident batch, loss.item() per step as upstream",
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    print(json.dumps(line), flush=True)


def sub_bench(a, impl, steps, warmup, extra=(), timeout=900):
    """Run another arm of this script in a child process (the reference's `builder` package cannot share a process with the
    repo's drop-in `builder` shim) and return its JSON line, or {"unavailable": why}."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", impl, "--steps", str(steps), "--warmup", str(warmup),
           "--batch", str(a.batch), "--tie-len", str(a.tie_len), "--layers", str(a.layers), "--multiimages",
           str(a.multiimages), "--dropout", str(a.dropout), *extra]
    if a.realistic:
        cmd.append("--realistic")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    except subprocess.TimeoutExpired:
        return {"unavailable": f"{impl}: timeout after {timeout}s"}
    for ln in reversed(r.stdout.strip().splitlines()):
        if ln.startswith("{"):
            try:
                return json.loads(ln)
            except ValueError:
                pass
    return {"unavailable": f"{impl}: rc={r.returncode} {r.stderr.strip()[-300:]}"}


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
        trainer.ddp_setup_env()      # NCCL may take as many CTAs as GradSync keeps SMs free of compute CTAs
        dist.init_process_group("nccl", device_id=dev, pg_options=trainer.ddp_pg_options())
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
        GradSync(model, overlap=os.environ.get("TMP_B200_DDP_NO_OVERLAP") is None)     # env: A/B of the comm-stream overlap
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
    # the work was real: finite loss and finite fused-path parameters after the timed steps (outside the timed region)
    last_loss = float(step_dev(a.steps).item())
    params_finite = bool(torch.isfinite(model._fused.flat_w).all().item())
    if not (last_loss == last_loss and abs(last_loss) < 1e4 and params_finite):
        raise RuntimeError(f"bench: non-finite training state after the timed steps (loss {last_loss}, finite params {params_finite})")

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
            "config": config_dict(a, world, graph=args.cuda_graph),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
            "loss_after_timed_steps": last_loss}
    if world == 1 and not a.no_gpu_eager:
        # the reference's own modules on this same B200 in PyTorch eager (child process; our graphs / workspaces stay alive,
        # so its memory comes on top: ~15 GB of 180)
        ge = sub_bench(a, "reference-gpu-eager", steps=3, warmup=2)
        line["gpu_eager_baseline"] = ({k: ge.get(k) for k in ("value", "unit", "ms_per_step", "dtype", "what", "peak_mem_gb")}
                                      if "value" in ge else ge)
    if world == 1 and not a.no_cpu_baseline:
        cb = sub_bench(a, "reference", steps=3, warmup=1, extra=("--cpu-seconds", str(a.cpu_seconds)))
        line["cpu_baseline"] = cb.get("cpu_baseline", cb)
    print(json.dumps(line), flush=True)
    _finish(world, model)


def _finish(world, model):
    """End of a multi-rank run: drop the captured step graphs (they reference the NCCL communicator, and
    `destroy_process_group()` blocks in ncclCommDestroy while they live), then tear the group down. A watchdog leaves the
    process if the teardown still does not return -- the result line is already flushed."""
    if world <= 1:
        return
    import threading
    import torch
    import torch.distributed as dist
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    threading.Timer(30.0, lambda: os._exit(0)).start()
    sync = getattr(model, "grad_sync", None)
    if sync is not None:
        sync.close()
    dist.destroy_process_group()
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
    # the call the training step makes: fused protocol (delta and the dQ zeroing come from the LayerNorm backward that
    # produces dO, tmp_layernorm_bwd_attn), ONE kernel launch. Repeated launches keep adding into the same dQ columns --
    # irrelevant for the timing.
    run = lambda: ops.attn_bwd(st["qkv"][l], st["O"][l], st["g_h"], kv, B, T, st["lse"][l], st["delta"], None, st["g_qkv"])
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
    peak = peaks["bf16_tflops"]          # the kernel is timed alone for a few ms: the burst figure (B200_PROFILING.md)
    traffic, src = profiled_traffic("attn_bwd_kernel") if (a.tie_len == 1000 and a.batch == 64 and not a.realistic) else (None, None)
    return {"kernel": "attn_bwd_kernel (vslt stream, one layer; timed alone -> burst peak)", "bound": "tensor",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "peak_source": peak_src,
            "ms_per_launch": ms, "flops_per_launch": flops, "traffic": traffic,
            "traffic_source": (f"profiles/{src}: dram read+write bytes of attn_bwd_kernel, one launch at these shapes"
                               if src else None)}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "reference-gpu-eager":
        run_reference_gpu_eager(a)
    else:
        run_b200(a)
