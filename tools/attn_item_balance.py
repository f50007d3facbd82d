"""CPU model of the attention-backward item schedule on ragged `kv_len` (no GPU needed): how much of the ragged-workload loss
is load imbalance of the STATIC item walk, and what a dynamic (atomic) work queue or an LPT order could still recover.

    python tools/attn_item_balance.py > profiles/r2_attn_bwd_item_balance.txt

The kernel (csrc/attn_bwd_tc05.cu) is persistent: CTA c processes items c, c + grid, ... ; item = (key tile jt, head h,
sample b). An item whose key tile starts past the sample's length is dead (skipped); a live one sweeps the sample's
n_q = ceil(len / 128) query tiles. Cost model: live item = c_item + n_q * c_tile, calibrated on the three uniform shapes of
`profiles/r2f_attn_sweep.json` (S = 1024 / 2048 / 4096, all 2 048 items, 14 rounds on 148 CTAs); the kernel time is the
busiest CTA's sum. The ragged lengths are regenerated with the sweep's own generator (tools/profile_kernels.py --sweep).
Orders modelled:
  index   item % n_jt = key tile, samples in batch order                     (the kernel before commit a8b5025)
  shipped samples longest first + key tile skewed by (sample, head)          (the kernel now)
  queue   the shipped order handed out dynamically (atomic counter): each CTA takes the next item when it is free
  lpt     items sorted by cost, longest first, handed out dynamically       (the best a work queue can do, <= 4/3 optimal)
  bound   max(total work / CTAs, largest item)                              (no schedule can beat it)"""
import heapq
import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMS, H, BT = 148, 4, 128


def sweep_lengths():
    """kv_len of the sweep cases, same generator calls in the same order as tools/profile_kernels.py --sweep (B = 64)."""
    gen = torch.Generator().manual_seed(0)
    out = {}
    for T in (256, 512, 1024, 2048, 4096):
        Bs = max(8, min(64, (64 * 1024) // T))
        rag = torch.randint(T // 4, T + 1, (Bs,), generator=gen)
        rag[0] = T
        half = rag.clone()
        half[1::2] = 0
        out[T] = {"full": [T] * Bs, "ragged": rag.tolist(), "half_missing": half.tolist()}
    return out


def calibrate(meas):
    """c_item, c_tile (µs) from the uniform shapes: 14 rounds x (c_item + n_q * c_tile)."""
    t2, t4 = meas["sweep_attn_bwd_S2048_full"], meas["sweep_attn_bwd_S4096_full"]
    rounds = -(-2048 // SMS)
    c_tile = (t4 - t2) / rounds / 16.0
    c_item = t2 / rounds - 16.0 * c_tile
    return c_item, c_tile


def item_costs(lens, T, order, c_item, c_tile, c_dead=0.03):
    """Costs in walk order. order: 'index' | 'shipped'."""
    n_jt = -(-T // BT)
    B = len(lens)
    if order == "shipped":
        rank = sorted(range(B), key=lambda b: (-lens[b], b))
    else:
        rank = list(range(B))
    costs = []
    for item in range(n_jt * H * B):
        t = item // n_jt
        jt = (item - t * n_jt + t) % n_jt if order == "shipped" else item % n_jt
        b = rank[t // H]
        ln = min(lens[b], T)
        live = jt * BT < ln
        costs.append(c_item + (-(-ln // BT)) * c_tile if live else c_dead)
    return costs


def static_makespan(costs):
    load = [0.0] * SMS
    for i, c in enumerate(costs):
        load[i % SMS] += c
    return max(load)


def dynamic_makespan(costs):
    heap = [0.0] * SMS
    heapq.heapify(heap)
    for c in costs:
        heapq.heappush(heap, heapq.heappop(heap) + c)
    return max(heap)


def main():
    meas = {r["kernel"]: r["ms"] * 1e3 for r in json.load(open(os.path.join(ROOT, "profiles", "r2f_attn_sweep.json")))}
    before = {}
    p = os.path.join(ROOT, "profiles", "r1h_attn_sweep.json")
    if os.path.exists(p):
        before = {r["kernel"]: r["ms"] * 1e3 for r in json.load(open(p))}
    c_item, c_tile = calibrate(meas)
    print(f"cost model from the uniform shapes of r2f_attn_sweep.json: c_item = {c_item:.2f} µs, c_tile = {c_tile:.3f} µs "
          f"(= {c_item * 1965:.0f} / {c_tile * 1965:.0f} cycles at 1 965 MHz); check S=1024 full: model "
          f"{14 * (c_item + 8 * c_tile):.1f} µs, measured {meas['sweep_attn_bwd_S1024_full']:.1f} µs\n")
    print(f"{'case':28s} {'measured':>9s} | {'index':>7s} {'shipped':>8s} {'queue':>7s} {'lpt':>7s} {'bound':>7s} | "
          f"{'shipped/bound':>13s} {'lpt gain':>9s}")
    lens = sweep_lengths()
    for T in (1024, 2048, 4096):
        for pat in ("full", "ragged", "half_missing"):
            L = lens[T][pat]
            idx = item_costs(L, T, "index", c_item, c_tile)
            shp = item_costs(L, T, "shipped", c_item, c_tile)
            m_idx, m_shp = static_makespan(idx), static_makespan(shp)
            m_q = dynamic_makespan(shp)
            m_lpt = dynamic_makespan(sorted(shp, reverse=True))
            bound = max(sum(shp) / SMS, max(shp))
            k = f"sweep_attn_bwd_S{T}_{pat}"
            print(f"S={T:<5d} {pat:20s} {meas[k]:9.1f} | {m_idx:7.1f} {m_shp:8.1f} {m_q:7.1f} {m_lpt:7.1f} {bound:7.1f} | "
                  f"{m_shp / bound:13.2f} {100 * (1 - m_lpt / m_shp):8.1f}%")
    print("\nmeasured = r2f_attn_sweep.json (shipped order, µs). Columns 'index' ... 'bound' are the model's kernel times (µs) "
          "for each order.\n'shipped/bound' = how far the shipped static order is from a perfect balance; 'lpt gain' = what a "
          "cost-sorted dynamic queue would still save according to the model.")


if __name__ == "__main__":
    main()
