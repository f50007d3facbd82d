// attn_bwd_tc05.cu -- backward of the modality-aware attention (reference attention.py:24-49 under autograd)
// as a tcgen05/TMEM kernel. One CTA owns a 128-key tile of one (sample, head) and sweeps the live query tiles:
//
//   S^T  = K_j Q_i^T            dP^T = V_j dO_i^T                      (MMAs into TMEM, in two 64-query halves)
//   P^T  = exp2(S^T*c - LSE_i)  dS^T = P^T o (dP^T - delta_i) / 8      (registers -> TMEM / swizzled smem, fp16)
//   dV_j += P^T dO_i            dK_j += dS^T Q_i       dQ_i = dS K_j   (MMAs; dQ leaves through a TMA reduce-add)
//
// Operand placement follows the measured tcgen05.mma costs on B200 (tools/microbench/ub_mma.cu, cta_group::1, K=16,
// M=128, N=64): 75 cycles with both operands in shared memory (operand-read bound; the math is 32), 42 cycles with
// the A operand in TMEM. Four of the five products therefore take A from TMEM:
//   * K_j and V_j (fixed for the CTA's lifetime) are copied once into TMEM as packed fp16 (2 x 32 columns) and are the
//     A operands of S^T and dP^T;
//   * P^T and dS^T are written by the compute warps as packed fp16 INTO the S^T / dP^T accumulator columns they were
//     computed from (the warp that owns S^T columns [16k,16k+16) of a half writes P^T K-step k into columns
//     [16k,16k+8) -- no cross-warp hazard) and are the A operands of dV and dK. The next S^T/dP^T MMAs into that
//     half are issued after dV/dK (tcgen05.mma executes in issue order), so the overwrite is safe;
//   * dQ = dS.K needs dS with queries on the M axis, i.e. the transpose of what the compute warps hold: dS^T also goes
//     to swizzled shared memory and is read MN-major (the one SS product left, 8 x 75 cycles per tile).
// Pipeline: every 128x128 score tile is processed as two 64-query halves with separate TMEM accumulators, so the
// S^T/dP^T MMAs of the next half run while the 16 compute warps do the exponentials of the current one. dS^T in
// shared memory is double-buffered by tile parity and dQ(i) is issued LAST in iteration i (after S^T/dP^T of tile
// i+1), so it never sits between a compute step and the MMAs that step is waiting for; every buffer reuse is ordered
// by the sdp_full barrier the compute warps wait on anyway (tcgen05.commit covers all earlier MMAs of the thread).
// Q_i / dO_i are consumed as B operands straight from their natural [row, d] layout; LSE_i / delta_i ride in the same
// TMA ring stage as Q_i / dO_i (bulk copies). dQ_i (fp32, 128x64) is staged in swizzled smem per warp and added into
// the fp32 accumulator with cp.reduce.async.bulk.tensor (no per-thread atomics). The key-padding mask is kv_len[b]
// applied in-register, only in boundary tiles (P^T rows of masked keys are exactly 0, so dK/dV of pad rows are exactly
// 0, as in the reference).
#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int kComputeWarps = 16;  // 4 per TMEM lane quarter: each owns 16 of the 64 query columns of a half-step
constexpr int kFlushWarps = 8;     // compute warps 0..7 also move dQ / dK / dV out of TMEM (32 columns each)
constexpr int kThreads = 64 + 32 * kComputeWarps;
constexpr int BT = 128;  // tile rows (keys per CTA, queries per iteration)
constexpr int HD = 64;
constexpr int kStages = 3;   // Q_i / dO_i ring: a stage is released when dV/dK of tile i retire; 2 stages left every tile
                             // waiting a full TMA round trip
constexpr int kTile = BT * HD * 2;   // 16 KB : [128 rows x 64 fp16]
constexpr int kStatBytes = 2 * BT * 4;  // lse[128] | delta[128] fp32 per ring stage

constexpr int kHalf = BT * 128;                           // 16 KB: one [128 keys x 64 queries] fp16 sub-tile
constexpr int kSmemK = 0;                                 // K_j: B operand of dQ for the whole sweep
constexpr int kSmemV = kSmemK + kTile;                    // V_j: only until it has been copied to TMEM, then ...
// dS^T (shared-memory copy, read MN-major by dQ): two [half 0 | half 1] pairs, alternating with the query tile's
// parity. Pair 1 starts in the V_j buffer (dead after the copy to TMEM) and runs 16 KB past it.
constexpr int kSmemDST1 = kSmemV;                         // pair of odd tiles  : 32 KB
constexpr int kSmemDST0 = kSmemDST1 + 2 * kHalf;          // pair of even tiles : 32 KB
constexpr int kSmemQ = kSmemDST0 + 2 * kHalf;
constexpr int kSmemDO = kSmemQ + kStages * kTile;
constexpr int kSmemDQ = kSmemDO + kStages * kTile;        // per flush warp: [32 rows x 32 fp32], 128B swizzle
constexpr int kSmemStat = kSmemDQ + kFlushWarps * 4096;
constexpr int kSmemBar = kSmemStat + kStages * kStatBytes;
constexpr int kMaxSortB = 256;                               // samples ordered by length inside the kernel up to this B
constexpr int kSmemOrder = kSmemBar + 256;                   // uint16 order[kMaxSortB]
constexpr int kSmemTotal = kSmemOrder + 2 * kMaxSortB + 1024;
static_assert(kSmemTotal <= 232448, "attention backward: shared memory");

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Bars {
  uint64_t kv_full;
  uint64_t kv_tmem;      // compute -> MMA : K_j / V_j copied to TMEM (4 warps, one per lane quarter)
  uint64_t kv_free;      // MMA -> MMA thread : every MMA of the item retired, K_j / V_j shared buffers reusable
  uint64_t qdo_full[kStages], qdo_empty[kStages];
  uint64_t sdp_full[2];  // MMA -> compute : S^T and dP^T of half h in TMEM
  uint64_t pds_full[2];  // compute -> MMA : P^T and dS^T of half h in smem (and the TMEM half is drained)
  uint64_t dq_full;      // MMA -> compute : dQ tile in TMEM, all MMAs of this iteration retired
  uint64_t dq_empty;     // compute -> MMA : dQ TMEM drained
  uint32_t tmem_slot;
};

// Work item = (key tile jt, head h, sample b), jt fastest. The kernel is PERSISTENT: grid = min(items, SMs) and every CTA
// walks items blockIdx.x, += gridDim.x. One CTA per item cost ~17 k cycles of launch gap, barrier/TMEM set-up, first
// TMA round trip and pipeline fill/drain against ~2.7 k cycles per 128x128 tile (T=1005 vs T=2005 timings: 44 % of the
// kernel at 8 tiles per item); inside the loop only the K_j/V_j round trip and the fill/drain remain: barriers and TMEM
// are set up once, and the Q/dO ring (its own producer warp) runs ahead into the next item while this one drains.
// All roles enumerate the same item sequence; ring stages and barrier phases follow running counters (gq = query tiles
// processed so far, it = live items so far).
// DQ16: dQ tiles are reduce-added in fp16 straight into the dQ columns of dQKV (pre-zeroed by the caller, see
// tmp_layernorm_bwd_attn) instead of into an fp32 accumulator that needs a memset before and a convert pass after.
template <bool DQ16>
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDKV,
                const int32_t* __restrict__ kv_len, int T, int n_jt, int H, int q_tiles_max,
                int n_items, const float* __restrict__ lse2, const float* __restrict__ delta, int T_lse,
                uint16_t* __restrict__ dQKV, float scale_log2, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  Bars* bars = (Bars*)(smem + kSmemBar);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQKV);
    prefetch_tmap(&tmDO);
    prefetch_tmap(&tmDQ);
    prefetch_tmap(&tmDKV);
    mbar_init(&bars->kv_full, 1);
    mbar_init(&bars->kv_tmem, 4);
    mbar_init(&bars->kv_free, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars->qdo_full[s], 1);
      mbar_init(&bars->qdo_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->sdp_full[s], 1);
      mbar_init(&bars->pds_full[s], kComputeWarps);
    }
    mbar_init(&bars->dq_full, 1);
    mbar_init(&bars->dq_empty, kFlushWarps);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  const uint32_t tm_ST = tmem_base + 0;     // [2] x 64 columns
  const uint32_t tm_DPT = tmem_base + 128;  // [2] x 64 columns
  const uint32_t tm_DV = tmem_base + 256;
  const uint32_t tm_DK = tmem_base + 320;
  const uint32_t tm_DQ = tmem_base + 384;
  const uint32_t tm_K = tmem_base + 448;    // K_j as packed fp16: 128 keys (lanes) x 64 d = 32 columns
  const uint32_t tm_V = tmem_base + 480;    // V_j likewise
  // P^T / dS^T (packed fp16) alias the S^T / dP^T columns: K-step k (16 queries) of half hh at column hh*64 + 16*k

  // Item order for ragged kv_len (static stride over items, so the ORDER decides the balance):
  //  * samples are taken longest first (rank computed here, once per CTA, B <= kMaxSortB): a CTA's items i, i + grid, ...
  //    then sample every cost level (stratified) instead of a random subset of samples;
  //  * the key tile of an item is skewed by its (sample, head) index: with jt = item % n_jt and a stride of 148 = 4 mod 8 a
  //    CTA only ever saw two key-tile indices -- the CTAs stuck with jt = 3 / 7 got the tiles that are dead for short
  //    samples, the others all the live ones. The n_jt items of one (sample, head) stay adjacent (they share Q / dO in L2).
  uint16_t* order = reinterpret_cast<uint16_t*>(smem + kSmemOrder);
  const int n_b = n_items / (n_jt * H);
  const bool sorted = kv_len != nullptr && n_b <= kMaxSortB;
  if (sorted) {
    for (int b = threadIdx.x; b < n_b; b += blockDim.x) {
      const int lb = __ldg(kv_len + b);
      int rank = 0;
      for (int o = 0; o < n_b; ++o) {
        const int lo = __ldg(kv_len + o);
        rank += (lo > lb) || (lo == lb && o < b);
      }
      order[rank] = (uint16_t)b;
    }
    __syncthreads();
  }
  // item -> (jt, h, b); live length of the sample
  struct Item { int jt, h, b, k0, len, n_q, row_base; };
  auto decode = [&](int item) {
    Item w;
    const int t = item / n_jt;
    w.jt = (item - t * n_jt + t) % n_jt;
    w.h = t % H;
    w.b = sorted ? (int)order[t / H] : t / H;
    w.k0 = w.jt * BT;
    w.len = kv_len ? min(__ldg(kv_len + w.b), T) : T;
    w.n_q = min((w.len + BT - 1) / BT, q_tiles_max);   // live query tiles (rows >= len are padding: dO == 0; rows past
                                                       // q_tiles_max tiles carry no gradient by the caller's contract)
    w.row_base = w.b * T;
    return w;
  };

  if (warp == 0) {
    // ===================== TMA producer: the Q_i / dO_i / lse_i / delta_i ring =====================
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item w = decode(item);
        if (w.k0 >= w.len) continue;
        const size_t stat_base = ((size_t)w.b * H + w.h) * T_lse;
        for (int i = 0; i < w.n_q; ++i) {
          mbar_wait(&bars->qdo_empty[st], ph ^ 1);
          mbar_expect_tx(&bars->qdo_full[st], 2 * kTile + kStatBytes);
          tma_load_2d(smem + kSmemQ + st * kTile, &tmQKV, &bars->qdo_full[st], w.h * HD, w.row_base + i * BT);
          tma_load_2d(smem + kSmemDO + st * kTile, &tmDO, &bars->qdo_full[st], w.h * HD, w.row_base + i * BT);
          bulk_load_1d(smem + kSmemStat + st * kStatBytes, lse2 + stat_base + i * BT, BT * 4, &bars->qdo_full[st]);
          bulk_load_1d(smem + kSmemStat + st * kStatBytes + BT * 4, delta + stat_base + i * BT, BT * 4,
                       &bars->qdo_full[st]);
          if (++st == kStages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (also loads K_j / V_j: it is the one that knows when they are dead) =====================
    if (elect_one()) {
      // all operands fp16 (kind::f16 needs A and B in the same format; gradients are fp16 with a host-side scale)
      constexpr uint32_t idesc_s = make_idesc(BT, 64, 0, 0, FMT_F16, FMT_F16);   // S^T_h / dP^T_h : [128 keys x 64 q]
      constexpr uint32_t idesc_kv = make_idesc(BT, HD, 0, 1, FMT_F16, FMT_F16);  // dV += P^T dO, dK += dS^T Q (B MN-major)
      constexpr uint32_t idesc_dq = make_idesc(BT, HD, 1, 1, FMT_F16, FMT_F16);  // dQ   = dS K     (A, B MN-major)
      const uint32_t sK = smem_u32(smem + kSmemK);
      auto issue_sdp = [&](uint32_t g, int hh) {   // S^T_hh = K_j Q[hh]^T, dP^T_hh = V_j dO[hh]^T   (A in TMEM)
        const int st = g % kStages;
        const uint32_t sQ = smem_u32(smem + kSmemQ + st * kTile) + hh * 8192;    // 64 query rows = 8192 B
        const uint32_t sDO = smem_u32(smem + kSmemDO + st * kTile) + hh * 8192;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_ts(tm_ST + hh * 64, tm_K + k * 8, make_sdesc_sw128(sQ + k * 32, 16, 1024), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_ts(tm_DPT + hh * 64, tm_V + k * 8, make_sdesc_sw128(sDO + k * 32, 16, 1024), idesc_s, k != 0);
        umma_commit(&bars->sdp_full[hh]);
      };
      auto issue_dvdk = [&](uint32_t g, int i, int hh) {  // dV += P^T_hh dO[hh], dK += dS^T_hh Q[hh]  (reduction over 64 queries)
        const int st = g % kStages;
        const uint32_t sQ = smem_u32(smem + kSmemQ + st * kTile);
        const uint32_t sDO = smem_u32(smem + kSmemDO + st * kTile);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // A = P^T_hh K-step k: 8 packed columns at the head of S^T columns [16k, 16k+16)
          umma_ts(tm_DV, tm_ST + hh * 64 + k * 16, make_sdesc_sw128(sDO + (hh * 4 + k) * 2048, BT * 128, 1024), idesc_kv,
                  (i | hh | k) != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // A = dS^T_hh K-step k, same placement inside the dP^T columns
          umma_ts(tm_DK, tm_DPT + hh * 64 + k * 16, make_sdesc_sw128(sQ + (hh * 4 + k) * 2048, BT * 128, 1024), idesc_kv,
                  (i | hh | k) != 0);
      };
      uint32_t gq = 0;
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item w = decode(item);
        if (w.k0 >= w.len) continue;
        // K_j / V_j of this item. Their buffers (V's doubles as dS^T pair 1) are dead once the last dQ of the previous
        // item has retired.
        if (it > 0) mbar_wait(&bars->kv_free, (it - 1) & 1);
        if (!(dbg & 1) || it == 0) {   // dbg bit 0 (timing experiment only): keep the first item's K_j / V_j
          mbar_expect_tx(&bars->kv_full, 2 * kTile);
          tma_load_2d(smem + kSmemK, &tmQKV, &bars->kv_full, 256 + w.h * HD, w.row_base + w.k0);
          tma_load_2d(smem + kSmemV, &tmQKV, &bars->kv_full, 512 + w.h * HD, w.row_base + w.k0);
        }
        mbar_wait(&bars->kv_tmem, it & 1);
        mbar_wait(&bars->qdo_full[gq % kStages], (gq / kStages) & 1);
        tc_fence_after();
        issue_sdp(gq, 0);
        issue_sdp(gq, 1);
        for (int i = 0; i < w.n_q; ++i) {
          const uint32_t g = gq + i;
          const uint32_t ph = g & 1;
          const bool more = i + 1 < w.n_q;
          // ---- half 0 ----
          mbar_wait(&bars->pds_full[0], ph);
          tc_fence_after();
          issue_dvdk(g, i, 0);
          if (more) {
            mbar_wait(&bars->qdo_full[(g + 1) % kStages], ((g + 1) / kStages) & 1);
            tc_fence_after();
            issue_sdp(g + 1, 0);
          }
          // ---- half 1 ----
          mbar_wait(&bars->pds_full[1], ph);
          tc_fence_after();
          issue_dvdk(g, i, 1);
          if (more) issue_sdp(g + 1, 1);
          if (g > 0) {
            mbar_wait(&bars->dq_empty, (g - 1) & 1);
            tc_fence_after();
          }
          // dQ = dS K_j  (reduction over the 128 keys; A = the shared-memory copy of dS^T read MN-major: the two
          // 64-query halves are kHalf apart)
          const uint32_t sDS = smem_u32(smem + ((g & 1) ? kSmemDST1 : kSmemDST0));
#pragma unroll
          for (int k = 0; k < BT / 16; ++k)
            umma_ss(tm_DQ, make_sdesc_sw128(sDS + k * 2048, kHalf, 1024),
                    make_sdesc_sw128(sK + k * 2048, BT * 128, 1024), idesc_dq, k != 0);
          umma_commit(&bars->dq_full);
          umma_commit(&bars->qdo_empty[g % kStages]);
        }
        umma_commit(&bars->kv_free);
        gq += w.n_q;
        ++it;
      }
    }
  } else {
    // ===================== compute warps =====================
    const int cw = warp - 2;
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int colq = cw >> 2;           // which 16 of the 64 query columns of a half-step (S^T / dP^T)
    const int colhalf = (cw >> 2) & 1;  // flush warps (cw < 8): which 32 of the 64 dQ / dK / dV columns
    const bool flusher = cw < kFlushWarps;
    const int r = quarter * 32 + lane;  // TMEM lane: key row for S^T/dP^T/dK/dV, query row for dQ
    uint8_t* sDQ = smem + kSmemDQ + (cw & (kFlushWarps - 1)) * 4096;
    uint32_t gq = 0;
    int it = 0;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item w = decode(item);
      const int h = w.h, k0 = w.k0, len = w.len, n_q = w.n_q, row_base = w.row_base;
      if (k0 >= len) {
        // masked / padding keys: dK = dV = 0
        if (colq == 0 && k0 + r < T) {
          uint4* dk = reinterpret_cast<uint4*>(dQKV + (size_t)(row_base + k0 + r) * 768 + 256 + h * HD);
          uint4* dv = reinterpret_cast<uint4*>(dQKV + (size_t)(row_base + k0 + r) * 768 + 512 + h * HD);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            dk[i] = make_uint4(0, 0, 0, 0);
            dv[i] = make_uint4(0, 0, 0, 0);
          }
        }
        continue;
      }
      const bool key_ok = (k0 + r) < len;
      const bool key_tile_partial = (k0 + BT) > len;

      if (colq == 0) {
        // K_j, V_j -> TMEM (A operands of S^T / dP^T): thread = key row = TMEM lane, 64 fp16 = 32 packed columns each.
        // One warp per lane quarter; the swizzled 16 B chunks of a row are read in logical order. (tm_K / tm_V are
        // free: this warp has consumed the last S^T/dP^T of the previous item.)
        if (!(dbg & 1) || it == 0) mbar_wait(&bars->kv_full, it & 1);
#pragma unroll 1
        for (int which = 0; which < 2 && (!(dbg & 1) || it == 0); ++which) {
          const uint8_t* src = smem + (which == 0 ? kSmemK : kSmemV);
          uint32_t wv[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 u = *reinterpret_cast<const uint4*>(src + sw128_offset(r, c));
            wv[c * 4] = u.x; wv[c * 4 + 1] = u.y; wv[c * 4 + 2] = u.z; wv[c * 4 + 3] = u.w;
          }
          tmem_st32(tmem_addr(which == 0 ? tm_K : tm_V, quarter * 32, 0), wv);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->kv_tmem);
      }

      auto dq_flush = [&](int i) {   // dQ(i): TMEM -> swizzled smem box -> TMA reduce-add into the fp32 accumulator
        mbar_wait(&bars->dq_full, (gq + i) & 1);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem_addr(tm_DQ, quarter * 32, colhalf * 32), v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->dq_empty);
        if (dbg & 2) return;          // dbg bit 1 (timing experiment only): no write-out
        tma_store_wait_read0();   // previous reduce / store of this warp has finished reading the staging box (every lane
                                  // waits: bulk groups are per thread, lanes without any return at once)
        __syncwarp();
        if (DQ16) {
          // fp16 box [32 rows x 32 cols], 64B swizzle (dense 64 B rows): chunk q of row `lane` at lane*64 + ((q ^ ((lane>>1)&3))<<4)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(sDQ + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) =
                make_uint4(pack_f16x2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1])),
                           pack_f16x2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3])),
                           pack_f16x2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5])),
                           pack_f16x2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7])));
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(sDQ + sw128_offset(lane, q)) = make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(&tmDQ, sDQ, h * HD + colhalf * 32, row_base + i * BT + quarter * 32);
          tma_store_commit();
        }
      };

      for (int i = 0; i < n_q; ++i) {
        const uint32_t g = gq + i;
        const int st = g % kStages;
        const float* st_lse = (const float*)(smem + kSmemStat + st * kStatBytes);
        const float* st_dl = st_lse + BT;
        const bool need_mask = key_tile_partial || (i * BT + BT > len);
        uint8_t* sDST = smem + ((g & 1) ? kSmemDST1 : kSmemDST0);
        // the ring stage (Q_i, dO_i, lse_i, delta_i) is complete before the MMA warp could issue S^T(i); observing
        // it here orders our generic-proxy reads of lse/delta after the bulk copies.
        mbar_wait(&bars->qdo_full[st], (g / kStages) & 1);
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          mbar_wait(&bars->sdp_full[hh], g & 1);
          tc_fence_after();
          uint32_t s[16], dp[16];
          tmem_ld16(tmem_addr(tm_ST + hh * 64, quarter * 32, colq * 16), s);
          tmem_ld16(tmem_addr(tm_DPT + hh * 64, quarter * 32, colq * 16), dp);
          tmem_ld_wait();
          const int qc0 = hh * 64 + colq * 16;   // first query column (within the 128-query tile) of this thread's run
          // P^T = p / 8 and dS^T = (p / 8) (dP^T - delta): the exponent carries lse + 3 (log2 units), so the 1/sqrt(d) of
          // dS costs nothing, and the factor 8 missing from P^T is applied to dV once, when the CTA writes it out (exact
          // powers of two on both sides). All fp32 math runs as packed pairs (FFMA2 / FMUL2); masked (key, query) pairs
          // get a score of -inf, which makes both p and dS exactly 0.
          uint32_t pp[8], dd[8];
          const f32x2_t c2 = f2_pack(scale_log2, scale_log2), m1 = f2_pack(-1.f, -1.f), m3 = f2_pack(-3.f, -3.f);
          const int n_ok = key_ok ? (len - i * BT - qc0) : 0;   // valid query columns in this thread's run of 16
#pragma unroll
          for (int t = 0; t < 16; t += 4) {
            const float4 l4 = *reinterpret_cast<const float4*>(st_lse + qc0 + t);
            const float4 d4 = *reinterpret_cast<const float4*>(st_dl + qc0 + t);
            if (need_mask) {
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (t + u >= n_ok) s[t + u] = 0xff800000u;   // -inf
            }
#pragma unroll
            for (int u = 0; u < 4; u += 2) {
              const f32x2_t nl = f2_fma(u == 0 ? f2_pack(l4.x, l4.y) : f2_pack(l4.z, l4.w), m1, m3);   // -(lse + 3)
              const f32x2_t x = f2_fma(f2_pack(__uint_as_float(s[t + u]), __uint_as_float(s[t + u + 1])), c2, nl);
              float x0, x1;
              f2_unpack(x, x0, x1);
              const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
              const f32x2_t gd = f2_fma(u == 0 ? f2_pack(d4.x, d4.y) : f2_pack(d4.z, d4.w), m1,
                                        f2_pack(__uint_as_float(dp[t + u]), __uint_as_float(dp[t + u + 1])));   // dP^T - delta
              float d0, d1;
              f2_unpack(f2_mul(f2_pack(p0, p1), gd), d0, d1);
              pp[(t + u) >> 1] = pack_f16x2(p0, p1);
              dd[(t + u) >> 1] = pack_f16x2(d0, d1);
            }
          }
          // The S^T / dP^T columns just read become P^T / dS^T (packed fp16, A operands of dV / dK): this warp owns columns
          // [16 colq, 16 colq + 16) of both accumulators and writes the 8 packed columns of K-step colq at their head.
          // The shared-memory copy of dS^T (for dQ) goes to the pair of this tile's parity: its last reader, the dQ two
          // tiles back, was issued before S^T/dP^T of this half-step, whose commit (sdp_full, observed above) covers it.
          const uint32_t sub = hh * kHalf;
#pragma unroll
          for (int q4 = 0; q4 < 2; ++q4) {
            const uint32_t off = sub + sw128_offset(r, colq * 2 + q4);
            *reinterpret_cast<uint4*>(sDST + off) = make_uint4(dd[q4 * 4], dd[q4 * 4 + 1], dd[q4 * 4 + 2], dd[q4 * 4 + 3]);
          }
          tmem_st8(tmem_addr(tm_ST + hh * 64, quarter * 32, colq * 16), pp);
          tmem_st8(tmem_addr(tm_DPT + hh * 64, quarter * 32, colq * 16), dd);
          // the shared-memory copy is only read by dQ(i), issued after BOTH halves have arrived: one generic->async proxy
          // fence per tile (in half 1, covering this thread's stores of both halves) instead of one per half-step
          if (hh == 1) fence_proxy_async_smem();
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->pds_full[hh]);
          // dQ(i-1) is the last MMA of iteration i-1 (issued after S^T/dP^T of this tile): it has retired by now
          if (hh == 1 && i > 0 && flusher) dq_flush(i - 1);
        }
      }
      if (flusher) {
        dq_flush(n_q - 1);
        // dK_j, dV_j (all MMAs of the item retired: last dq_full). Flush warp (quarter, colhalf) writes all 64 columns of
        // 32 key rows of ONE of the two (colhalf 0: dK, 1: dV): a [32 x 128 B] box, staged in the warp's swizzled dQ
        // box and written with one TMA store. (Per-thread 16 B stores to 32 different rows per instruction -- 32
        // partial sectors each -- cost 1.8 us per item, 25 us of the 267 us kernel at T=1005.)
        if (dbg & 2) { tc_fence_before(); gq += n_q; ++it; continue; }
        const int which = colhalf;
        const uint32_t src = which == 0 ? tm_DK : tm_DV;
        const float osc = which == 0 ? 1.f : 8.f;   // dV was accumulated from P^T / 8
        const int kr = k0 + r;
        const bool full_tile = (k0 + BT) <= T;      // the 2-D map cannot clip at the sample boundary: last tile by hand
        if (full_tile) {
          tma_store_wait_read0();  // last dQ reduce has finished reading the staging box
          __syncwarp();
        }
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_addr(src, quarter * 32, c * 32), v);
          tmem_ld_wait();
          uint4 u[4];
#pragma unroll
          for (int t = 0; t < 4; ++t)
            u[t] = make_uint4(pack_f16x2(__uint_as_float(v[t * 8 + 0]) * osc, __uint_as_float(v[t * 8 + 1]) * osc),
                              pack_f16x2(__uint_as_float(v[t * 8 + 2]) * osc, __uint_as_float(v[t * 8 + 3]) * osc),
                              pack_f16x2(__uint_as_float(v[t * 8 + 4]) * osc, __uint_as_float(v[t * 8 + 5]) * osc),
                              pack_f16x2(__uint_as_float(v[t * 8 + 6]) * osc, __uint_as_float(v[t * 8 + 7]) * osc));
          if (full_tile) {
#pragma unroll
            for (int t = 0; t < 4; ++t) *reinterpret_cast<uint4*>(sDQ + sw128_offset(lane, c * 4 + t)) = u[t];
          } else if (kr < T) {
            uint4* dst = reinterpret_cast<uint4*>(dQKV + (size_t)(row_base + kr) * 768 + (which == 0 ? 256 : 512) + h * HD +
                                                  c * 32);
#pragma unroll
            for (int t = 0; t < 4; ++t) dst[t] = u[t];
          }
        }
        if (full_tile) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmDKV, sDQ, (which == 0 ? 256 : 512) + h * HD, row_base + k0 + quarter * 32);
            tma_store_commit();
          }
        }
        // the accumulators are read: the next item's first dV/dK MMAs (accumulate = 0) are issued only after this warp
        // has arrived on pds_full of that item's first tile, i.e. after this point in program order
        tc_fence_before();
      }
      gq += n_q;
      ++it;
    }
    if (flusher) tma_store_wait_read0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// Single-query attention backward. In the last fused layer under --mbt-only-vslt 1 only the CLS row of the vslt stream has a
// consumer, so dO is zero everywhere but in ONE query row per sample. The tile kernel above spends 75 us on that (2 048
// items of one query tile each: all item overhead) plus three memsets; the job is ~165 MB of traffic:
//   p_k = exp2((q . k_k) * scale_log2 - lse2[b,h,q_row])   for k < len
//   dV_k = p_k dO        dS_k = p_k (dO . v_k - delta)      dK_k = dS_k q / 8      dQ[q_row] = sum_k dS_k k_k / 8
// every other dQ row and the dK / dV rows of masked keys are zero. One CTA per (sample, head), one thread per key (stride
// 128), fp32 math on the CUDA cores, no atomics. dO_row / O_row: [B, 256] fp16 (the CLS rows only), gradients scaled like
// everywhere else on the 16-bit path.
// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_bwd_single_query_kernel(const uint16_t* __restrict__ qkv,
                                                                    const uint16_t* __restrict__ dO_row,
                                                                    const uint16_t* __restrict__ O_row,
                                                                    const int32_t* __restrict__ kv_len, int T, int H,
                                                                    int q_row, const float* __restrict__ lse2, int T_lse,
                                                                    uint16_t* __restrict__ dQKV, float scale_log2) {
  __shared__ float sq[HD], sdo[HD], sprod[HD];
  __shared__ float sred[4][HD];
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int len = kv_len ? min(__ldg(kv_len + b), T) : T;
  const size_t row_base = (size_t)b * T;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const bool live = q_row < len;
  if (t < HD) {
    const float qv = __half2float(__ushort_as_half(qkv[(row_base + q_row) * 768 + h * HD + t]));
    const float dv = __half2float(__ushort_as_half(dO_row[(size_t)b * 256 + h * HD + t]));
    const float ov = __half2float(__ushort_as_half(O_row[(size_t)b * 256 + h * HD + t]));
    sq[t] = qv;
    sdo[t] = dv;
    sprod[t] = dv * ov;
  }
  __syncthreads();
  float delta = 0.f;
#pragma unroll
  for (int d = 0; d < HD; ++d) delta += sprod[d];     // 64 broadcast reads: same order in every thread
  const float lse = lse2[((size_t)b * H + h) * T_lse + q_row];
  float dq[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) dq[d] = 0.f;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (int k = t; k < T; k += 128) {
    uint16_t* out_row = dQKV + (row_base + k) * 768 + h * HD;
    uint4* dq_out = reinterpret_cast<uint4*>(out_row);
    uint4* dk_out = reinterpret_cast<uint4*>(out_row + 256);
    uint4* dv_out = reinterpret_cast<uint4*>(out_row + 512);
    if (k != q_row) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dq_out[i] = zero4;
    }
    if (!live || k >= len) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { dk_out[i] = zero4; dv_out[i] = zero4; }
      continue;
    }
    const uint4* kp = reinterpret_cast<const uint4*>(qkv + (row_base + k) * 768 + 256 + h * HD);
    const uint4* vp = reinterpret_cast<const uint4*>(qkv + (row_base + k) * 768 + 512 + h * HD);
    uint4 kr[8], vr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { kr[i] = __ldg(kp + i); vr[i] = __ldg(vp + i); }
    float s = 0.f, dp = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t kw[4] = {kr[i].x, kr[i].y, kr[i].z, kr[i].w};
      const uint32_t vw[4] = {vr[i].x, vr[i].y, vr[i].z, vr[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 kf = unpack2<FMT_F16>(kw[j]), vf = unpack2<FMT_F16>(vw[j]);
        const int d = i * 8 + j * 2;
        s = fmaf(sq[d], kf.x, s); s = fmaf(sq[d + 1], kf.y, s);
        dp = fmaf(sdo[d], vf.x, dp); dp = fmaf(sdo[d + 1], vf.y, dp);
      }
    }
    const float p = ex2_approx(fmaf(s, scale_log2, -lse));
    const float ds = p * (dp - delta) * 0.125f;      // d/d(q.k) incl. the 1/sqrt(64) of the score
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t kw[4] = {kr[i].x, kr[i].y, kr[i].z, kr[i].w};
      uint32_t ok[4], ov[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 kf = unpack2<FMT_F16>(kw[j]);
        const int d = i * 8 + j * 2;
        dq[d] = fmaf(ds, kf.x, dq[d]);
        dq[d + 1] = fmaf(ds, kf.y, dq[d + 1]);
        ok[j] = pack_f16x2(ds * sq[d], ds * sq[d + 1]);
        ov[j] = pack_f16x2(p * sdo[d], p * sdo[d + 1]);
      }
      dk_out[i] = make_uint4(ok[0], ok[1], ok[2], ok[3]);
      dv_out[i] = make_uint4(ov[0], ov[1], ov[2], ov[3]);
    }
  }
  // dQ[q_row] = sum over the keys of all threads
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    float v = dq[d];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sred[warp][d] = v;
  }
  __syncthreads();
  if (t < HD) {
    const float v = live ? (sred[0][t] + sred[1][t]) + (sred[2][t] + sred[3][t]) : 0.f;
    dQKV[(row_base + q_row) * 768 + h * HD + t] = __half_as_ushort(__float2half_rn(v));
  }
}

// delta[b,h,q] = sum_d dO[b,q,h*64+d] * O[b,q,h*64+d]   (one warp per row; rows past T_lse padding are zeroed)
__global__ void attn_bwd_delta_kernel(const uint16_t* __restrict__ O, const uint16_t* __restrict__ dO, int ld, int B, int T,
                                      int H, float* __restrict__ delta, int T_lse) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B * T_lse) return;
  const int b = row / T_lse, q = row % T_lse;
  float acc = 0.f;
  if (q < T) {
    const uint4 o = *reinterpret_cast<const uint4*>(O + (size_t)(b * T + q) * ld + lane * 8);
    const uint4 g = *reinterpret_cast<const uint4*>(dO + (size_t)(b * T + q) * ld + lane * 8);
    const uint32_t ow[4] = {o.x, o.y, o.z, o.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int t = 0; t < 4; ++t)
    {
      const float2 of = unpack2<FMT_F16>(ow[t]);
      const float2 gf = unpack2<FMT_F16>(gw[t]);
      acc += of.x * gf.x + of.y * gf.y;
    }
  }
  // 8 lanes per head (8 lanes x 8 columns = 64)
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if ((lane & 7) == 0) delta[((size_t)b * H + (lane >> 3)) * T_lse + q] = acc;
}

// dQKV[:, 0:256] = fp16(dQ_acc)
__global__ void attn_bwd_dq_convert_kernel(const float* __restrict__ dQ_acc, uint16_t* __restrict__ dQKV, size_t rows) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread = 8 columns
  if (idx >= rows * 32) return;
  const size_t row = idx >> 5;
  const int c8 = (int)(idx & 31) * 8;
  const float4 a = *reinterpret_cast<const float4*>(dQ_acc + row * 256 + c8);
  const float4 b2 = *reinterpret_cast<const float4*>(dQ_acc + row * 256 + c8 + 4);
  *reinterpret_cast<uint4*>(dQKV + row * 768 + c8) =
      make_uint4(pack_f16x2(a.x, a.y), pack_f16x2(a.z, a.w), pack_f16x2(b2.x, b2.y), pack_f16x2(b2.z, b2.w));
}

}  // namespace

// qkv [B*T,768] and O [B*T,ld] fp16 (forward quantities), dO [B*T,ld] fp16 (scaled gradient); lse2 from the forward; delta [B,H,T_lse] and dQ_acc [B*T,256] fp32 are
// workspaces (dQ_acc is zeroed here); dQKV [B*T,768] fp16 receives dQ|dK|dV.
// Two protocols:
//  * dQ_acc != NULL (stand-alone): delta is computed here, dQ is accumulated in the fp32 workspace dQ_acc (zeroed here)
//    and converted to fp16 at the end -- three extra passes over [B*T, 256] tensors;
//  * dQ_acc == NULL (fused, what the training step uses): the caller has ALREADY written delta and zeroed the dQ columns
//    of dQKV (tmp_layernorm_bwd_attn does both while it produces dO); dQ tiles are reduce-added in fp16 in place.
// q_rows: only the leading q_rows query rows of every sample carry a gradient (rounded up to whole tiles; T = all): the
// query sweep of every key tile stops there (the last layer of `--mbt-only-vslt 1`: only the CLS row has dO != 0).
extern "C" int tmp_mma_attn_bwd(const void* qkv, const void* O, const void* dO, int ld, const int32_t* kv_len, int B,
                                int T, int H, const float* lse2, int T_lse, float* delta, float* dQ_acc, void* dQKV,
                                int q_rows, void* stream) {
  TMP_REQUIRE(q_rows > 0 && q_rows <= T, "attn_bwd: q_rows must be in [1, T]");
  const int q_tiles_max = (q_rows + BT - 1) / BT;
  TMP_REQUIRE(qkv && O && dO && lse2 && delta && dQKV, "attn_bwd: null operand");
  TMP_REQUIRE(B > 0 && T > 0 && H == 4 && ld == 256, "attn_bwd: need H==4, ld==256 (B=%d T=%d H=%d ld=%d)", B, T, H, ld);
  TMP_REQUIRE(T_lse % BT == 0 && T_lse >= T, "attn_bwd: T_lse must be a multiple of 128 and >= T");
  cudaStream_t st = (cudaStream_t)stream;
  const bool fused = dQ_acc == nullptr;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (e != cudaSuccess) {
      tmp::set_error("cudaFuncSetAttribute(attn_bwd): %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  CUtensorMap tmQKV, tmDO, tmDQ, tmDKV;
  int rc = tmp::encode_tmap_2d_h16(&tmQKV, qkv, 768, (uint64_t)B * T, 768 * 2, HD, BT);
  if (rc) return rc;
  rc = tmp::encode_tmap_2d_h16(&tmDO, dO, 256, (uint64_t)B * T, (uint64_t)ld * 2, HD, BT);
  if (rc) return rc;
  rc = tmp::encode_tmap_2d_h16(&tmDKV, dQKV, 768, (uint64_t)B * T, 768 * 2, 64, 32);   // dK / dV boxes [32 rows x 64 cols]
  if (rc) return rc;
  if (fused)   // fp16 reduce-add boxes [32 rows x 32 cols] into dQKV[:, 0:256]
    rc = tmp::encode_tmap_2d_f16_sw64(&tmDQ, dQKV, 256, (uint64_t)B * T, 768 * 2, 32, 32);
  else         // fp32 reduce-add boxes [32 rows x 32 fp32]
    rc = tmp::encode_tmap_2d_f32(&tmDQ, dQ_acc, 256, (uint64_t)B * T, 256 * 4, 32, 32);
  if (rc) return rc;
  if (!fused) {
    const int rows = B * T_lse;
    attn_bwd_delta_kernel<<<(rows + 7) / 8, 256, 0, st>>>((const uint16_t*)O, (const uint16_t*)dO, ld, B, T, H, delta, T_lse);
    rc = tmp::check_launch("attn_bwd_delta_kernel");
    if (rc) return rc;
    cudaError_t e = cudaMemsetAsync(dQ_acc, 0, (size_t)B * T * 256 * sizeof(float), st);
    if (e != cudaSuccess) {
      tmp::set_error("attn_bwd memset: %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  static const int dbg = getenv("TMP_B200_ATTN_BWD_DBG") ? atoi(getenv("TMP_B200_ATTN_BWD_DBG")) : 0;   // timing experiments
  const int n_jt = (T + BT - 1) / BT;
  const int n_items = n_jt * H * B;
  const int sms = tmp::num_sms();
  const int grid = n_items < sms ? n_items : sms;
  if (fused)
    attn_bwd_kernel<true><<<grid, kThreads, kSmemTotal, st>>>(tmQKV, tmDO, tmDQ, tmDKV, kv_len, T, n_jt, H, q_tiles_max, n_items,
                                                              lse2, delta, T_lse, (uint16_t*)dQKV, kLog2e / 8.0f, dbg);
  else
    attn_bwd_kernel<false><<<grid, kThreads, kSmemTotal, st>>>(tmQKV, tmDO, tmDQ, tmDKV, kv_len, T, n_jt, H, q_tiles_max, n_items,
                                                               lse2, delta, T_lse, (uint16_t*)dQKV, kLog2e / 8.0f, dbg);
  rc = tmp::check_launch("attn_bwd_kernel");
  if (rc || fused) return rc;
  const size_t rows = (size_t)B * T;
  attn_bwd_dq_convert_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(dQ_acc, (uint16_t*)dQKV, rows);
  return tmp::check_launch("attn_bwd_dq_convert_kernel");
}

// Single-query form (see attn_bwd_single_query_kernel): dO_row / O_row [B,256] fp16 hold the one query row `q_row` of every
// sample that carries a gradient; writes ALL of dQKV [B*T,768] (zeros where nothing flows).
extern "C" int tmp_attn_bwd_single_query(const void* qkv, const void* dO_row, const void* O_row, const int32_t* kv_len, int B,
                                         int T, int H, int q_row, const float* lse2, int T_lse, void* dQKV, void* stream) {
  TMP_REQUIRE(qkv && dO_row && O_row && lse2 && dQKV, "attn_bwd_single_query: null operand");
  TMP_REQUIRE(B > 0 && T > 0 && H == 4 && q_row >= 0 && q_row < T && T_lse >= T,
              "attn_bwd_single_query: need B>0, T>0, H==4, 0 <= q_row < T <= T_lse");
  attn_bwd_single_query_kernel<<<B * H, 128, 0, (cudaStream_t)stream>>>(
      (const uint16_t*)qkv, (const uint16_t*)dO_row, (const uint16_t*)O_row, kv_len, T, H, q_row, lse2, T_lse,
      (uint16_t*)dQKV, kLog2e / 8.0f);
  return tmp::check_launch("attn_bwd_single_query_kernel");
}
