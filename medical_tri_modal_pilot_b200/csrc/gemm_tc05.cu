// gemm_tc05.cu -- tcgen05/TMEM/TMA GEMMs for the fusion-encoder blocks:
//   * gemm_tn : C[M,N] = epilogue(A[M,K] . B[N,K]^T)      both operands K-major (activations x nn.Linear /
//               Conv1d(k=1) weights as stored by the reference, `[out,in]`). Used for QKV, FFN1, FFN2, the
//               768->256 projections and (with pre-transposed weights) every dgrad.
//   * gemm_wgrad : dW[N,K] += dY[M,N]^T . X[M,K]          both operands MN-major (reduction over tokens),
//               split over M across CTAs, fp32 TMA reduce-adds into the parameter gradient.
// Warp-specialised: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc; both issue through elect.sync, see
// tc05::elect_one), warps 2-17 = epilogue (TMEM -> registers -> fused bias/activation/gate/dropout/residual -> swizzled
// smem box -> TMA store). Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile
// i+1; the kernel is persistent over output tiles.
#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int FMT_MASK = 3;   // gate given as the bit mask written by `mask_out` (uint32 words, ld_gate words per row)
constexpr int BM = 128;
constexpr int BK = 64;  // 64 x 16-bit = 128 B = one swizzle row
constexpr int kEpiWarps = 16;
constexpr int kThreads = 64 + 32 * kEpiWarps;   // warp0 TMA, warp1 MMA, warps 2..17 epilogue
constexpr int kThreadsW = 192;                  // wgrad kernel: 4 epilogue warps
constexpr int kBiasSmemFloats = 1024;

struct EpiParams {
  int M, N, K;
  float alpha;            // scale on the accumulator
  const float* bias;      // [N] fp32 or null
  int relu;               // activation after bias: 0 none, 1 ReLU, 2 GELU (erf)
  const uint16_t* gate;   // [M, ld_gate] 16-bit: multiply by (gate > 0)  (ReLU backward) or null
  int ld_gate;
  const uint16_t* residual;  // [M, ld_res] 16-bit, added last, or null
  int ld_res;
  int a_fmt, b_fmt;       // operand formats (FMT_F16 / FMT_BF16)
  int out_fmt, gate_fmt, res_fmt;
  uint32_t drop_thr16;    // 0 = no dropout; else round(p*65536)
  float drop_scale;       // 1/(1-p)
  int drop_fold;          // 1: 1/(1-p) is already folded into alpha and the bias (every epilogue but GELU is positively
                          // homogeneous), dropout only zeroes
  float bias_scale;       // factor on the bias (1/(1-p) when folded, else 1)
  uint32_t drop_seed, drop_salt;   // mask = f(dropout_key(drop_seed + *drop_seed_dev, drop_salt), element index)
  const uint32_t* drop_seed_dev;   // optional device word added to the seed (CUDA-graph replays), or null
  uint32_t* mask_out;     // optional [M, N/32] bit mask of (result > 0) -- the ReLU/dropout gate of the backward pass in 1 bit per
                          // element instead of re-reading the 16-bit activation (gate_fmt == FMT_MASK consumes it)
  const uint8_t* row_live;   // optional: liveness per group of `rows_per_group` consecutive rows (an image of the Swin feed);
  int rows_per_group;        // an m-tile whose rows all belong to dead groups is skipped by every role
  uint16_t* out;          // [M, ld_out] 16-bit in out_fmt (or null) -- written through tmOut (TMA store)
  float* out_f32;         // [M, ld_out] fp32 (or null) -- direct stores
  int ld_out;
  int mode;               // >= 0: compile-time specialised epilogue (bit set of EPI_*), -1: the generic run-time one
};

// Epilogue features that exist as compile-time specialisations. The generic epilogue tests every feature at run time in
// ~1 600 SASS instructions of branchy code; with 16 epilogue warps a bias-only [128 x 256] tile took ~3 500 cycles of
// epilogue against 2 048 cycles of MMA, i.e. the QKV GEMM was epilogue-bound (43 us; 34 us with a straight-line epilogue).
enum : int {
  EPI_BIAS = 1,        // + bias (staged in shared memory, already multiplied by bias_scale)
  EPI_RELU = 2,
  EPI_GELU = 4,
  EPI_GATEMASK = 8,    // multiply by the 1-bit gate written by an earlier mask_out
  EPI_DROPFOLD = 16,   // dropout with 1/(1-p) folded into alpha / bias: zeroing only
  EPI_MASKOUT = 32,    // write the (result > 0) bit mask
  EPI_RES16 = 64,      // + fp16 residual
  EPI_ALPHA = 128,     // accumulator scale != 1
};
// the hot call sites of the training step and the image-encoder feed (everything else runs the generic epilogue)
#define TMP_EPI_MODES(X)                                                                                              \
  X(0)                                                        /* dgrads: plain */                                     \
  X(EPI_BIAS)                                                 /* QKV, 768->256 projections, Swin qkv / proj / fc2 */   \
  X(EPI_BIAS | EPI_RELU | EPI_DROPFOLD | EPI_MASKOUT | EPI_ALPHA)   /* FFN1 forward */                                \
  X(EPI_BIAS | EPI_DROPFOLD | EPI_RES16 | EPI_ALPHA)          /* FFN2 forward */                                      \
  X(EPI_GATEMASK | EPI_ALPHA)                                 /* FFN2 input gradient */                                \
  X(EPI_BIAS | EPI_GELU)                                      /* Swin fc1 */


// WS ("weight-stationary") variant: for K <= kWsMaxK the whole B operand of an n-block ([BN x K], <= 128 KB) is loaded ONCE per
// CTA and stays in shared memory while the CTA walks over m-tiles of that n-block; the TMA ring then carries A tiles only.
// At K = 256 a [128 x 256] output tile needs 64 KB (A) instead of 192 KB (A + B) from L2 -- the non-stationary kernel is
// bound by that L2 -> SM operand traffic (profiles/r1b: 24 % tensor-pipe active at 4.8 TB/s of operand reads).
template <int BN, bool WS>
struct SmemLayout {
  static constexpr int kWsMaxK = (BN == 256) ? 256 : 512;
  static constexpr int kStages = WS ? 4 : ((BN == 256) ? 4 : 5);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;                          // one 64-wide K slice of B
  static constexpr int kBResBytes = WS ? BN * kWsMaxK * 2 : 0;         // resident B (WS): K/64 slices
  static constexpr int kRingOffset = kBResBytes;
  static constexpr int kStageBytes = WS ? kABytes : kABytes + kBBytes;
  static constexpr int kOutOffset = kRingOffset + kStages * kStageBytes;
  static constexpr int kOutBufs = 1;                                   // per epilogue warp: [32 rows x 64 B] boxes
  static constexpr int kOutBoxBytes = 2048;
  static constexpr int kOutBytes = kEpiWarps * kOutBufs * kOutBoxBytes;
  static constexpr int kBiasOffset = kOutOffset + kOutBytes;
  static constexpr int kBiasFloats = WS ? BN : (BN == 256 ? 256 : kBiasSmemFloats);
  static constexpr int kBarOffset = kBiasOffset + kBiasFloats * 4;
  static constexpr int kTotal = kBarOffset + 256 + 1024;  // + barriers + alignment slack
};
static_assert(SmemLayout<256, true>::kTotal <= 232448 && SmemLayout<128, true>::kTotal <= 232448, "WS smem budget");
static_assert(SmemLayout<256, false>::kTotal <= 232448 && SmemLayout<128, false>::kTotal <= 232448, "smem budget");

// erf GELU with erf from Abramowitz-Stegun 7.1.28 (|abs error| <= 3e-7, far below the fp16 output rounding):
//   1 - erf(z) = q = 1 / (1 + a1 z + ... + a6 z^6)^16,   z = |x| / sqrt 2   (z >= 0)
//   gelu(x) = 0.5 x (1 + erf(x / sqrt 2)) = relu(x) - (|x| / 2) q          (both signs of x, no select)
// evaluated for a PAIR of elements on the packed fp32 pipe (FFMA2 / FMUL2). ONE MUFU per element (the reciprocal; the 16th
// power is four packed squarings): the FC1 epilogue of the image encoder touches 1.1 G elements per step and with the
// rcp + ex2 form of 7.1.26 (two MUFU per element) a [128 x 256] tile cost 4 096 cycles of the 16/clk/SM MUFU pipe against
// 3 072 cycles of MMA at K = 384 -- the GELU GEMMs were MUFU-bound (MMA warp waiting 54 % of its loop for accumulators).
__device__ __forceinline__ void gelu_erf_pair(float& x0, float& x1) {
  const f32x2_t z = f2_mul(f2_pack(fabsf(x0), fabsf(x1)), f2_pack(0.70710678118654752f, 0.70710678118654752f));
  f32x2_t poly = f2_fma(z, f2_pack(0.0000430638f, 0.0000430638f), f2_pack(0.0002765672f, 0.0002765672f));
  poly = f2_fma(poly, z, f2_pack(0.0001520143f, 0.0001520143f));
  poly = f2_fma(poly, z, f2_pack(0.0092705272f, 0.0092705272f));
  poly = f2_fma(poly, z, f2_pack(0.0422820123f, 0.0422820123f));
  poly = f2_fma(poly, z, f2_pack(0.0705230784f, 0.0705230784f));
  poly = f2_fma(poly, z, f2_pack(1.f, 1.f));
  float u0, u1;
  f2_unpack(poly, u0, u1);
  float r0, r1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(u0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(u1));
  f32x2_t q = f2_pack(r0, r1);
  q = f2_mul(q, q);
  q = f2_mul(q, q);
  q = f2_mul(q, q);
  q = f2_mul(q, q);                                                                // 1 - erf(|z|)  in (0, 1]
  const f32x2_t wn = f2_mul(z, f2_pack(-0.70710678118654752f, -0.70710678118654752f));   // -|x| / 2
  f2_unpack(f2_fma(wn, q, f2_pack(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), x0, x1);
}

// Epilogue of one 32-column chunk held by one thread (= one output row): v <- fused epilogue of the accumulators.
__device__ __forceinline__ void epilogue_math(float (&v)[32], const EpiParams& p, const float* sbias, int row, bool row_ok,
                                              int col0, uint32_t drop_key) {
  if (p.alpha != 1.f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
  }
  if (p.bias) {
    if (sbias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(sbias + col0 + j);   // smem broadcast
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
        v[j] = fmaf(b.x, p.bias_scale, v[j]); v[j + 1] = fmaf(b.y, p.bias_scale, v[j + 1]);
        v[j + 2] = fmaf(b.z, p.bias_scale, v[j + 2]); v[j + 3] = fmaf(b.w, p.bias_scale, v[j + 3]);
      }
    }
  }
  if (p.relu == 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (p.relu == 2) {   // erf GELU (nn.GELU() of the Swin MLP): 0.5 x (1 + erf(x / sqrt 2))
#pragma unroll
    for (int j = 0; j < 32; j += 2) gelu_erf_pair(v[j], v[j + 1]);
  }
  if (p.gate && row_ok && p.gate_fmt == FMT_MASK) {  // 1 bit per element: one 4-byte load per 32-column chunk
    const uint32_t m = __ldg(reinterpret_cast<const uint32_t*>(p.gate) + (size_t)row * p.ld_gate + (col0 >> 5));
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (!((m >> j) & 1u)) v[j] = 0.f;
  } else if (p.gate && row_ok && p.gate_fmt == FMT_F32) {   // fp32 mode: the gate tensor is stored in fp32
    const float4* g = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.gate) + (size_t)row * p.ld_gate + col0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 u = __ldg(g + q);
      if (!(u.x > 0.f)) v[q * 4] = 0.f;
      if (!(u.y > 0.f)) v[q * 4 + 1] = 0.f;
      if (!(u.z > 0.f)) v[q * 4 + 2] = 0.f;
      if (!(u.w > 0.f)) v[q * 4 + 3] = 0.f;
    }
  } else if (p.gate && row_ok) {
    const uint4* g = reinterpret_cast<const uint4*>(p.gate + (size_t)row * p.ld_gate + col0);
    uint4 u[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) u[q] = __ldg(g + q);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t w[4] = {u[q].x, u[q].y, u[q].z, u[q].w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 gv = unpack2_rt(w[t], p.gate_fmt);
        if (!(gv.x > 0.f)) v[q * 8 + t * 2] = 0.f;
        if (!(gv.y > 0.f)) v[q * 8 + t * 2 + 1] = 0.f;
      }
    }
  }
  if (p.drop_thr16) {
    const uint32_t idx0 = (uint32_t)row * (uint32_t)p.N + (uint32_t)col0;
    if (p.drop_fold) dropout_zero_run<32>(v, drop_key, idx0, p.drop_thr16);
    else dropout_apply_run<32>(v, drop_key, idx0, p.drop_thr16, p.drop_scale);
  }
  if (p.mask_out && row_ok) {     // before the residual: the gate is the ReLU / dropout pattern of this GEMM's own result
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) m |= (v[j] > 0.f ? 1u : 0u) << j;
    p.mask_out[(size_t)row * (p.N >> 5) + (col0 >> 5)] = m;
  }
  if (p.residual && row_ok && p.res_fmt == FMT_F32) {   // fp32 mode: fp32 residual stream
    const float4* g = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + (size_t)row * p.ld_res + col0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 u = __ldg(g + q);
      v[q * 4] += u.x; v[q * 4 + 1] += u.y; v[q * 4 + 2] += u.z; v[q * 4 + 3] += u.w;
    }
  } else if (p.residual && row_ok) {
    const uint4* g = reinterpret_cast<const uint4*>(p.residual + (size_t)row * p.ld_res + col0);
    uint4 u[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) u[q] = __ldg(g + q);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t w[4] = {u[q].x, u[q].y, u[q].z, u[q].w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 rv = unpack2_rt(w[t], p.res_fmt);
        v[q * 8 + t * 2] += rv.x;
        v[q * 8 + t * 2 + 1] += rv.y;
      }
    }
  }
}

// Compile-time specialised form (MODE >= 0: only the features in the bit set, 16-bit fp16 output, bias in shared memory,
// fp16 residual); MODE < 0 = the generic function above.
template <int MODE>
__device__ __forceinline__ void epilogue_math_t(float (&v)[32], const EpiParams& p, const float* sbias, int row, bool row_ok,
                                                int col0, uint32_t drop_key) {
  if constexpr (MODE < 0) {
    epilogue_math(v, p, sbias, row, row_ok, col0, drop_key);
  } else {
    if constexpr ((MODE & EPI_BIAS) != 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(sbias + col0 + j);   // smem broadcast
        if constexpr ((MODE & EPI_ALPHA) != 0) {
          v[j] = fmaf(v[j], p.alpha, b.x); v[j + 1] = fmaf(v[j + 1], p.alpha, b.y);
          v[j + 2] = fmaf(v[j + 2], p.alpha, b.z); v[j + 3] = fmaf(v[j + 3], p.alpha, b.w);
        } else {
          v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
        }
      }
    } else if constexpr ((MODE & EPI_ALPHA) != 0) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
    }
    if constexpr ((MODE & EPI_RELU) != 0) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if constexpr ((MODE & EPI_GELU) != 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) gelu_erf_pair(v[j], v[j + 1]);
    }
    if constexpr ((MODE & EPI_GATEMASK) != 0) {
      if (row_ok) {
        const uint32_t m = __ldg(reinterpret_cast<const uint32_t*>(p.gate) + (size_t)row * p.ld_gate + (col0 >> 5));
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (!((m >> j) & 1u)) v[j] = 0.f;
      }
    }
    if constexpr ((MODE & EPI_DROPFOLD) != 0) {
      const uint32_t idx0 = (uint32_t)row * (uint32_t)p.N + (uint32_t)col0;
      dropout_zero_run<32>(v, drop_key, idx0, p.drop_thr16);
    }
    if constexpr ((MODE & EPI_MASKOUT) != 0) {
      if (row_ok) {
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) m |= (v[j] > 0.f ? 1u : 0u) << j;
        p.mask_out[(size_t)row * (p.N >> 5) + (col0 >> 5)] = m;
      }
    }
    if constexpr ((MODE & EPI_RES16) != 0) {
      if (row_ok) {
        const uint4* g = reinterpret_cast<const uint4*>(p.residual + (size_t)row * p.ld_res + col0);
        uint4 u[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) u[q] = __ldg(g + q);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t w[4] = {u[q].x, u[q].y, u[q].z, u[q].w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 rv = unpack2<FMT_F16>(w[t]);
            v[q * 8 + t * 2] += rv.x;
            v[q * 8 + t * 2 + 1] += rv.y;
          }
        }
      }
    }
  }
}

__device__ __forceinline__ bool tile_dead(const EpiParams& p, int m0) {
  if (!p.row_live) return false;
  const int g1 = (min(m0 + BM, p.M) - 1) / p.rows_per_group;
  for (int g = m0 / p.rows_per_group; g <= g1; ++g)
    if (__ldg(p.row_live + g)) return false;
  return true;
}

// fp32 accumulator buffers in TMEM: all 512 columns, i.e. 2 at BN = 256 and 4 at BN = 128. With 2 buffers the MMA warp of
// the GELU GEMMs waited 54 % of its loop for the epilogue to release one (clock64 trace): the accumulator round trip
// MMA -> epilogue -> MMA, not the epilogue's issue rate, set the tile period.
template <int BN>
constexpr int kAccBufs = 512 / BN;

struct EpiCtx {
  uint8_t* out_stage;           // staging boxes of all epilogue warps
  uint64_t* tfull_bar;          // [kAcc] accumulator complete
  uint64_t* tempty_bar;         // [kAcc] accumulator drained
  const float* sb_ptr;          // bias in shared memory, indexed by the global column (or null)
  uint32_t tmem_base;
  int it_first, it_end, it_step, n_blks, n_fixed, warp, lane;
};

// The epilogue role (warps 2..17) of gemm_tn_kernel, templated on the epilogue MODE (see TMP_EPI_MODES).
// 16 warps = 4 per scheduler: the fused epilogue is a dependent chain per warp (tcgen05.ld -> bias/activation/dropout
// -> pack -> smem -> TMA store) and ran at 43-48 % issue utilisation with 2 warps per scheduler (profiles/r1g).
// warp -> (TMEM lane quarter it may access, column quarter of the tile). Each thread owns one output row and BN/4
// columns in 32-column chunks; 16-bit results are staged as [32 rows x 32 cols] boxes (64 B rows inside the 128B
// swizzle pattern: two rows per 128 B line) and written with TMA stores (coalesced, asynchronous, M-tail clipped by
// the tensor map). Lane 0 issues the stores; every lane executes the bulk-group waits (groups are per thread, lanes
// without any return at once).
template <int BN, bool WS, int MODE>
__device__ __forceinline__ void epilogue_role(const EpiCtx& cx, const EpiParams& p, const CUtensorMap& tmOut) {
  using L = SmemLayout<BN, WS>;
  constexpr bool kGeneric = MODE < 0;
  const int warp = cx.warp, lane = cx.lane;
  const int e = warp - 2;
  const int quarter = warp & 3;
  const int cg = e >> 2;
  constexpr int kChunks = BN / 128;   // 32-column chunks per warp
  uint8_t* stage_buf = cx.out_stage + e * (L::kOutBufs * L::kOutBoxBytes);
  int sbuf = 0;
  int acc = 0;
  uint32_t acc_phase = 0;
  const bool has_drop = kGeneric ? p.drop_thr16 != 0 : (MODE & EPI_DROPFOLD) != 0;
  const uint32_t drop_key = has_drop ? dropout_key(effective_seed(p.drop_seed, p.drop_seed_dev), p.drop_salt) : 0u;
  for (int it = cx.it_first; it < cx.it_end; it += cx.it_step) {
    const int m0 = (WS ? it : it / cx.n_blks) * BM, n0 = WS ? cx.n_fixed : (it % cx.n_blks) * BN;
    if (tile_dead(p, m0)) continue;
    mbar_wait(&cx.tfull_bar[acc], acc_phase);
    tc_fence_after();
    const int row = m0 + quarter * 32 + lane;
    const bool row_ok = row < p.M;
    const int colw = n0 + cg * (BN / 4);
    const uint32_t taddr = tmem_addr(cx.tmem_base, quarter * 32, acc * BN + cg * (BN / 4));
#pragma unroll 1
    for (int c = 0; c < kChunks; ++c) {
      uint32_t r[32];
      tmem_ld32(taddr + c * 32, r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (c == kChunks - 1) {
        // every accumulator column of this warp is in registers: hand the TMEM buffer back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&cx.tempty_bar[acc]);
      }
      const int col0 = colw + c * 32;
      epilogue_math_t<MODE>(v, p, cx.sb_ptr, row, row_ok, col0, drop_key);
      if (kGeneric && p.out_f32 && row_ok) {
        float4* o = reinterpret_cast<float4*>(p.out_f32 + (size_t)row * p.ld_out + col0);
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      }
      if (!kGeneric || p.out) {
        uint8_t* sbox = stage_buf + sbuf * L::kOutBoxBytes;
        // the TMA store that last used this buffer must have finished reading it
        // (every lane executes the wait: bulk groups are per thread, lanes without any return at once -- no reliance on
        // elect.sync picking the same lane that committed the store)
        if (L::kOutBufs == 2) tma_store_wait_read1();
        else tma_store_wait_read0();
        __syncwarp();
        uint4 u[4];
        if (!kGeneric || p.out_fmt == FMT_F16) {   // uniform branch: one pack per pair
#pragma unroll
          for (int q = 0; q < 4; ++q)
            u[q] = make_uint4(pack_f16x2(v[q * 8 + 0], v[q * 8 + 1]), pack_f16x2(v[q * 8 + 2], v[q * 8 + 3]),
                              pack_f16x2(v[q * 8 + 4], v[q * 8 + 5]), pack_f16x2(v[q * 8 + 6], v[q * 8 + 7]));
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            u[q] = make_uint4(pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]), pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]),
                              pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]));
        }
        // 64B-swizzled box, dense 64 B rows: 16 B chunk q of row `lane` at lane*64 + ((q ^ ((lane >> 1) & 3)) << 4)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(sbox + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = u[q];
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {   // fixed lane: its bulk groups gate the reuse of this warp's staging boxes
          tma_store_2d(&tmOut, sbox, col0, m0 + quarter * 32);
          tma_store_commit();
        }
        if (L::kOutBufs == 2) sbuf ^= 1;
      }
    }
    if (++acc == kAccBufs<BN>) { acc = 0; acc_phase ^= 1; }
  }
  tma_store_wait_read0();
}

template <int BN, bool WS>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmOut, EpiParams p) {
  using L = SmemLayout<BN, WS>;
  constexpr int kStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* full_bar = (uint64_t*)(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  constexpr int kAcc = kAccBufs<BN>;
  uint64_t* tfull_bar = empty_bar + kStages;   // [kAcc]
  uint64_t* tempty_bar = tfull_bar + kAcc;     // [kAcc]
  uint64_t* bres_bar = tempty_bar + kAcc;      // WS: [8] K slice kb of the resident B landed
  uint32_t* tmem_slot = (uint32_t*)(bres_bar + 8);
  static_assert((2 * kStages + 2 * kAcc + 8) * 8 + 4 <= 256, "barrier block");
  float* sbias = (float*)(smem + L::kBiasOffset);
  uint8_t* ring = smem + L::kRingOffset;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_blks = (p.M + BM - 1) / BM;
  const int n_blks = p.N / BN;
  const int k_blks = p.K / BK;
  // tile walk. non-WS: tile = blockIdx.x, += gridDim.x over all (m, n) tiles. WS: this CTA owns n-block
  // blockIdx.x % n_blks and walks m-blocks blockIdx.x / n_blks, += gridDim.x / n_blks (gridDim.x % n_blks == 0).
  const int it_first = WS ? (int)blockIdx.x / n_blks : (int)blockIdx.x;
  const int it_step = WS ? (int)gridDim.x / n_blks : (int)gridDim.x;
  const int it_end = WS ? m_blks : m_blks * n_blks;
  const int n_fixed = WS ? ((int)blockIdx.x % n_blks) * BN : 0;
  const bool bias_in_smem = p.bias && (WS || p.N <= L::kBiasFloats);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (p.out) prefetch_tmap(&tmOut);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAcc; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps);  // one arrive per epilogue warp
    }
    for (int s = 0; s < 8; ++s) mbar_init(&bres_bar[s], 1);
    fence_barrier_init();
  }
  if (bias_in_smem) {
    if (WS) {
      for (int i = threadIdx.x; i < BN; i += kThreads) sbias[i] = p.bias[n_fixed + i] * p.bias_scale;
    } else {
      for (int i = threadIdx.x; i < p.N; i += kThreads) sbias[i] = p.bias[i] * p.bias_scale;
    }
  }
  if (warp == 1) tmem_alloc(tmem_slot, kAcc * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      // WS: the K slices of the resident B are requested one by one, interleaved with the A stages of the first tile, each
      // on its own barrier: the first MMAs start after 1/k_blks of B instead of all of it (128 KB per CTA at kernel start)
      bool b_pending = WS;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = it_first; it < it_end; it += it_step) {
        const int m0 = (WS ? it : it / n_blks) * BM, n0 = WS ? n_fixed : (it % n_blks) * BN;
        if (tile_dead(p, m0)) continue;
        for (int kb = 0; kb < k_blks; ++kb) {
          if (b_pending) {
            mbar_expect_tx(&bres_bar[kb], (uint32_t)L::kBBytes);
            tma_load_2d(smem + kb * L::kBBytes, &tmB, &bres_bar[kb], kb * BK, n_fixed);
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = ring + stage * L::kStageBytes;
          mbar_expect_tx(&full_bar[stage], L::kStageBytes);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
          if (!WS) tma_load_2d(sa + L::kABytes, &tmB, &full_bar[stage], kb * BK, n0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        b_pending = false;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc = make_idesc(BM, BN, 0, 0, p.a_fmt, p.b_fmt);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      bool b_pending = WS;
      for (int it = it_first; it < it_end; it += it_step) {
        if (tile_dead(p, (WS ? it : it / n_blks) * BM)) continue;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blks; ++kb) {
          if (b_pending) mbar_wait(&bres_bar[kb], 0);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * L::kStageBytes);
          const uint32_t sb = WS ? smem_u32(smem + kb * L::kBBytes) : sa + L::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = make_sdesc_sw128(sa + k * 32, 16, 1024);
            const uint64_t bdesc = make_sdesc_sw128(sb + k * 32, 16, 1024);
            umma_ss(d_tmem, adesc, bdesc, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == kAcc) { acc = 0; acc_phase ^= 1; }
        b_pending = false;
      }
    }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    const EpiCtx cx{smem + L::kOutOffset, tfull_bar, tempty_bar, bias_in_smem ? sbias - n_fixed : nullptr, tmem_base,
                    it_first, it_end, it_step, n_blks, n_fixed, warp, lane};
    switch (p.mode) {
#define X(M) case (M): epilogue_role<BN, WS, (M)>(cx, p, tmOut); break;
      TMP_EPI_MODES(X)
#undef X
      default: epilogue_role<BN, WS, -1>(cx, p, tmOut); break;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kAcc * BN);
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad: dW[N,K] (+)= sum_m dY[m,N]^T X[m,K].  A = dY^T (MN-major), B = X^T (MN-major).
// grid = (N/128, K/BNW, splits).  Each CTA reduces rows [m_begin, m_end) and atomically adds its
// 128 x BNW fp32 tile into dW.  Bias gradient (column sums of dY) rides along when `dbias` is given: in the CTAs of
// the first k-column (blockIdx.y == 0) the four epilogue warps, idle during the main loop, read every dY stage out of
// shared memory (the MMA's A operand: 2 swizzled [64 x 64] chunks) and keep fp32 column sums in registers -- the
// separate colsum pass over dY (one more HBM read of every gradient tensor) disappears. The stage is released to the
// TMA producer by the MMA commit AND one arrival per summing warp.
// ------------------------------------------------------------------------------------------------
template <int BNW>
struct WgradSmem {
  static constexpr int kStages = (BNW == 256) ? 4 : 6;
  static constexpr int kABytes = 128 * BK * 2;   // 2 chunks of [64 m-rows x 64 n-cols]
  static constexpr int kBBytes = BNW * BK * 2;   // BNW/64 chunks of [64 m-rows x 64 k-cols]
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;
};

template <int BNW>
__global__ void __launch_bounds__(kThreadsW, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX,
                  const __grid_constant__ CUtensorMap tmDW, int M, int ldw,
                  float* __restrict__ dW, float* __restrict__ dbias, int rows_per_split, int y_fmt, int x_fmt) {
  using L = WgradSmem<BNW>;
  constexpr int kStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* full_bar = (uint64_t*)(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint32_t* tmem_slot = (uint32_t*)(tfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 128;   // rows of dW
  const int k0 = blockIdx.y * BNW;   // cols of dW
  const int m_begin = blockIdx.z * rows_per_split;
  const int m_end = min(M, m_begin + rows_per_split);
  const int k_blks = (m_end - m_begin + BK - 1) / BK;  // reduction blocks (OOB rows are zero-filled by TMA)
  const bool do_bias = dbias != nullptr && blockIdx.y == 0;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmY);
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmDW);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], do_bias ? 5 : 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BNW);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (k_blks > 0) {
    if (warp == 0) {
      if (elect_one()) {
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < k_blks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          const int m = m_begin + kb * BK;
          mbar_expect_tx(&full_bar[stage], L::kStageBytes);
#pragma unroll
          for (int c = 0; c < 2; ++c) tma_load_2d(sa + c * (BK * 128), &tmY, &full_bar[stage], n0 + c * 64, m);
#pragma unroll
          for (int c = 0; c < BNW / 64; ++c) tma_load_2d(sb + c * (BK * 128), &tmX, &full_bar[stage], k0 + c * 64, m);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        const uint32_t idesc = make_idesc(128, BNW, 1, 1, y_fmt, x_fmt);
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < k_blks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // MN-major: 16 reduction rows = 2048 B per UMMA_K step; LBO = one 64-wide chunk = BK*128 B
            const uint64_t adesc = make_sdesc_sw128(sa + k * 2048, BK * 128, 1024);
            const uint64_t bdesc = make_sdesc_sw128(sb + k * 2048, BK * 128, 1024);
            umma_ss(tmem_base, adesc, bdesc, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar);
      }
    } else {
      if (do_bias) {
        // thread = (column pair p of the 128 dY columns, half h of the stage's 64 token rows); a warp reads one whole
        // 128 B swizzled row per LDS.32 -> conflict-free
        const int t = threadIdx.x - 64;
        const int p = t & 63, h = t >> 6;
        const uint32_t col_off = (uint32_t)(p >> 5) * (BK * 128);
        const uint32_t q = p & 31;
        float s0 = 0.f, s1 = 0.f;
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < k_blks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          const uint8_t* sa = smem + stage * L::kStageBytes + col_off;
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            const uint32_t m = (uint32_t)(h * 32 + r);
            const uint32_t w = *reinterpret_cast<const uint32_t*>(sa + sw128_offset(m, q >> 2) + (q & 3) * 4);
            const float2 v = unpack2_rt(w, y_fmt);
            s0 += v.x;
            s1 += v.y;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        atomicAdd(dbias + n0 + 2 * p, s0);
        atomicAdd(dbias + n0 + 2 * p + 1, s1);
        // The dW write-out below stages its boxes in the first 32 KB of the ring (stage 0). `tfull_bar` only says that the
        // MMAs have retired; a summing warp that is still reading the LAST k-block -- which sits in stage 0 whenever
        // (k_blks - 1) % kStages == 0, e.g. 53 blocks per split at M = 64 320 -- would read another warp's fp32 boxes as
        // dY (random values, NaN included: found as sporadic non-finite bias gradients at the bench shape). All four
        // epilogue warps therefore meet here before any of them reuses the ring.
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      const int quarter = warp & 3;
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      // dW tile += accumulators. Each warp stages [32 rows x 32 cols] fp32 boxes (128B swizzle, two per warp, inside the
      // first pipeline stage: every MMA has retired, the ring is idle) and adds them with cp.reduce.async.bulk.tensor.
      // Per-thread red.global.add.v4 put 32 different rows (32 partial sectors) into every instruction.
      uint8_t* sbox0 = smem + (warp - 2) * 8192;
#pragma unroll 1
      for (int c = 0; c < BNW / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_addr(tmem_base, quarter * 32, c * 32), r);
        tmem_ld_wait();
        uint8_t* sbox = sbox0 + (c & 1) * 4096;
        if (c >= 2) tma_store_wait_read1();   // the reduce that last read this box is done with it (groups live in lane 0)
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(sbox + sw128_offset(lane, q)) = make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(&tmDW, sbox, k0 + c * 32, n0 + quarter * 32);
          tma_store_commit();
        }
      }
      tma_store_wait_read0();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, BNW);
  }
}

template <int BN, bool WS>
int launch_tn(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const EpiParams& p, int grid,
              cudaStream_t st) {
  using L = SmemLayout<BN, WS>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel<BN, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) {
      tmp::set_error("cudaFuncSetAttribute(gemm_tn): %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  gemm_tn_kernel<BN, WS><<<grid, kThreads, L::kTotal, st>>>(tmA, tmB, tmOut, p);
  return tmp::check_launch("gemm_tn_kernel");
}

template <int BNW>
int launch_wgrad(const CUtensorMap& tmY, const CUtensorMap& tmX, const CUtensorMap& tmDW, int M, int N, int K, float* dW,
                 float* dbias, int y_fmt,
                 int x_fmt, cudaStream_t st) {
  using L = WgradSmem<BNW>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(gemm_wgrad_kernel<BNW>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) {
      tmp::set_error("cudaFuncSetAttribute(gemm_wgrad): %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  // one CTA per SM (197 KB of shared memory): tiles * splits must not exceed the SM count, or the few CTAs of a second
  // wave double the kernel time (8 tiles x 19 splits = 152 CTAs on 148 SMs ran as two waves)
  const int tiles = (N / 128) * (K / BNW);
  int splits = tmp::num_sms() / tiles;
  if (splits < 1) splits = 1;
  int rows_per_split = (M + splits - 1) / splits;
  rows_per_split = ((rows_per_split + BK - 1) / BK) * BK;
  if (rows_per_split < BK) rows_per_split = BK;
  splits = (M + rows_per_split - 1) / rows_per_split;
  dim3 grid(N / 128, K / BNW, splits);
  gemm_wgrad_kernel<BNW><<<grid, kThreadsW, L::kTotal, st>>>(tmY, tmX, tmDW, M, K, dW, dbias, rows_per_split, y_fmt,
                                                             x_fmt);
  return tmp::check_launch("gemm_wgrad_kernel");
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
static bool fmt_ok(int f) { return f == FMT_F16 || f == FMT_BF16; }
static bool fmt_ok32(int f) { return fmt_ok(f) || f == FMT_F32; }   // gate / residual may be fp32 (fp32 mode)

extern "C" int tmp_gemm_bias_act_fwd(const void* A, int a_fmt, int lda, const void* B, int b_fmt, int ldb, int M, int N,
                                     int K, float alpha, const float* bias, int relu, const void* gate, int gate_fmt,
                                     int ld_gate, const void* residual, int res_fmt, int ld_res, float drop_p,
                                     uint32_t seed, uint32_t salt, const uint32_t* seed_dev, void* out16, int out_fmt,
                                     float* out_f32, int ld_out, uint32_t* mask_out, const uint8_t* row_live,
                                     int rows_per_group, void* stream) {
  void* out_h16 = out16;
  TMP_REQUIRE(!row_live || rows_per_group > 0, "gemm: row_live needs rows_per_group > 0");
  TMP_REQUIRE(A && B && (out_h16 || out_f32), "gemm: null operand");
  TMP_REQUIRE(fmt_ok(a_fmt) && fmt_ok(b_fmt) && fmt_ok(out_fmt) && (!gate || fmt_ok32(gate_fmt) || gate_fmt == FMT_MASK) &&
                  (!residual || fmt_ok32(res_fmt)),
              "gemm: operand / output formats must be 0 (fp16) or 1 (bf16); gate / residual may also be 2 (fp32)");
  TMP_REQUIRE(a_fmt == b_fmt, "gemm: tcgen05 kind::f16 needs A and B in the same 16-bit format");
  TMP_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  TMP_REQUIRE(N % 128 == 0 && K % BK == 0, "gemm: N must be a multiple of 128 and K of 64 (N=%d K=%d)", N, K);
  TMP_REQUIRE(ld_out % 8 == 0 && (!gate || gate_fmt == FMT_MASK || ld_gate % 8 == 0) && (!residual || ld_res % 8 == 0),
              "gemm: leading dimensions must be multiples of 8");
  TMP_REQUIRE(!gate || gate_fmt != FMT_MASK || ld_gate == N / 32, "gemm: a bit-mask gate has N/32 words per row");
  TMP_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "gemm: dropout p out of range");
  const int sms = tmp::num_sms();
  const int m_blks = (M + BM - 1) / BM;
  int BN = 256;
  if (N % 256 != 0 || m_blks * (N / 256) < sms) BN = 128;
  static const int bn_force = getenv("TMP_B200_GEMM_BN") ? atoi(getenv("TMP_B200_GEMM_BN")) : 0;   // A/B timing only
  if (bn_force == 128 || (bn_force == 256 && N % 256 == 0)) BN = bn_force;
  CUtensorMap tmA, tmB;
  int rc = tmp::encode_tmap_2d_h16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, BK, BM);
  if (rc) return rc;
  rc = tmp::encode_tmap_2d_h16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, BK, BN);
  if (rc) return rc;
  EpiParams p;
  p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.bias = bias; p.relu = relu;
  p.gate = (const uint16_t*)gate; p.ld_gate = ld_gate;
  p.residual = (const uint16_t*)residual; p.ld_res = ld_res;
  p.a_fmt = a_fmt; p.b_fmt = b_fmt; p.out_fmt = out_fmt; p.gate_fmt = gate_fmt; p.res_fmt = res_fmt;
  p.drop_thr16 = drop_p > 0.f ? (uint32_t)(drop_p * 65536.f + 0.5f) : 0;
  p.drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  p.drop_fold = (drop_p > 0.f && relu != 2) ? 1 : 0;
  p.bias_scale = p.drop_fold ? p.drop_scale : 1.f;
  if (p.drop_fold) p.alpha = alpha * p.drop_scale;
  p.drop_seed = seed; p.drop_salt = salt; p.drop_seed_dev = seed_dev;
  p.out = (uint16_t*)out_h16; p.out_f32 = out_f32; p.ld_out = ld_out;
  p.mask_out = mask_out;
  p.row_live = row_live; p.rows_per_group = rows_per_group;
  // compile-time specialised epilogue when the call is one of the hot combinations (TMP_EPI_MODES), else the generic one
  p.mode = -1;
  static const bool epi_generic = getenv("TMP_B200_GEMM_GENERIC_EPILOGUE") != nullptr;   // A/B timing
  int mode_cand = -1;
  if (!epi_generic && out_h16 && !out_f32 && out_fmt == FMT_F16 && (!gate || gate_fmt == FMT_MASK) &&
      (!residual || res_fmt == FMT_F16) && (drop_p == 0.f || p.drop_fold) && relu >= 0 && relu <= 2) {
    int m = 0;
    if (bias) m |= EPI_BIAS;
    if (relu == 1) m |= EPI_RELU;
    if (relu == 2) m |= EPI_GELU;
    if (gate) m |= EPI_GATEMASK;
    if (p.drop_fold) m |= EPI_DROPFOLD;
    if (mask_out) m |= EPI_MASKOUT;
    if (residual) m |= EPI_RES16;
    if (p.alpha != 1.f) m |= EPI_ALPHA;
    switch (m) {
#define X(M) case (M):
      TMP_EPI_MODES(X)
#undef X
        mode_cand = m;
        break;
      default: break;
    }
  }
  // the specialised bias path reads the bias from shared memory: only where the kernel variant stages it
  auto with_mode = [&](bool ws, int bn) -> const EpiParams& {
    p.mode = (!bias || ws || N <= (bn == 256 ? 256 : kBiasSmemFloats)) ? mode_cand : -1;
    return p;
  };
  CUtensorMap tmOut;
  if (out_h16) {
    // 16-bit output written by TMA: boxes of [32 rows x 32 cols], 64B swizzle; rows >= M are clipped
    rc = tmp::encode_tmap_2d_h16_sw64(&tmOut, out_h16, (uint64_t)N, (uint64_t)M, (uint64_t)ld_out * 2, 32, 32);
    if (rc) return rc;
  } else {
    tmOut = tmA;
  }
  const int tiles = m_blks * (N / BN);
  const int n_blks = N / BN;
  // weight-stationary variant: worth it once every CTA amortises its resident B over >= 2 m-tiles
  static const bool ws_off = getenv("TMP_B200_GEMM_NO_WS") != nullptr;
  const int ws_max_k = BN == 256 ? SmemLayout<256, true>::kWsMaxK : SmemLayout<128, true>::kWsMaxK;
  if (!ws_off && K <= ws_max_k && n_blks <= sms && tiles >= 2 * sms) {
    const int grid = (sms / n_blks) * n_blks;
    if (BN == 256) return launch_tn<256, true>(tmA, tmB, tmOut, with_mode(true, 256), grid, (cudaStream_t)stream);
    return launch_tn<128, true>(tmA, tmB, tmOut, with_mode(true, 128), grid, (cudaStream_t)stream);
  }
  const int grid = tiles < sms ? tiles : sms;
  if (BN == 256) return launch_tn<256, false>(tmA, tmB, tmOut, with_mode(false, 256), grid, (cudaStream_t)stream);
  return launch_tn<128, false>(tmA, tmB, tmOut, with_mode(false, 128), grid, (cudaStream_t)stream);
}

extern "C" int tmp_gemm_wgrad(const void* dY, int y_fmt, int ldy, const void* X, int x_fmt, int ldx, int M, int N, int K,
                              float* dW, float* dbias, void* stream) {
  TMP_REQUIRE(dY && X && dW, "wgrad: null operand");
  TMP_REQUIRE(fmt_ok(y_fmt) && fmt_ok(x_fmt) && y_fmt == x_fmt, "wgrad: dY and X must share one 16-bit format");
  TMP_REQUIRE(M > 0 && N % 128 == 0 && K % 128 == 0, "wgrad: need N,K multiples of 128 (M=%d N=%d K=%d)", M, N, K);
  CUtensorMap tmY, tmX;
  int rc = tmp::encode_tmap_2d_h16(&tmY, dY, (uint64_t)N, (uint64_t)M, (uint64_t)ldy * 2, 64, BK);
  if (rc) return rc;
  rc = tmp::encode_tmap_2d_h16(&tmX, X, (uint64_t)K, (uint64_t)M, (uint64_t)ldx * 2, 64, BK);
  if (rc) return rc;
  CUtensorMap tmDW;   // dW [N, K] fp32, reduce-add boxes [32 rows x 32 cols]
  rc = tmp::encode_tmap_2d_f32(&tmDW, dW, (uint64_t)K, (uint64_t)N, (uint64_t)K * 4, 32, 32);
  if (rc) return rc;
  if (K % 256 == 0) return launch_wgrad<256>(tmY, tmX, tmDW, M, N, K, dW, dbias, y_fmt, x_fmt, (cudaStream_t)stream);
  return launch_wgrad<128>(tmY, tmX, tmDW, M, N, K, dW, dbias, y_fmt, x_fmt, (cudaStream_t)stream);
}
