// embed.cu -- UMSE / TIE embedding and the encoder prologue (SURVEY.md §8 a1, a2 (UMSE add), a5).
//
// Reference (tri_mbt_vsltcls.py:183-190, 216-224; mbt_encoder.py:697-729):
//   E[b,l,:]  = ReLU(LN(x_val * w_v + b_v)) + ReLU(LN(x_time * w_t + b_t)) + W_feat[int(x_feat)]      (vslt)
//   E[b,j,:]  = proj[b,j,:] + ReLU(LN(time_b * w_t + b_t)) + W_feat[18 | 19]                          (img | txt)
//   X0[b,:,:] = [ bottlenecks(4) ; Dropout(LN_in([CLS ; E]) (+ PE for txt)) ]
// Linear(1,256)->LayerNorm is rank-1 in the scalar, so mean/variance over the 256 channels are closed-form
// polynomials of the scalar (no cross-lane reduction): mean = s*wbar + bbar, var = s^2*A + 2s*C + Dv.
//
// One warp per token row, lane l owns channels 8l..8l+7: every global access is a 16 B vector per lane,
// 512 B contiguous per warp (16-bit rows). Forward and gradient tensors are fp16 (gradients carry a static power-of-two scale).
#include "common.cuh"
#include "rowwise.cuh"

using namespace tc05;
using namespace rw;

namespace {

struct Branch {            // nn.Sequential(Linear(1,256), LayerNorm(256, eps=1e-5), ReLU)
  const float* w;          // [256]  (Linear.weight[:,0])
  const float* b;          // [256]
  const float* g;          // [256]  LayerNorm.weight
  const float* be;         // [256]  LayerNorm.bias
};

struct BranchLane {        // per-lane constants of one branch
  float wc[8], bc[8];      // centred: w - mean(w), b - mean(b)
  float g[8], be[8];
  float A, C, Dv;          // var(w), cov(w,b), var(b)   (population, over the 256 channels)
};

__device__ __forceinline__ void branch_setup(const Branch& br, int lane, BranchLane& L) {
  float w[8], b[8];
  load8_f32(br.w + lane * 8, w);
  load8_f32(br.b + lane * 8, b);
  load8_f32(br.g + lane * 8, L.g);
  load8_f32(br.be + lane * 8, L.be);
  float sw = 0.f, sb = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { sw += w[i]; sb += b[i]; }
  warp_sum2(sw, sb);
  const float wm = sw * (1.f / D), bm = sb * (1.f / D);
  float a = 0.f, c = 0.f, d = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    L.wc[i] = w[i] - wm;
    L.bc[i] = b[i] - bm;
    a += L.wc[i] * L.wc[i];
    c += L.wc[i] * L.bc[i];
    d += L.bc[i] * L.bc[i];
  }
  warp_sum2(a, c);
  d = warp_sum(d);
  L.A = a * (1.f / D); L.C = c * (1.f / D); L.Dv = d * (1.f / D);
}
__device__ __forceinline__ float branch_rstd(const BranchLane& L, float s) {
  const float var = fmaxf(fmaf(s, fmaf(s, L.A, 2.f * L.C), L.Dv), 0.f);
  return rsqrtf(var + 1e-5f);
}
// acc[i] += ReLU(LN(s*w+b))[channel i of this lane]
__device__ __forceinline__ void branch_add(const BranchLane& L, float s, float rstd, float (&acc)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float zh = fmaf(s, L.wc[i], L.bc[i]) * rstd;
    acc[i] += fmaxf(fmaf(zh, L.g[i], L.be[i]), 0.f);
  }
}

// ------------------------------------------------------------------------------------------------
// Forward-only form of a branch on the packed-fp32 pipe (FFMA2 / FADD2, sm_100a): with wg = (w-wbar)*gamma,
// bg = (b-bbar)*gamma folded per lane, ReLU(LN(s*w+b))[i] = max(a*wg[i] + (r*bg[i] + beta[i]), 0) where the two
// per-token scalars are r = rstd(s) and a = r*s. One FFMA2 pair handles two channels: 8 FFMA2 + 8 FMNMX per
// branch per lane-row instead of 40 scalar instructions.
// ------------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 relu2(f32x2 v) {
  float lo, hi;
  upk2(v, lo, hi);
  return pk2(fmaxf(lo, 0.f), fmaxf(hi, 0.f));
}

struct BranchFwd {         // per-lane constants, channel pairs (8l+2i, 8l+2i+1)
  f32x2 wg[4], bg[4], be[4];
  float A, C, Dv;
};
__device__ __forceinline__ void branch_fwd_setup(const Branch& br, int lane, BranchFwd& F) {
  BranchLane L;
  branch_setup(br, lane, L);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    F.wg[i] = pk2(L.wc[2 * i] * L.g[2 * i], L.wc[2 * i + 1] * L.g[2 * i + 1]);
    F.bg[i] = pk2(L.bc[2 * i] * L.g[2 * i], L.bc[2 * i + 1] * L.g[2 * i + 1]);
    F.be[i] = pk2(L.be[2 * i], L.be[2 * i + 1]);
  }
  F.A = L.A; F.C = L.C; F.Dv = L.Dv;
}
__device__ __forceinline__ float branch_fwd_rstd(const BranchFwd& F, float s) {
  const float var = fmaxf(fmaf(s, fmaf(s, F.A, 2.f * F.C), F.Dv), 0.f);
  return rsqrtf(var + 1e-5f);
}
// ReLU(LN(s*w+b)) for the lane's 4 channel pairs; a = rstd*s, r = rstd
__device__ __forceinline__ f32x2 branch_fwd_pair(const BranchFwd& F, int i, f32x2 a2, f32x2 r2) {
  return relu2(ffma2(a2, F.wg[i], ffma2(r2, F.bg[i], F.be[i])));
}

// ------------------------------------------------------------------------------------------------
// raw UMSE/TIE embedding (a1): x[n_tok,3] (time,value,feat) -> E[n_tok,256]; used for the bit-exact gather
// parity test and the HBM-roofline measurement (12 B in + 512 B fp16 out per token).
// A warp takes 32 tokens at a time: lane j reads token j's triple, reduces it to (a_v, r_v, a_t, r_t, feature id)
// and parks the five words in the warp's smem slot; the row loop then costs 2 broadcast LDS + 2 LDS.128 of the
// feature-table row + 16 FFMA2 + 16 FMNMX + 8 FADD2 + 4 packs + one 16 B store per lane.
// ------------------------------------------------------------------------------------------------
template <bool OUT_16>
__global__ void __launch_bounds__(256) umse_embed_fwd_kernel(const float* __restrict__ x, long long n_tok, Branch val,
                                                             Branch tim, const float* __restrict__ Wfeat,
                                                             void* __restrict__ out) {
  __shared__ __align__(16) float sW[20 * D];
  __shared__ __align__(16) float4 sTok[8][32];
  __shared__ int sFid[8][32];
  for (int i = threadIdx.x; i < 20 * D; i += blockDim.x) sW[i] = Wfeat[i];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  BranchFwd V, Tm;
  branch_fwd_setup(val, lane, V);
  branch_fwd_setup(tim, lane, Tm);
  __syncthreads();
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + wid;
  const long long n_grp = (n_tok + 31) / 32;
  const float* sWl = sW + lane * 8;
  for (long long grp = gw; grp < n_grp; grp += warps) {
    const long long tok = grp * 32 + lane;
    float xt = 0.f, xv = 0.f, xf = 0.f;
    if (tok < n_tok) {
      xt = __ldg(x + tok * 3);
      xv = __ldg(x + tok * 3 + 1);
      xf = __ldg(x + tok * 3 + 2);
    }
    const float rv = branch_fwd_rstd(V, xv), rt = branch_fwd_rstd(Tm, xt);
    int fid = __float2int_rz(xf);  // C truncation == .type(torch.IntTensor) (tri_mbt_vsltcls.py:187)
    fid = min(max(fid, 0), 19);
    __syncwarp();
    sTok[wid][lane] = make_float4(rv * xv, rv, rt * xt, rt);
    sFid[wid][lane] = fid * D;
    __syncwarp();
    const int cnt = (int)min(32LL, n_tok - grp * 32);
    char* dst = (char*)out + (size_t)(grp * 32) * D * (OUT_16 ? 2 : 4) + lane * (OUT_16 ? 16 : 32);
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const float4 t4 = sTok[wid][j];
      const float* fr = sWl + sFid[wid][j];
      const float4 f0 = *reinterpret_cast<const float4*>(fr);
      const float4 f1 = *reinterpret_cast<const float4*>(fr + 4);
      const f32x2 av = pk2(t4.x, t4.x), rvv = pk2(t4.y, t4.y), at = pk2(t4.z, t4.z), rtt = pk2(t4.w, t4.w);
      const f32x2 fe[4] = {pk2(f0.x, f0.y), pk2(f0.z, f0.w), pk2(f1.x, f1.y), pk2(f1.z, f1.w)};
      float e[8];
      // reference order: value_embedding + time_embedding + feat_embedding (tri_mbt_vsltcls.py:189)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const f32x2 s = fadd2(fadd2(branch_fwd_pair(V, i, av, rvv), branch_fwd_pair(Tm, i, at, rtt)), fe[i]);
        upk2(s, e[2 * i], e[2 * i + 1]);
      }
      if (OUT_16) store8<ACT>((h16*)(dst + (size_t)j * D * 2), e);
      else store8_f32((float*)(dst + (size_t)j * D * 4), e);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// stream prologue: builds the layer-0 input of one modality stream, X0[B, T=5+n, 256] fp16.
// KIND 0 = vslt (UMSE triples), KIND 1 = img/txt (projected rows + shared time branch + constant feature id).
// ------------------------------------------------------------------------------------------------
struct PrologueParams {
  int B, n, T;                 // T = 5 + n
  // KIND 0
  const float* x;              // [B, n, 3]
  Branch val;
  // KIND 1
  const void* proj;            // [B*n, 256] fp16 (fp32 in the fp32 mode)
  const float* times;          // [B * n_slots]
  int n_slots, rows_per_slot;  // n = n_slots * rows_per_slot
  int feat_id;
  // common
  Branch tim;
  const float* Wfeat;          // [20,256]
  const float* cls;            // [256]
  const float* bottlenecks;    // [4,256]
  const float* ln_g;           // layer_norms_in[m] weight/bias (nn.LayerNorm eps 1e-5)
  const float* ln_b;
  const float* pe;             // [>=T-4, 256] or null
  uint32_t drop_thr16; float drop_scale; uint32_t seed, salt;
  const uint32_t* seed_dev;    // optional device word added to `seed` (per-step counter of a captured CUDA graph)
  void* X0;                    // [B, T, 256] fp16 (fp32 in the fp32 mode)
  int rows_per_group;          // forward: rows a warp takes at a time (32; 8 for short streams, see prologue_fwd_impl)
};

// Forward. A warp takes 32 consecutive rows of the [B*T, 256] output at a time, in the manner of umse_embed_fwd_kernel: lane j
// turns row j's inputs (the (time, value, feature) triple, or the slot time of an img / txt row) into the two per-token
// scalars of each rank-1 LayerNorm branch and parks them in the warp's shared-memory slot; the row loop then runs on
// broadcast reads and packed fp32 math (FFMA2), two rows at a time so that the two shuffle trees of the row LayerNorm (sum and
// sum of squares, one pass) overlap. The first version walked one row per warp iteration with every load, both branch
// evaluations and two dependent warp reductions in one serial chain: 0.65 TB/s cold at the bench shape, 0.97 TB/s at 512 k rows.
struct RowMeta { int code; int proj_row; int t; int pad; };   // code: >= 0 feature-table offset (fid*D), -1..-4 bottleneck row, -5 CLS

template <int KIND, int ST>
__global__ void __launch_bounds__(256, KIND == 0 ? 2 : 3) stream_prologue_fwd_kernel(PrologueParams p) {
  __shared__ __align__(16) float sW[20 * D];
  __shared__ __align__(16) float4 sTok[8][32];
  __shared__ __align__(16) RowMeta sMeta[8][32];
  for (int i = threadIdx.x; i < 20 * D; i += blockDim.x) sW[i] = p.Wfeat[i];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  BranchFwd V, Tm;
  if (KIND == 0) branch_fwd_setup(p.val, lane, V);
  branch_fwd_setup(p.tim, lane, Tm);
  float lg[8], lb[8];
  load8_f32(p.ln_g + lane * 8, lg);
  load8_f32(p.ln_b + lane * 8, lb);
  __syncthreads();
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long rows = (long long)p.B * p.T;
  const int rpg = p.rows_per_group;
  const long long n_grp = (rows + rpg - 1) / rpg;
  const uint32_t dkey = p.drop_thr16 ? dropout_key(effective_seed(p.seed, p.seed_dev), p.salt) : 0u;
  const float* sWl = sW + lane * 8;

  // e <- [CLS ; E] row `j` of the group (before the input LayerNorm); returns false for a bottleneck row (copied verbatim)
  auto embed = [&](int j, long long row, float (&e)[8]) -> bool {
    const RowMeta m = sMeta[wid][j];
    if (m.code < 0 && m.code > -5) {
      load8_f32(p.bottlenecks + (-m.code - 1) * D + lane * 8, e);
      st8<ST>(p.X0, (size_t)row * D + lane * 8, e);
      return false;
    }
    if (m.code == -5) {
      load8_f32(p.cls + lane * 8, e);
      return true;
    }
    const float4 t4 = sTok[wid][j];
    const float4 f0 = *reinterpret_cast<const float4*>(sWl + m.code);
    const float4 f1 = *reinterpret_cast<const float4*>(sWl + m.code + 4);
    f32x2 acc[4] = {pk2(f0.x, f0.y), pk2(f0.z, f0.w), pk2(f1.x, f1.y), pk2(f1.z, f1.w)};
    const f32x2 at = pk2(t4.z, t4.z), rtt = pk2(t4.w, t4.w);
    if (KIND == 0) {
      const f32x2 av = pk2(t4.x, t4.x), rvv = pk2(t4.y, t4.y);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fadd2(fadd2(branch_fwd_pair(V, i, av, rvv), branch_fwd_pair(Tm, i, at, rtt)), acc[i]);
    } else {
      float pr[8];
      ld8<ST>(p.proj, (size_t)m.proj_row * D + lane * 8, pr);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        acc[i] = fadd2(fadd2(pk2(pr[2 * i], pr[2 * i + 1]), branch_fwd_pair(Tm, i, at, rtt)), acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) upk2(acc[i], e[2 * i], e[2 * i + 1]);
    return true;
  };
  // layer_norms_in (nn.LayerNorm, eps 1e-5) + PE + dropout + store, given the row's sum and sum of squares
  auto finish = [&](int j, long long row, float (&e)[8], float s1, float s2) {
    const float mean = s1 * (1.f / D);
    const float rstd = rsqrtf(fmaxf(s2 * (1.f / D) - mean * mean, 0.f) + 1e-5f);
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = fmaf((e[i] - mean) * rstd, lg[i], lb[i]);
    if (p.pe) {
      float pe[8];
      load8_f32(p.pe + (size_t)(sMeta[wid][j].t - 4) * D + lane * 8, pe);
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] += pe[i];
    }
    if (p.drop_thr16) dropout_apply_run<8>(y, dkey, (uint32_t)row * D + lane * 8, p.drop_thr16, p.drop_scale);
    st8<ST>(p.X0, (size_t)row * D + lane * 8, y);
  };
  auto sums = [&](const float (&e)[8], float& s1, float& s2) {
    s1 = 0.f; s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s1 += e[i]; s2 = fmaf(e[i], e[i], s2); }
  };

  for (long long grp = (long long)blockIdx.x * (blockDim.x >> 5) + wid; grp < n_grp; grp += warps) {
    const long long row0 = grp * rpg;
    {   // lane j < rows_per_group prepares row row0 + j
      const long long r = row0 + lane;
      RowMeta m{-1, 0, 0, 0};
      float4 tok = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lane < rpg && r < rows) {
        const int b = (int)(r / p.T), t = (int)(r - (long long)b * p.T);
        m.t = t;
        if (t < 4) m.code = -1 - t;
        else if (t == 4) m.code = -5;
        else {
          const int j = t - 5;
          float st;
          if (KIND == 0) {
            const float* xr = p.x + ((size_t)b * p.n + j) * 3;
            st = __ldg(xr);
            const float sv = __ldg(xr + 1);
            m.code = min(max(__float2int_rz(__ldg(xr + 2)), 0), 19) * D;   // C truncation == .type(torch.IntTensor)
            const float rv = branch_fwd_rstd(V, sv);
            tok.x = rv * sv; tok.y = rv;
          } else {
            m.proj_row = b * p.n + j;
            st = __ldg(p.times + b * p.n_slots + j / p.rows_per_slot);
            m.code = p.feat_id * D;
          }
          const float rt = branch_fwd_rstd(Tm, st);
          tok.z = rt * st; tok.w = rt;
        }
      }
      __syncwarp();
      sTok[wid][lane] = tok;
      sMeta[wid][lane] = m;
      __syncwarp();
    }
    const int cnt = (int)min((long long)rpg, rows - row0);
    for (int j = 0; j < cnt; j += 2) {
      float e0[8], e1[8];
      const bool two = j + 1 < cnt;
      const bool n0 = embed(j, row0 + j, e0);
      const bool n1 = two && embed(j + 1, row0 + j + 1, e1);
      float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
      if (n0) sums(e0, a1, a2);
      if (n1) sums(e1, b1, b2);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {      // four interleaved shuffle trees
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        b1 += __shfl_xor_sync(0xffffffffu, b1, o);
        b2 += __shfl_xor_sync(0xffffffffu, b2, o);
      }
      if (n0) finish(j, row0 + j, e0, a1, a2);
      if (n1) finish(j + 1, row0 + j + 1, e1, b1, b2);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward of the stream prologue. Gradient accumulators (fp32, atomically added):
//   g_val / g_tim : [4,256] each = (dLinear.weight, dLinear.bias, dLN.weight, dLN.bias)
//   g_feat [20,256], g_cls [256], g_bott [4,256], g_ln [2,256] = (dLN_in.weight, dLN_in.bias)
//   dproj [B*n,256] fp16 (KIND 1): gradient wrt the projected rows
// ------------------------------------------------------------------------------------------------
struct PrologueBwdParams {
  PrologueParams f;
  const void* dX0;   // [B, T, 256] fp16 gradient (fp32 in the fp32 mode)
  float* g_val; float* g_tim; float* g_feat; float* g_cls; float* g_bott; float* g_ln;
  void* dproj;       // [B*n, 256] fp16 gradient (fp32 in the fp32 mode)
};

struct BranchAcc { float dw[8], db[8], dg[8], dbe[8]; };

__device__ __forceinline__ void flush8(float* sm, const float (&v)[8], int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) atomicAdd(&sm[lane * 8 + i], v[i]);
}

// Backward, same row grouping as the forward: lane j prepares row j's scalars (inputs, both branch rstd's, feature id), the
// row loop reads them by broadcast, the gradient row of the NEXT row is prefetched while the current one is reduced, the
// sum / sum-of-squares of the input LayerNorm come from one pass (one shuffle tree instead of two), the two rank-1 branch
// backward reductions share one tree, and the feature-embedding gradient goes to a WARP-PRIVATE [20,256] table in shared
// memory (plain read-modify-write, each lane owns its 8 channels) instead of shared-memory atomics that eight warps
// contended for on the five common feature ids. First version: 155-178 us at the bench shape (0.2 TB/s).
//
// Second version (the kernel sits on the serial tail of the step, alone on the GPU at 8 warps per SM -- registers AND shared
// memory both cap it there): (1) rows are split evenly over the warps (a contiguous range each, walked in chunks of <= 32)
// instead of whole 32-row groups, which left 826 of 1 184 warps with 64 rows and the rest with 32 at the bench shape and
// ran the 9 728-row image stream on 38 blocks; (2) the gradient rows (and the projected rows of the img / txt streams)
// come through a warp-private cp.async ring of kRing rows instead of a one-row register prefetch, so a row no longer
// costs at least one global-load latency. Each lane copies and later reads its own 16 / 32 bytes: no cross-lane hand-off.
struct BwdTok { float sv, rv, st, rt; };
constexpr int kRing16 = 4, kRing32 = 2;       // ring depth in rows: 16-bit rows (512 B) / fp32 rows (1 KB)
__device__ __forceinline__ void cp_async16_g2s(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// one lane's 8 elements of a row: global -> ring slot (16 B, or 2 x 16 B in the fp32 mode)
template <int ST>
__device__ __forceinline__ void ring_fill(void* slot_row, const void* base, size_t elem, int lane) {
  if (ST == FMT_F32) {
    const float* src = static_cast<const float*>(base) + elem;
    float* dst = static_cast<float*>(slot_row) + lane * 8;
    cp_async16_g2s(dst, src);
    cp_async16_g2s(dst + 4, src + 4);
  } else {
    cp_async16_g2s(static_cast<h16*>(slot_row) + lane * 8, static_cast<const h16*>(base) + elem);
  }
}

template <int KIND, int ST>
__global__ void __launch_bounds__(256) stream_prologue_bwd_kernel(PrologueBwdParams q) {
  const PrologueParams& p = q.f;
  extern __shared__ __align__(16) float sm[];
  float* sW = sm;                              // [20*256] forward table
  float* sAcc = sm + 20 * D;                   // [16*256]: val(4) tim(4) cls(1) bott(4) ln(2) feat-const(1, KIND 1)
  BwdTok* sTok = reinterpret_cast<BwdTok*>(sAcc + 16 * D);           // [8][32]
  RowMeta* sMeta = reinterpret_cast<RowMeta*>(sTok + 8 * 32);        // [8][32]
  float* sG = reinterpret_cast<float*>(sMeta + 8 * 32);              // KIND 0: [8 warps][20*256] feature-embedding gradient
  constexpr int kRing = ST == FMT_F32 ? kRing32 : kRing16;
  constexpr int kRowBytes = D * (ST == FMT_F32 ? 4 : 2);
  char* sRing = reinterpret_cast<char*>(sG + (KIND == 0 ? 8 * 20 * D : 0));   // [8 warps][kRing][row]: dX0 rows (+ the same for proj, KIND 1)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  char* ringG = sRing + (size_t)wid * kRing * kRowBytes;
  char* ringP = sRing + (size_t)(8 + wid) * kRing * kRowBytes;
  for (int i = threadIdx.x; i < 20 * D; i += blockDim.x) sW[i] = p.Wfeat[i];
  for (int i = threadIdx.x; i < 16 * D; i += blockDim.x) sAcc[i] = 0.f;
  if (KIND == 0)
    for (int i = threadIdx.x; i < 8 * 20 * D; i += blockDim.x) sG[i] = 0.f;
  BranchLane V, Tm;
  if (KIND == 0) branch_setup(p.val, lane, V);
  branch_setup(p.tim, lane, Tm);
  float lg[8];
  load8_f32(p.ln_g + lane * 8, lg);
  BranchAcc aV, aT;
  float a_cls[8], a_lng[8], a_lnb[8], a_feat[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    aV.dw[i] = aV.db[i] = aV.dg[i] = aV.dbe[i] = 0.f;
    aT.dw[i] = aT.db[i] = aT.dg[i] = aT.dbe[i] = 0.f;
    a_cls[i] = a_lng[i] = a_lnb[i] = a_feat[i] = 0.f;
  }
  __syncthreads();
  float* sGw = sG + (size_t)wid * 20 * D + lane * 8;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long rows = (long long)p.B * p.T;
  const long long per_warp = (rows + warps - 1) / warps;
  const long long r_begin = min(rows, ((long long)blockIdx.x * (blockDim.x >> 5) + wid) * per_warp);
  const long long r_end = min(rows, r_begin + per_warp);
  const uint32_t dkey = p.drop_thr16 ? dropout_key(effective_seed(p.seed, p.seed_dev), p.salt) : 0u;

  for (long long row0 = r_begin; row0 < r_end; row0 += 32) {
    {
      const long long r = row0 + lane;
      RowMeta m{-1, 0, 0, 0};
      BwdTok tok{0.f, 0.f, 0.f, 0.f};
      if (r < r_end) {
        const int b = (int)(r / p.T), t = (int)(r - (long long)b * p.T);
        m.t = t;
        if (t < 4) m.code = -1 - t;
        else if (t == 4) m.code = -5;
        else {
          const int j = t - 5;
          if (KIND == 0) {
            const float* xr = p.x + ((size_t)b * p.n + j) * 3;
            tok.st = __ldg(xr);
            tok.sv = __ldg(xr + 1);
            m.code = min(max(__float2int_rz(__ldg(xr + 2)), 0), 19) * D;
            tok.rv = branch_rstd(V, tok.sv);
          } else {
            m.proj_row = b * p.n + j;
            tok.st = __ldg(p.times + b * p.n_slots + j / p.rows_per_slot);
            m.code = p.feat_id * D;
          }
          tok.rt = branch_rstd(Tm, tok.st);
        }
      }
      __syncwarp();
      sTok[wid * 32 + lane] = tok;
      sMeta[wid * 32 + lane] = m;
      __syncwarp();
    }
    const int cnt = (int)min(32LL, r_end - row0);
    // ring: row j of the chunk lives in slot j % kRing; one commit group per row (empty past the end of the chunk)
    auto fill = [&](int j) {
      if (j < cnt) {
        ring_fill<ST>(ringG + (j % kRing) * kRowBytes, q.dX0, (size_t)(row0 + j) * D + lane * 8, lane);
        if (KIND == 1) {
          const RowMeta mj = sMeta[wid * 32 + j];
          if (mj.code >= 0) ring_fill<ST>(ringP + (j % kRing) * kRowBytes, p.proj, (size_t)mj.proj_row * D + lane * 8, lane);
        }
      }
      cp_async_commit();
    };
#pragma unroll
    for (int j = 0; j < kRing; ++j) fill(j);
    for (int j = 0; j < cnt; ++j) {
      const long long row = row0 + j;
      cp_async_wait<kRing - 1>();        // this lane's copy of row j has landed (it reads back only its own bytes)
      float g[8];
      ld8<ST>(ringG + (j % kRing) * kRowBytes, (size_t)lane * 8, g);
      const RowMeta m = sMeta[wid * 32 + j];
      if (m.code < 0 && m.code > -5) {
        flush8(sAcc + (9 + (-m.code - 1)) * D, g, lane);      // bottleneck parameter rows (4 of T rows)
      } else {
        const BwdTok tk = sTok[wid * 32 + j];
        float e[8];
        if (m.code == -5) {
          load8_f32(p.cls + lane * 8, e);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) e[i] = 0.f;
          if (KIND == 0) branch_add(V, tk.sv, tk.rv, e);
          else ld8<ST>(ringP + (j % kRing) * kRowBytes, (size_t)lane * 8, e);
          branch_add(Tm, tk.st, tk.rt, e);
          const float4 f0 = *reinterpret_cast<const float4*>(&sW[m.code + lane * 8]);
          const float4 f1 = *reinterpret_cast<const float4*>(&sW[m.code + lane * 8 + 4]);
          e[0] += f0.x; e[1] += f0.y; e[2] += f0.z; e[3] += f0.w;
          e[4] += f1.x; e[5] += f1.y; e[6] += f1.z; e[7] += f1.w;
        }
        if (p.drop_thr16) dropout_apply_run<8>(g, dkey, (uint32_t)row * D + lane * 8, p.drop_thr16, p.drop_scale);
        // input LayerNorm statistics, one pass
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { s1 += e[i]; s2 = fmaf(e[i], e[i], s2); }
        warp_sum2(s1, s2);
        const float mean = s1 * (1.f / D);
        const float rstd = rsqrtf(fmaxf(s2 * (1.f / D) - mean * mean, 0.f) + 1e-5f);
        float de[8], m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float yh = (e[i] - mean) * rstd;
          a_lng[i] += g[i] * yh;
          a_lnb[i] += g[i];
          de[i] = g[i] * lg[i];   // d yhat
          m1 += de[i];
          m2 = fmaf(de[i], yh, m2);
          e[i] = yh;
        }
        warp_sum2(m1, m2);
        m1 *= (1.f / D); m2 *= (1.f / D);
#pragma unroll
        for (int i = 0; i < 8; ++i) de[i] = rstd * (de[i] - m1 - e[i] * m2);
        if (m.code == -5) {
#pragma unroll
          for (int i = 0; i < 8; ++i) a_cls[i] += de[i];
        } else {
          if (KIND == 0) {
            float* gw = sGw + m.code;          // warp-private table: no atomics
            float4 u0 = *reinterpret_cast<float4*>(gw), u1 = *reinterpret_cast<float4*>(gw + 4);
            u0.x += de[0]; u0.y += de[1]; u0.z += de[2]; u0.w += de[3];
            u1.x += de[4]; u1.y += de[5]; u1.z += de[6]; u1.w += de[7];
            *reinterpret_cast<float4*>(gw) = u0; *reinterpret_cast<float4*>(gw + 4) = u1;
          } else {
            st8<ST>(q.dproj, (size_t)m.proj_row * D + lane * 8, de);
#pragma unroll
            for (int i = 0; i < 8; ++i) a_feat[i] += de[i];
          }
          // both rank-1 LayerNorm branches: ReLU gate, LN-parameter gradients, then ONE shuffle tree for the four sums
          float zv[8], dzv[8], zt[8], dzt[8];
          float v1 = 0.f, v2 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (KIND == 0) {
              zv[i] = fmaf(tk.sv, V.wc[i], V.bc[i]) * tk.rv;
              const float gm = fmaf(zv[i], V.g[i], V.be[i]) > 0.f ? de[i] : 0.f;
              aV.dbe[i] += gm;
              aV.dg[i] = fmaf(gm, zv[i], aV.dg[i]);
              dzv[i] = gm * V.g[i];
              v1 += dzv[i];
              v2 = fmaf(dzv[i], zv[i], v2);
            }
            zt[i] = fmaf(tk.st, Tm.wc[i], Tm.bc[i]) * tk.rt;
            const float gt = fmaf(zt[i], Tm.g[i], Tm.be[i]) > 0.f ? de[i] : 0.f;
            aT.dbe[i] += gt;
            aT.dg[i] = fmaf(gt, zt[i], aT.dg[i]);
            dzt[i] = gt * Tm.g[i];
            t1 += dzt[i];
            t2 = fmaf(dzt[i], zt[i], t2);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            if (KIND == 0) {
              v1 += __shfl_xor_sync(0xffffffffu, v1, o);
              v2 += __shfl_xor_sync(0xffffffffu, v2, o);
            }
            t1 += __shfl_xor_sync(0xffffffffu, t1, o);
            t2 += __shfl_xor_sync(0xffffffffu, t2, o);
          }
          v1 *= (1.f / D); v2 *= (1.f / D); t1 *= (1.f / D); t2 *= (1.f / D);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (KIND == 0) {
              const float dz = tk.rv * (dzv[i] - v1 - zv[i] * v2);
              aV.dw[i] = fmaf(dz, tk.sv, aV.dw[i]);
              aV.db[i] += dz;
            }
            const float dz = tk.rt * (dzt[i] - t1 - zt[i] * t2);
            aT.dw[i] = fmaf(dz, tk.st, aT.dw[i]);
            aT.db[i] += dz;
          }
        }
      }
      fill(j + kRing);      // refill the slot just consumed (its values went through the arithmetic above)
    }
    cp_async_wait<0>();
    __syncwarp();           // the next chunk overwrites this warp's sTok / sMeta
  }
  if (KIND == 0) {
    flush8(sAcc + 0 * D, aV.dw, lane); flush8(sAcc + 1 * D, aV.db, lane);
    flush8(sAcc + 2 * D, aV.dg, lane); flush8(sAcc + 3 * D, aV.dbe, lane);
  }
  flush8(sAcc + 4 * D, aT.dw, lane); flush8(sAcc + 5 * D, aT.db, lane);
  flush8(sAcc + 6 * D, aT.dg, lane); flush8(sAcc + 7 * D, aT.dbe, lane);
  flush8(sAcc + 8 * D, a_cls, lane);
  flush8(sAcc + 13 * D, a_lng, lane); flush8(sAcc + 14 * D, a_lnb, lane);
  if (KIND == 1) flush8(sAcc + 15 * D, a_feat, lane);
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * D; i += blockDim.x) {
    if (KIND == 0) atomicAdd(&q.g_val[i], sAcc[i]);
    atomicAdd(&q.g_tim[i], sAcc[4 * D + i]);
    atomicAdd(&q.g_bott[i], sAcc[9 * D + i]);
  }
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(&q.g_cls[i], sAcc[8 * D + i]);
    atomicAdd(&q.g_ln[i], sAcc[13 * D + i]);
    atomicAdd(&q.g_ln[D + i], sAcc[14 * D + i]);
  }
  if (KIND == 0) {
    for (int i = threadIdx.x; i < 20 * D; i += blockDim.x) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += sG[(size_t)w * 20 * D + i];
      if (v != 0.f) atomicAdd(&q.g_feat[i], v);
    }
  } else {
    for (int i = threadIdx.x; i < D; i += blockDim.x) atomicAdd(&q.g_feat[p.feat_id * D + i], sAcc[15 * D + i]);
  }
}

}  // namespace

// branch parameter block: 4 pointers (Linear.weight[256,1], Linear.bias, LayerNorm.weight, LayerNorm.bias)
extern "C" int tmp_umse_embed_fwd(const float* x, long long n_tok, const float* const* val4, const float* const* tim4,
                                  const float* Wfeat, void* out, int out_is_fp16, void* stream) {
  TMP_REQUIRE(x && val4 && tim4 && Wfeat && out && n_tok >= 0, "umse_embed_fwd: bad argument");
  if (n_tok == 0) return TMP_OK;
  Branch v{val4[0], val4[1], val4[2], val4[3]}, t{tim4[0], tim4[1], tim4[2], tim4[3]};
  long long grp = (n_tok + 31) / 32;
  long long blocks = (grp + 7) / 8;
  const long long cap = (long long)tmp::num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (out_is_fp16)
    umse_embed_fwd_kernel<true><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, n_tok, v, t, Wfeat, out);
  else
    umse_embed_fwd_kernel<false><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, n_tok, v, t, Wfeat, out);
  return tmp::check_launch("umse_embed_fwd_kernel");
}

static int fill_prologue(PrologueParams& p, int kind, int B, int n, const float* x, const float* const* val4,
                         const void* proj, const float* times, int n_slots, int feat_id, const float* const* tim4,
                         const float* Wfeat, const float* cls, const float* bottlenecks, const float* ln_g,
                         const float* ln_b, const float* pe, float drop_p, uint32_t seed, uint32_t salt,
                         const uint32_t* seed_dev, void* X0) {
  TMP_REQUIRE(kind == 0 || kind == 1, "prologue: kind must be 0 (vslt) or 1 (img/txt)");
  TMP_REQUIRE(B > 0 && n >= 0 && tim4 && Wfeat && cls && bottlenecks && ln_g && ln_b && X0, "prologue: bad argument");
  TMP_REQUIRE(kind == 1 || (x && val4), "prologue: vslt stream needs x and the value branch");
  TMP_REQUIRE(kind == 0 || (proj && times && n_slots > 0 && n % n_slots == 0 && feat_id >= 0 && feat_id < 20),
              "prologue: img/txt stream needs proj, times, n_slots | n, feat id in [0,20)");
  TMP_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "prologue: dropout p out of range");
  p.B = B; p.n = n; p.T = 5 + n;
  p.x = x;
  if (val4) p.val = Branch{val4[0], val4[1], val4[2], val4[3]}; else p.val = Branch{nullptr, nullptr, nullptr, nullptr};
  p.proj = proj; p.times = times; p.n_slots = n_slots > 0 ? n_slots : 1;
  p.rows_per_slot = n_slots > 0 ? n / n_slots : n; if (p.rows_per_slot < 1) p.rows_per_slot = 1;
  p.feat_id = feat_id;
  p.tim = Branch{tim4[0], tim4[1], tim4[2], tim4[3]};
  p.Wfeat = Wfeat; p.cls = cls; p.bottlenecks = bottlenecks; p.ln_g = ln_g; p.ln_b = ln_b; p.pe = pe;
  p.drop_thr16 = drop_p > 0.f ? (uint32_t)(drop_p * 65536.f + 0.5f) : 0;
  p.drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  p.seed = seed; p.salt = salt; p.seed_dev = seed_dev;
  p.X0 = X0;
  return TMP_OK;
}

static int prologue_fwd_impl(int stf, int kind, int B, int n, const float* x, const float* const* val4, const void* proj,
                             const float* times, int n_slots, int feat_id, const float* const* tim4, const float* Wfeat,
                             const float* cls, const float* bottlenecks, const float* ln_g, const float* ln_b,
                             const float* pe, float drop_p, uint32_t seed, uint32_t salt, const uint32_t* seed_dev,
                             void* X0, void* stream) {
  PrologueParams p;
  int rc = fill_prologue(p, kind, B, n, x, val4, proj, times, n_slots, feat_id, tim4, Wfeat, cls, bottlenecks, ln_g,
                         ln_b, pe, drop_p, seed, salt, seed_dev, X0);
  if (rc) return rc;
  // one warp per group of rows. A warp walks its group two rows at a time, i.e. a chain of rows_per_group / 2 dependent
  // load -> LayerNorm -> store steps: with 32-row groups the 9 728-row image stream ran on 38 blocks for 32 us -- on the
  // critical path between the image encoder and the first exchange -- whatever the grid. Short streams take 8-row groups.
  const long long rows_total = (long long)B * p.T;
  p.rows_per_group = rows_total >= 32LL * 8 * tmp::num_sms() ? 32 : 8;
  long long fblocks = ((rows_total + p.rows_per_group - 1) / p.rows_per_group + 7) / 8;
  const long long fcap = (long long)tmp::num_sms() * (kind == 0 ? 2 : 3);
  if (fblocks > fcap) fblocks = fcap;
  const int grid = (int)(fblocks < 1 ? 1 : fblocks);
  cudaStream_t s = (cudaStream_t)stream;
  if (stf == FMT_F32) {
    if (kind == 0) stream_prologue_fwd_kernel<0, FMT_F32><<<grid, 256, 0, s>>>(p);
    else stream_prologue_fwd_kernel<1, FMT_F32><<<grid, 256, 0, s>>>(p);
  } else {
    if (kind == 0) stream_prologue_fwd_kernel<0, ACT><<<grid, 256, 0, s>>>(p);
    else stream_prologue_fwd_kernel<1, ACT><<<grid, 256, 0, s>>>(p);
  }
  return tmp::check_launch("stream_prologue_fwd_kernel");
}

extern "C" int tmp_stream_prologue_fwd(int kind, int B, int n, const float* x, const float* const* val4,
                                       const void* proj, const float* times, int n_slots, int feat_id,
                                       const float* const* tim4, const float* Wfeat, const float* cls,
                                       const float* bottlenecks, const float* ln_g, const float* ln_b, const float* pe,
                                       float drop_p, uint32_t seed, uint32_t salt, const uint32_t* seed_dev, void* X0,
                                       void* stream) {
  return prologue_fwd_impl(ACT, kind, B, n, x, val4, proj, times, n_slots, feat_id, tim4, Wfeat, cls, bottlenecks, ln_g,
                           ln_b, pe, drop_p, seed, salt, seed_dev, X0, stream);
}
// fp32 mode: proj and X0 are fp32
extern "C" int tmp_stream_prologue_fwd_f32(int kind, int B, int n, const float* x, const float* const* val4,
                                           const float* proj, const float* times, int n_slots, int feat_id,
                                           const float* const* tim4, const float* Wfeat, const float* cls,
                                           const float* bottlenecks, const float* ln_g, const float* ln_b,
                                           const float* pe, float drop_p, uint32_t seed, uint32_t salt,
                                           const uint32_t* seed_dev, float* X0, void* stream) {
  return prologue_fwd_impl(FMT_F32, kind, B, n, x, val4, proj, times, n_slots, feat_id, tim4, Wfeat, cls, bottlenecks,
                           ln_g, ln_b, pe, drop_p, seed, salt, seed_dev, X0, stream);
}

static int prologue_bwd_impl(int stf, int kind, int B, int n, const float* x, const float* const* val4, const void* proj,
                             const float* times, int n_slots, int feat_id, const float* const* tim4, const float* Wfeat,
                             const float* cls, const float* bottlenecks, const float* ln_g, const float* ln_b,
                             const float* pe, float drop_p, uint32_t seed, uint32_t salt, const uint32_t* seed_dev,
                             const void* dX0, float* g_val, float* g_tim, float* g_feat, float* g_cls, float* g_bott,
                             float* g_ln, void* dproj, void* stream) {
  PrologueBwdParams q;
  // X0 is not written by the backward; pass dX0 to satisfy the non-null check
  int rc = fill_prologue(q.f, kind, B, n, x, val4, proj, times, n_slots, feat_id, tim4, Wfeat, cls, bottlenecks, ln_g,
                         ln_b, pe, drop_p, seed, salt, seed_dev, const_cast<void*>(dX0));
  if (rc) return rc;
  TMP_REQUIRE(dX0 && g_tim && g_feat && g_cls && g_bott && g_ln, "prologue_bwd: null gradient buffer");
  TMP_REQUIRE(kind == 1 || g_val, "prologue_bwd: vslt needs g_val");
  TMP_REQUIRE(kind == 0 || dproj, "prologue_bwd: img/txt needs dproj");
  q.dX0 = dX0;
  q.g_val = g_val; q.g_tim = g_tim; q.g_feat = g_feat; q.g_cls = g_cls; q.g_bott = g_bott; q.g_ln = g_ln;
  q.dproj = dproj;
  // forward table + accumulators + per-warp row scalars (+ 8 warp-private feature-gradient tables for the vslt stream)
  // + the cp.async rings: 8 warps x 4 rows x 512 B (2 rows x 1 KB in the fp32 mode) = 16 KB, twice for the img / txt kernel
  const int ring = 8 * kRing16 * D * 2;
  static_assert(kRing16 * 2 == kRing32 * 4, "ring bytes are the same in both storage formats");
  const int smem1 = (20 + 16) * D * 4 + 2 * 8 * 32 * 16 + 2 * ring;
  const int smem0 = (20 + 16) * D * 4 + 2 * 8 * 32 * 16 + 8 * 20 * D * 4 + ring;
  const int smem = kind == 0 ? smem0 : smem1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(stream_prologue_bwd_kernel<0, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem0);
    cudaFuncSetAttribute(stream_prologue_bwd_kernel<1, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
    cudaFuncSetAttribute(stream_prologue_bwd_kernel<0, FMT_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem0);
    cudaFuncSetAttribute(stream_prologue_bwd_kernel<1, FMT_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
    attr_set = true;
  }
  // rows are split evenly over the warps of the grid, one block per SM at most (220 KB of shared memory for the vslt kernel,
  // ~200 registers per thread for both): the few thousand rows of an img / txt stream spread over every SM, ~8 per warp
  long long blocks = ((long long)B * q.f.T + 63) / 64;
  const long long cap = (long long)tmp::num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cudaStream_t s = (cudaStream_t)stream;
  if (stf == FMT_F32) {
    if (kind == 0) stream_prologue_bwd_kernel<0, FMT_F32><<<(int)blocks, 256, smem, s>>>(q);
    else stream_prologue_bwd_kernel<1, FMT_F32><<<(int)blocks, 256, smem, s>>>(q);
  } else {
    if (kind == 0) stream_prologue_bwd_kernel<0, ACT><<<(int)blocks, 256, smem, s>>>(q);
    else stream_prologue_bwd_kernel<1, ACT><<<(int)blocks, 256, smem, s>>>(q);
  }
  return tmp::check_launch("stream_prologue_bwd_kernel");
}

extern "C" int tmp_stream_prologue_bwd(int kind, int B, int n, const float* x, const float* const* val4,
                                       const void* proj, const float* times, int n_slots, int feat_id,
                                       const float* const* tim4, const float* Wfeat, const float* cls,
                                       const float* bottlenecks, const float* ln_g, const float* ln_b, const float* pe,
                                       float drop_p, uint32_t seed, uint32_t salt, const uint32_t* seed_dev,
                                       const void* dX0, float* g_val,
                                       float* g_tim, float* g_feat, float* g_cls, float* g_bott, float* g_ln,
                                       void* dproj, void* stream) {
  return prologue_bwd_impl(ACT, kind, B, n, x, val4, proj, times, n_slots, feat_id, tim4, Wfeat, cls, bottlenecks, ln_g,
                           ln_b, pe, drop_p, seed, salt, seed_dev, dX0, g_val, g_tim, g_feat, g_cls, g_bott, g_ln, dproj,
                           stream);
}
extern "C" int tmp_stream_prologue_bwd_f32(int kind, int B, int n, const float* x, const float* const* val4,
                                           const float* proj, const float* times, int n_slots, int feat_id,
                                           const float* const* tim4, const float* Wfeat, const float* cls,
                                           const float* bottlenecks, const float* ln_g, const float* ln_b,
                                           const float* pe, float drop_p, uint32_t seed, uint32_t salt,
                                           const uint32_t* seed_dev, const float* dX0, float* g_val, float* g_tim,
                                           float* g_feat, float* g_cls, float* g_bott, float* g_ln, float* dproj,
                                           void* stream) {
  return prologue_bwd_impl(FMT_F32, kind, B, n, x, val4, proj, times, n_slots, feat_id, tim4, Wfeat, cls, bottlenecks,
                           ln_g, ln_b, pe, drop_p, seed, salt, seed_dev, dX0, g_val, g_tim, g_feat, g_cls, g_bott, g_ln,
                           dproj, stream);
}
