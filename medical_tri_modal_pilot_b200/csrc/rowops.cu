// rowops.cu -- HBM-bound pieces of the fusion encoder:
//   * tmp_build_lengths / tmp_debug_materialize_mask : key-padding lengths (mbt_encoder.py:703-714,748;
//     tri_mbt_vsltcls.py:226-237; utils.py:79-125) kept on device as int32, never as a [B*H,T,T] bool tensor
//   * tmp_layernorm_fwd / bwd : the reference's hand-written LayerNorm (module.py:130-144: unbiased std,
//     eps added to std) with the residual add / residual gradient fused, one warp per row
//   * tmp_bottleneck_mix_fwd / bwd : modality-aware bottleneck exchange (mbt_encoder.py:764-776)
//   * tmp_colsum : bias gradients; tmp_dropout_apply; tmp_cast_weights : fp32 master -> fp16 (+transposed) copies
#include "common.cuh"
#include "rowwise.cuh"

using namespace tc05;
using namespace rw;

namespace {

// ------------------------------------------------------------------------------------------------
// lengths (a3, a4)
// ------------------------------------------------------------------------------------------------
__global__ void build_lengths_kernel(const long long* __restrict__ input_lengths, const long long* __restrict__ txt_lengths,
                                     const float* __restrict__ img_time, int n_img, int multiimages,
                                     const long long* __restrict__ missing, int skip_missing, int B, int T_v, int T_i,
                                     int T_t, int32_t* __restrict__ kv_len) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  // vslt: input_lengths + 1 (CLS, mbt_encoder.py:704) + 4 bottleneck keys (:748)
  int v = (int)input_lengths[b] + 1 + 4;
  // img: 49 * #(img_time != 10) (tri_mbt_vsltcls.py:229-232) + 1 + 4; unmasked when --multiimages 0 (:144,:234)
  int im = T_i;
  if (multiimages) {
    int cnt = 0;
    for (int k = 0; k < n_img; ++k) cnt += (img_time[b * n_img + k] - 10.0f) != 0.0f;
    im = 49 * cnt + 1 + 4;
  }
  // txt: caller passes txt_lengths + 2 (tri_mbt_vsltcls.py:237), encoder adds 1 and maps 3 -> 0 (mbt_encoder.py:704-707)
  int t = (int)txt_lengths[b] + 3;
  if (t == 3) t = 0;
  t += 4;
  if (skip_missing && missing) {
    const long long m = missing[b];  // 0 = all, 1 = txt missing, 2 = img missing, 3 = both (trainer.py:68-84)
    if (m == 2 || m == 3) im = 0;
    if (m == 1 || m == 3) t = 0;
  }
  kv_len[b] = min(max(v, 0), T_v);
  kv_len[B + b] = min(max(im, 0), T_i);
  kv_len[2 * B + b] = min(max(t, 0), T_t);
}

// mask[b,q,k] = (k >= kv_len[b])   -- test-only restatement of get_attn_pad_mask
__global__ void materialize_mask_kernel(const int32_t* __restrict__ kv_len, int B, int T, uint8_t* __restrict__ mask) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * T * T;
  if (idx >= total) return;
  const int k = (int)(idx % T);
  const int b = (int)(idx / ((size_t)T * T));
  mask[idx] = k >= kv_len[b];
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (reference module.py:130-144): y = gamma * (z - mean) / (std_unbiased + eps) + beta
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ln_stats(float (&c)[8], float& r, float& s_std) {
  float s1 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s1 += c[i];
  const float mean = warp_sum(s1) * (1.f / D);
  float s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i] -= mean; s2 += c[i] * c[i]; }
  s_std = sqrtf(warp_sum(s2) * (1.f / (D - 1)));
  r = 1.f / (s_std + 1e-6f);
}

// ADD: h = x + o written to `sum_out`, then normalised.  x,o,sum_out,y: [rows,256] fp16
template <bool ADD, int ST>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const void* __restrict__ x, const void* __restrict__ o,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, long long rows,
                                                            void* __restrict__ sum_out, void* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  float g[8], be[8];
  load8_f32(gamma + lane * 8, g);
  load8_f32(beta + lane * 8, be);
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
    float c[8];
    const size_t at = (size_t)row * D + lane * 8;
    ld8<ST>(x, at, c);
    if (ADD) {
      float a[8];
      ld8<ST>(o, at, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) c[i] += a[i];
      st8<ST>(sum_out, at, c);
      // normalise the stored (fp16-rounded) sum: it is what the backward pass and the residual path see
      ld8<ST>(sum_out, at, c);
    }
    float r, sd;
    ln_stats(c, r, sd);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(c[i] * r, g[i], be[i]);
    st8<ST>(y, at, v);
  }
}

// dx = dres + LN'(dy; x).  Optionally also writes dx_drop = dropout_mask(seed,salt) * dx / (1-p)
// (the gradient entering the previous block's FFN2 when its output dropout is active).
// ATTN (the LayerNorm in front of the FFN, whose input gradient dx IS the attention's dO): the same pass also produces
// what the attention backward needs before it can start -- delta[b,h,q] = sum_d dO.O per head (8 lanes x 8 channels = one
// head) and zeroed dQ columns of dQKV (the kernel reduce-adds its dQ tiles there) -- instead of a delta kernel (reads O and
// dO again), a memset and an fp32 -> fp16 convert pass.
struct LnAttnArgs {
  const void* O;     // [rows, 256] attention output of the forward
  float* delta;      // [B, 4, T_lse]
  void* dQKV;        // [rows, 768]: columns [0, 256) are zeroed
  int T, T_lse;
};

template <int ST, bool ATTN>
__global__ void __launch_bounds__(256, 3) layernorm_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ x,
                                                            const void* __restrict__ dres,
                                                            const float* __restrict__ gamma, long long rows,
                                                            void* __restrict__ dx, void* __restrict__ dx_drop,
                                                            uint32_t drop_thr16, float drop_scale, uint32_t seed,
                                                            uint32_t salt, const uint32_t* __restrict__ seed_dev,
                                                            float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, LnAttnArgs at_) {
  __shared__ float sAcc[2 * D];
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sAcc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float g[8], ag[8], ab[8];
  load8_f32(gamma + lane * 8, g);
#pragma unroll
  for (int i = 0; i < 8; ++i) ag[i] = ab[i] = 0.f;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  // Software-pipelined over rows: the loads of the NEXT row (x, dy, dres, O: up to 64 B per lane) are issued before the
  // current row is reduced and stored, so every warp keeps two rows of loads in flight. One row at a time left the kernel
  // latency-bound at 2.5 TB/s (54 us for 64 320 rows, profiles/r2b_step_breakdown.txt). The row in flight is held in its
  // storage form (Raw8: 4 registers per tensor instead of 8) and the launch bounds ask for 3 resident blocks (80
  // registers, 48 B of spills): 24 instead of 16 warps per SM took the cold-L2 time from 46 to 34 us (57 -> 44 us with the
  // attention extras, 4.5 TB/s); 4 blocks (64 registers) spill too much and are slower again.
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  Raw8<ST> rc, rgy, ra, ro;   // the row in flight, in storage form (4 registers per 16-bit tensor)
  auto fetch = [&](long long r) {
    const size_t at = (size_t)r * D + lane * 8;
    rc.load(x, at);
    rgy.load(dy, at);
    if (dres) ra.load(dres, at);
    if (ATTN) ro.load(at_.O, at);
  };
  if (row < rows) fetch(row);
  for (; row < rows; row += warps) {
    const size_t at = (size_t)row * D + lane * 8;
    float c[8], gy[8], a[8], o[8];
    rc.get(c);
    rgy.get(gy);
    if (dres) ra.get(a);
    if (ATTN) ro.get(o);
    const long long nxt = row + warps;
    if (nxt < rows) fetch(nxt);
    float r, sd;
    ln_stats(c, r, sd);
    float gbar = 0.f, gc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      ab[i] += gy[i];
      ag[i] += gy[i] * c[i] * r;
      gy[i] *= g[i];
      gbar += gy[i];
      gc += gy[i] * c[i];
    }
    warp_sum2(gbar, gc);
    gbar *= (1.f / D);
    // d/dc: r*g - r^2 * (sum g.c) / ((n-1) * std) * c ; then subtract the mean (only the first term has one)
    const float k2 = sd > 0.f ? r * r * gc / ((D - 1) * sd) : 0.f;
    float out[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = r * (gy[i] - gbar) - k2 * c[i];
    if (dres) {
#pragma unroll
      for (int i = 0; i < 8; ++i) out[i] += a[i];
    }
    st8<ST>(dx, at, out);
    if (ATTN) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = fmaf(out[i], o[i], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      const long long b = row / at_.T;
      const int q = (int)(row - b * at_.T);
      if ((lane & 7) == 0) at_.delta[(b * 4 + (lane >> 3)) * at_.T_lse + q] = acc;
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      st8<ST>(at_.dQKV, (size_t)row * 768 + lane * 8, z);
    }
    if (dx_drop) {
      const uint32_t base = (uint32_t)row * D + lane * 8;
      dropout_apply_run<8>(out, dropout_key(effective_seed(seed, seed_dev), salt), base, drop_thr16, drop_scale);
      st8<ST>(dx_drop, at, out);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    atomicAdd(&sAcc[lane * 8 + i], ag[i]);
    atomicAdd(&sAcc[D + lane * 8 + i], ab[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(&dgamma[i], sAcc[i]);
    atomicAdd(&dbeta[i], sAcc[D + i]);
  }
}

// ------------------------------------------------------------------------------------------------
// bottleneck exchange (a7). Y_m: [B, T_m, 256] fp16 (forward and gradients); rows 0..3 of every present stream are replaced by the
// per-sample mean over the modalities selected by `missing` (0: v,i,t  1: v,i  2: v,t  3: v).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mix_weights(long long code, float (&w)[3]) {
  switch (code) {
    case 0: w[0] = w[1] = w[2] = 1.f / 3.f; break;
    case 1: w[0] = w[1] = 0.5f; w[2] = 0.f; break;
    case 2: w[0] = w[2] = 0.5f; w[1] = 0.f; break;
    default: w[0] = 1.f; w[1] = w[2] = 0.f; break;
  }
}

template <int ST>
__global__ void bottleneck_mix_fwd_kernel(void* __restrict__ Yv, void* __restrict__ Yi, void* __restrict__ Yt, int Tv,
                                          int Ti, int Tt, const long long* __restrict__ missing, int B) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B * 4) return;
  const int b = row >> 2, r = row & 3;
  float w[3];
  mix_weights(missing[b], w);
  const size_t pv = ((size_t)b * Tv + r) * D + lane * 8;
  const size_t pi = ((size_t)b * Ti + r) * D + lane * 8;
  const size_t pt = ((size_t)b * Tt + r) * D + lane * 8;
  float acc[8], a[8];
  ld8<ST>(Yv, pv, acc);
  // sum first, scale once: the reference takes torch.mean over the selected stack (mbt_encoder.py:765-768)
  if (w[1] != 0.f) { ld8<ST>(Yi, pi, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += a[i]; }
  if (w[2] != 0.f) { ld8<ST>(Yt, pt, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += a[i]; }
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] *= w[0];
  st8<ST>(Yv, pv, acc);
  st8<ST>(Yi, pi, acc);
  st8<ST>(Yt, pt, acc);
}

// gradient: g = sum over present dY_m rows; dY_m rows <- w_m * g.  A null pointer = stream absent in the upper layer.
// dYd_m (optional): the same rows after the NEXT layer-down's output dropout (mask of (seed, salt_m, element index in the
// stream's [B*T_m, 256] matrix)): the LayerNorm backward that produced dY_m already wrote dropout(dY_m) for every row, and
// this kernel is the only thing that changes rows afterwards.
struct MixDrop {
  void* dYd[3];
  uint32_t salt[3];
  uint32_t thr16, seed;
  float scale;
  const uint32_t* seed_dev;
};
template <int ST>
__global__ void bottleneck_mix_bwd_kernel(void* __restrict__ dYv, void* __restrict__ dYi, void* __restrict__ dYt,
                                          int Tv, int Ti, int Tt, int upper_has_it,
                                          const long long* __restrict__ missing, int B, MixDrop md) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B * 4) return;
  const int b = row >> 2, r = row & 3;
  float w[3];
  mix_weights(missing[b], w);
  const size_t pv = ((size_t)b * Tv + r) * D + lane * 8;
  const size_t pi = ((size_t)b * Ti + r) * D + lane * 8;
  const size_t pt = ((size_t)b * Tt + r) * D + lane * 8;
  float g[8], a[8], o[8];
  ld8<ST>(dYv, pv, g);
  if (upper_has_it) {
    ld8<ST>(dYi, pi, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] += a[i];
    ld8<ST>(dYt, pt, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] += a[i];
  }
  void* const dst[3] = {dYv, dYi, dYt};
  const size_t pos[3] = {pv, pi, pt};
#pragma unroll
  for (int m = 0; m < 3; ++m) {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = g[i] * w[m];
    st8<ST>(dst[m], pos[m], o);
    if (md.dYd[m]) {
      dropout_apply_run<8>(o, dropout_key(effective_seed(md.seed, md.seed_dev), md.salt[m]), (uint32_t)pos[m], md.thr16,
                           md.scale);
      st8<ST>(md.dYd[m], pos[m], o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// column sums (bias gradients): out[N] += sum_rows dY[rows, N]   (fp16 gradients in, fp32 atomics out)
// block = 256 threads = (N/8 column groups) x (rows in flight)
// ------------------------------------------------------------------------------------------------
template <int ST>
__global__ void __launch_bounds__(256) colsum_kernel(const void* __restrict__ dY, int ld, long long M, int N,
                                                     long long rows_per_block, float* __restrict__ out) {
  extern __shared__ float sacc[];  // [N]
  for (int i = threadIdx.x; i < N; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int groups = N / 8;
  const int lanes_r = blockDim.x / groups;   // row lanes
  const int cg = threadIdx.x % groups, rl = threadIdx.x / groups;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (rl < lanes_r) {
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = min(M, r0 + rows_per_block);
    for (long long r = r0 + rl; r < r1; r += lanes_r) {
      float v[8];
      ld8<ST>(dY, (size_t)r * ld + cg * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&sacc[cg * 8 + i], acc[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) atomicAdd(&out[i], sacc[i]);
}

template <int ST>
__global__ void dropout_apply_kernel(const void* __restrict__ in, void* __restrict__ out, long long n8,
                                     uint32_t thr16, float scale, uint32_t seed, uint32_t salt,
                                     const uint32_t* __restrict__ seed_dev) {
  const long long i8 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i8 >= n8) return;
  float v[8];
  ld8<ST>(in, (size_t)i8 * 8, v);
  const uint32_t base = (uint32_t)(i8 * 8);
  dropout_apply_run<8>(v, dropout_key(effective_seed(seed, seed_dev), salt), base, thr16, scale);
  st8<ST>(out, (size_t)i8 * 8, v);
}

// ------------------------------------------------------------------------------------------------
// weight refresh: fp32 master [R,C] -> fp16 [R,C] and fp16 transposed [C,R], batched over a descriptor table
// ------------------------------------------------------------------------------------------------
struct CastDesc { const float* src; __half* dst; __half* dst_t; int R, C; };

__global__ void cast_weights_kernel(const CastDesc* __restrict__ descs) {
  __shared__ float tile[32][33];
  const CastDesc d = descs[blockIdx.z];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  if (c0 >= d.C || r0 >= d.R) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    float v = 0.f;
    if (r < d.R && c < d.C) {
      v = d.src[(size_t)r * d.C + c];
      if (d.dst) d.dst[(size_t)r * d.C + c] = __float2half_rn(v);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  if (d.dst_t) {
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, r = r0 + tx;
      if (r < d.R && c < d.C) d.dst_t[(size_t)c * d.R + r] = __float2half_rn(tile[tx][i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// AdamW over the flat fp32 parameter / gradient buffers (reference 2_train.py:110 torch.optim.AdamW semantics:
// decoupled weight decay, bias-corrected moments, eps added after the sqrt(v)/sqrt(bc2) division).
// 28 B/parameter of HBM traffic (read w,g,m,v; write w,m,v), one launch for all fused-path parameters.
// ------------------------------------------------------------------------------------------------
template <bool DEV>
__global__ void __launch_bounds__(256) adamw_kernel(float4* __restrict__ w, const float4* __restrict__ g,
                                                    float4* __restrict__ m, float4* __restrict__ v, long long n4,
                                                    float lr, float b1, float b2, float eps, float wd,
                                                    float inv_bc1, float inv_sqrt_bc2,
                                                    const float* __restrict__ lr_dev,
                                                    int32_t* __restrict__ step_dev, int count_skip) {
  if (DEV) {
    // learning rate / step count live in device memory (CUDA-graph replays): bias corrections per thread, in double
    // like the host path (pow of a handful of values, once per thread).
    // step_dev = {calls, skipped, last call with a non-finite gradient}: a call flagged by grad_nonfinite_kernel leaves
    // w, m, v untouched and does not count as an optimizer step (what torch.cuda.amp.GradScaler does on the host).
    const int calls = step_dev[0], skipped = step_dev[1], bad = step_dev[2];
    if (bad == calls) {
      // nobody else reads it in a skipped call; count_skip == 0: a second launch of the same optimizer step (the head
      // parameters' flat buffer) that shares the words but must not count the skip twice
      if (count_skip && blockIdx.x == 0 && threadIdx.x == 0) step_dev[1] = skipped + 1;
      return;
    }
    lr = __ldg(lr_dev);
    const double t = (double)(calls - skipped);
    inv_bc1 = (float)(1.0 / (1.0 - pow((double)b1, t)));
    inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)b2, t)));
  }
  const float decay = 1.f - lr * wd;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 W = w[i], M = m[i], V = v[i];
    const float4 G = g[i];
    float* pw = &W.x; float* pm = &M.x; float* pv = &V.x; const float* pg = &G.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      pw[k] *= decay;
      pm[k] = b1 * pm[k] + (1.f - b1) * pg[k];
      pv[k] = b2 * pv[k] + (1.f - b2) * pg[k] * pg[k];
      const float denom = sqrtf(pv[k]) * inv_sqrt_bc2 + eps;
      pw[k] -= lr * inv_bc1 * pm[k] / denom;
    }
    w[i] = W; m[i] = M; v[i] = V;
  }
}

// state[2] <- state[0] (= the optimizer call in progress) when any of g[0..n) is inf / NaN. Runs on the fully reduced
// gradient (after the data-parallel all-reduce: every rank takes the same decision).
__global__ void __launch_bounds__(256) grad_nonfinite_kernel(const float4* __restrict__ g, long long n4,
                                                             int32_t* __restrict__ state) {
  bool bad = false;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint4 u = *reinterpret_cast<const uint4*>(g + i);
    bad |= ((u.x & 0x7f800000u) == 0x7f800000u) | ((u.y & 0x7f800000u) == 0x7f800000u) |
           ((u.z & 0x7f800000u) == 0x7f800000u) | ((u.w & 0x7f800000u) == 0x7f800000u);
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicMax(&state[2], state[0]);
}

int rows_grid(long long rows) {
  long long blocks = (rows + 7) / 8;
  const long long cap = (long long)tmp::num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

extern "C" int tmp_build_lengths(const long long* input_lengths, const long long* txt_lengths, const float* img_time,
                                 int n_img, int multiimages, const long long* missing, int skip_missing, int B, int T_v,
                                 int T_i, int T_t, int32_t* kv_len, void* stream) {
  TMP_REQUIRE(input_lengths && txt_lengths && kv_len && B > 0, "build_lengths: bad argument");
  TMP_REQUIRE(!multiimages || img_time, "build_lengths: --multiimages 1 needs img_time");
  build_lengths_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      input_lengths, txt_lengths, img_time, n_img, multiimages, missing, skip_missing, B, T_v, T_i, T_t, kv_len);
  return tmp::check_launch("build_lengths_kernel");
}

extern "C" int tmp_debug_materialize_mask(const int32_t* kv_len, int B, int T, uint8_t* mask, void* stream) {
  TMP_REQUIRE(kv_len && mask && B > 0 && T > 0, "materialize_mask: bad argument");
  const size_t total = (size_t)B * T * T;
  materialize_mask_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kv_len, B, T, mask);
  return tmp::check_launch("materialize_mask_kernel");
}

static int layernorm_fwd_impl(int st, const void* x, const void* add, const float* gamma, const float* beta,
                              long long rows, void* sum_out, void* y, void* stream) {
  TMP_REQUIRE(x && gamma && beta && y && rows >= 0, "layernorm_fwd: bad argument");
  TMP_REQUIRE(!add || sum_out, "layernorm_fwd: fused add needs sum_out");
  if (rows == 0) return TMP_OK;
  const int grid = rows_grid(rows);
  cudaStream_t s = (cudaStream_t)stream;
  if (st == FMT_F32) {
    if (add) layernorm_fwd_kernel<true, FMT_F32><<<grid, 256, 0, s>>>(x, add, gamma, beta, rows, sum_out, y);
    else layernorm_fwd_kernel<false, FMT_F32><<<grid, 256, 0, s>>>(x, nullptr, gamma, beta, rows, nullptr, y);
  } else {
    if (add) layernorm_fwd_kernel<true, ACT><<<grid, 256, 0, s>>>(x, add, gamma, beta, rows, sum_out, y);
    else layernorm_fwd_kernel<false, ACT><<<grid, 256, 0, s>>>(x, nullptr, gamma, beta, rows, nullptr, y);
  }
  return tmp::check_launch("layernorm_fwd_kernel");
}
extern "C" int tmp_layernorm_fwd(const void* x, const void* add, const float* gamma, const float* beta, long long rows,
                                 void* sum_out, void* y, void* stream) {
  return layernorm_fwd_impl(ACT, x, add, gamma, beta, rows, sum_out, y, stream);
}
extern "C" int tmp_layernorm_fwd_f32(const float* x, const float* add, const float* gamma, const float* beta,
                                     long long rows, float* sum_out, float* y, void* stream) {
  return layernorm_fwd_impl(FMT_F32, x, add, gamma, beta, rows, sum_out, y, stream);
}

static int layernorm_bwd_impl(int st, const void* dy, const void* x, const void* dres, const float* gamma, long long rows,
                              void* dx, void* dx_drop, float drop_p, uint32_t seed, uint32_t salt,
                              const uint32_t* seed_dev, float* dgamma, float* dbeta, void* stream) {
  TMP_REQUIRE(dy && x && gamma && dx && dgamma && dbeta && rows >= 0, "layernorm_bwd: bad argument");
  TMP_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "layernorm_bwd: dropout p out of range");
  if (rows == 0) return TMP_OK;
  long long blocks = (rows + 7) / 8;
  if (blocks > 3LL * tmp::num_sms()) blocks = 3LL * tmp::num_sms();   // 3 resident blocks per SM (launch bounds): one wave   // 3 resident blocks per SM (launch bounds): one wave
  const uint32_t thr = drop_p > 0.f ? (uint32_t)(drop_p * 65536.f + 0.5f) : 0;
  const float scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const LnAttnArgs none{nullptr, nullptr, nullptr, 1, 1};
  if (st == FMT_F32)
    layernorm_bwd_kernel<FMT_F32, false><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
        dy, x, dres, gamma, rows, dx, thr ? dx_drop : nullptr, thr, scale, seed, salt, seed_dev, dgamma, dbeta, none);
  else
    layernorm_bwd_kernel<ACT, false><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
        dy, x, dres, gamma, rows, dx, thr ? dx_drop : nullptr, thr, scale, seed, salt, seed_dev, dgamma, dbeta, none);
  return tmp::check_launch("layernorm_bwd_kernel");
}

// LayerNorm backward in front of the attention backward (16-bit path): dx (= dO of the attention), plus delta[B,4,T_lse] and
// zeroed dQ columns of dQKV [rows,768] -- see LnAttnArgs. rows = B*T.
extern "C" int tmp_layernorm_bwd_attn(const void* dy, const void* x, const void* dres, const float* gamma, long long rows,
                                      void* dx, float* dgamma, float* dbeta, const void* attn_O, int T, int T_lse,
                                      float* delta, void* dQKV, void* stream) {
  TMP_REQUIRE(dy && x && gamma && dx && dgamma && dbeta && attn_O && delta && dQKV && rows >= 0, "layernorm_bwd_attn: bad argument");
  TMP_REQUIRE(T > 0 && T_lse >= T && rows % T == 0, "layernorm_bwd_attn: rows must be B*T and T_lse >= T");
  if (rows == 0) return TMP_OK;
  long long blocks = (rows + 7) / 8;
  if (blocks > 3LL * tmp::num_sms()) blocks = 3LL * tmp::num_sms();   // 3 resident blocks per SM (launch bounds): one wave
  const LnAttnArgs a{attn_O, delta, dQKV, T, T_lse};
  layernorm_bwd_kernel<ACT, true><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(dy, x, dres, gamma, rows, dx, nullptr, 0u, 1.f,
                                                                                0u, 0u, nullptr, dgamma, dbeta, a);
  return tmp::check_launch("layernorm_bwd_kernel");
}
extern "C" int tmp_layernorm_bwd(const void* dy, const void* x, const void* dres, const float* gamma, long long rows,
                                 void* dx, void* dx_drop, float drop_p, uint32_t seed, uint32_t salt,
                                 const uint32_t* seed_dev, float* dgamma, float* dbeta, void* stream) {
  return layernorm_bwd_impl(ACT, dy, x, dres, gamma, rows, dx, dx_drop, drop_p, seed, salt, seed_dev, dgamma, dbeta, stream);
}
extern "C" int tmp_layernorm_bwd_f32(const float* dy, const float* x, const float* dres, const float* gamma,
                                     long long rows, float* dx, float* dx_drop, float drop_p, uint32_t seed,
                                     uint32_t salt, const uint32_t* seed_dev, float* dgamma, float* dbeta,
                                     void* stream) {
  return layernorm_bwd_impl(FMT_F32, dy, x, dres, gamma, rows, dx, dx_drop, drop_p, seed, salt, seed_dev, dgamma, dbeta,
                            stream);
}

static int mix_fwd_impl(int st, void* Yv, void* Yi, void* Yt, int Tv, int Ti, int Tt, const long long* missing, int B,
                        void* stream) {
  TMP_REQUIRE(Yv && Yi && Yt && missing && B > 0 && Tv >= 4 && Ti >= 4 && Tt >= 4, "bottleneck_mix_fwd: bad argument");
  if (st == FMT_F32)
    bottleneck_mix_fwd_kernel<FMT_F32><<<(B * 4 + 7) / 8, 256, 0, (cudaStream_t)stream>>>(Yv, Yi, Yt, Tv, Ti, Tt, missing, B);
  else
    bottleneck_mix_fwd_kernel<ACT><<<(B * 4 + 7) / 8, 256, 0, (cudaStream_t)stream>>>(Yv, Yi, Yt, Tv, Ti, Tt, missing, B);
  return tmp::check_launch("bottleneck_mix_fwd_kernel");
}
extern "C" int tmp_bottleneck_mix_fwd(void* Yv, void* Yi, void* Yt, int Tv, int Ti, int Tt, const long long* missing,
                                      int B, void* stream) {
  return mix_fwd_impl(ACT, Yv, Yi, Yt, Tv, Ti, Tt, missing, B, stream);
}
extern "C" int tmp_bottleneck_mix_fwd_f32(float* Yv, float* Yi, float* Yt, int Tv, int Ti, int Tt,
                                          const long long* missing, int B, void* stream) {
  return mix_fwd_impl(FMT_F32, Yv, Yi, Yt, Tv, Ti, Tt, missing, B, stream);
}

static int mix_bwd_impl(int st, void* dYv, void* dYi, void* dYt, int Tv, int Ti, int Tt, int upper_has_img_txt,
                        const long long* missing, int B, void* dYd_v, void* dYd_i, void* dYd_t, float drop_p, uint32_t seed,
                        const uint32_t* seed_dev, uint32_t salt_v, uint32_t salt_i, uint32_t salt_t, void* stream) {
  TMP_REQUIRE(dYv && dYi && dYt && missing && B > 0, "bottleneck_mix_bwd: bad argument");
  TMP_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "bottleneck_mix_bwd: dropout p out of range");
  TMP_REQUIRE((long long)B * (Tv > Ti ? (Tv > Tt ? Tv : Tt) : (Ti > Tt ? Ti : Tt)) * D < (1ll << 32),
              "bottleneck_mix_bwd: element index exceeds 32 bits");
  MixDrop md;
  const bool drop = drop_p > 0.f;
  md.dYd[0] = drop ? dYd_v : nullptr; md.dYd[1] = drop ? dYd_i : nullptr; md.dYd[2] = drop ? dYd_t : nullptr;
  md.salt[0] = salt_v; md.salt[1] = salt_i; md.salt[2] = salt_t;
  md.thr16 = drop ? (uint32_t)(drop_p * 65536.f + 0.5f) : 0; md.seed = seed;
  md.scale = drop ? 1.f / (1.f - drop_p) : 1.f;
  md.seed_dev = seed_dev;
  if (st == FMT_F32)
    bottleneck_mix_bwd_kernel<FMT_F32><<<(B * 4 + 7) / 8, 256, 0, (cudaStream_t)stream>>>(dYv, dYi, dYt, Tv, Ti, Tt,
                                                                                        upper_has_img_txt, missing, B, md);
  else
    bottleneck_mix_bwd_kernel<GRD><<<(B * 4 + 7) / 8, 256, 0, (cudaStream_t)stream>>>(dYv, dYi, dYt, Tv, Ti, Tt,
                                                                                    upper_has_img_txt, missing, B, md);
  return tmp::check_launch("bottleneck_mix_bwd_kernel");
}
extern "C" int tmp_bottleneck_mix_bwd(void* dYv, void* dYi, void* dYt, int Tv, int Ti, int Tt, int upper_has_img_txt,
                                      const long long* missing, int B, void* dYd_v, void* dYd_i, void* dYd_t,
                                      float drop_p, uint32_t seed, const uint32_t* seed_dev, uint32_t salt_v,
                                      uint32_t salt_i, uint32_t salt_t, void* stream) {
  return mix_bwd_impl(GRD, dYv, dYi, dYt, Tv, Ti, Tt, upper_has_img_txt, missing, B, dYd_v, dYd_i, dYd_t, drop_p, seed,
                      seed_dev, salt_v, salt_i, salt_t, stream);
}
extern "C" int tmp_bottleneck_mix_bwd_f32(float* dYv, float* dYi, float* dYt, int Tv, int Ti, int Tt,
                                          int upper_has_img_txt, const long long* missing, int B, float* dYd_v,
                                          float* dYd_i, float* dYd_t, float drop_p, uint32_t seed,
                                          const uint32_t* seed_dev, uint32_t salt_v, uint32_t salt_i, uint32_t salt_t,
                                          void* stream) {
  return mix_bwd_impl(FMT_F32, dYv, dYi, dYt, Tv, Ti, Tt, upper_has_img_txt, missing, B, dYd_v, dYd_i, dYd_t, drop_p, seed,
                      seed_dev, salt_v, salt_i, salt_t, stream);
}

// out[N] += column sums of dY [M, N]. The 16-bit path gets its bias gradients from the weight-gradient kernel
// (gemm_wgrad, fused); this stand-alone form serves the fp32 mode, whose split operands cannot be summed that way.
static int colsum_impl(int st, const void* dY, int ld, long long M, int N, float* out, void* stream) {
  TMP_REQUIRE(dY && out && M >= 0 && N > 0 && N % 8 == 0 && N <= 2048 && ld % 8 == 0, "colsum: bad argument");
  if (M == 0) return TMP_OK;
  const int groups = N / 8;
  TMP_REQUIRE(groups <= 256, "colsum: N too large");
  long long blocks = 2LL * tmp::num_sms();
  long long rpb = (M + blocks - 1) / blocks;
  if (rpb < 32) rpb = 32;
  blocks = (M + rpb - 1) / rpb;
  if (st == FMT_F32)
    colsum_kernel<FMT_F32><<<(int)blocks, 256, N * sizeof(float), (cudaStream_t)stream>>>(dY, ld, M, N, rpb, out);
  else
    colsum_kernel<GRD><<<(int)blocks, 256, N * sizeof(float), (cudaStream_t)stream>>>(dY, ld, M, N, rpb, out);
  return tmp::check_launch("colsum_kernel");
}
extern "C" int tmp_colsum(const void* dY, int ld, long long M, int N, float* out, void* stream) {
  return colsum_impl(GRD, dY, ld, M, N, out, stream);
}
extern "C" int tmp_colsum_f32(const float* dY, int ld, long long M, int N, float* out, void* stream) {
  return colsum_impl(FMT_F32, dY, ld, M, N, out, stream);
}

static int dropout_apply_impl(int st, const void* in, void* out, long long n, float drop_p, uint32_t seed, uint32_t salt,
                              const uint32_t* seed_dev, void* stream) {
  TMP_REQUIRE(in && out && n >= 0 && n % 8 == 0 && drop_p >= 0.f && drop_p < 1.f, "dropout_apply: bad argument");
  if (n == 0) return TMP_OK;
  const uint32_t thr = (uint32_t)(drop_p * 65536.f + 0.5f);
  const long long n8 = n / 8;
  const unsigned grid = (unsigned)((n8 + 255) / 256);
  if (st == FMT_F32)
    dropout_apply_kernel<FMT_F32><<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, n8, thr, 1.f / (1.f - drop_p), seed, salt,
                                                                        seed_dev);
  else
    dropout_apply_kernel<GRD><<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, n8, thr, 1.f / (1.f - drop_p), seed, salt,
                                                                    seed_dev);
  return tmp::check_launch("dropout_apply_kernel");
}
extern "C" int tmp_dropout_apply(const void* in, void* out, long long n, float drop_p, uint32_t seed, uint32_t salt,
                                 const uint32_t* seed_dev, void* stream) {
  return dropout_apply_impl(GRD, in, out, n, drop_p, seed, salt, seed_dev, stream);
}
extern "C" int tmp_dropout_apply_f32(const float* in, float* out, long long n, float drop_p, uint32_t seed, uint32_t salt,
                                     const uint32_t* seed_dev, void* stream) {
  return dropout_apply_impl(FMT_F32, in, out, n, drop_p, seed, salt, seed_dev, stream);
}

// descs: device array of n_desc {const float* src; h16* dst; h16* dst_t; int R; int C}; max_R/max_C bound the grid
extern "C" int tmp_cast_weights(const void* descs, int n_desc, int max_R, int max_C, void* stream) {
  TMP_REQUIRE(descs && n_desc > 0 && max_R > 0 && max_C > 0, "cast_weights: bad argument");
  dim3 grid((max_C + 31) / 32, (max_R + 31) / 32, n_desc);
  cast_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const CastDesc*)descs);
  return tmp::check_launch("cast_weights_kernel");
}

// w, g, m, v: fp32 [n], n % 4 == 0, 16-byte aligned. step >= 1 (bias correction).
static int adamw_grid(long long n4) {
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)tmp::num_sms() * 8;
  if (blocks > cap) blocks = cap;
  return (int)blocks;
}

extern "C" int tmp_adamw_step(float* w, const float* g, float* m, float* v, long long n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int step, void* stream) {
  TMP_REQUIRE(w && g && m && v && n >= 0 && n % 4 == 0 && step >= 1, "adamw_step: bad argument");
  if (n == 0) return TMP_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const long long n4 = n / 4;
  adamw_kernel<false><<<adamw_grid(n4), 256, 0, (cudaStream_t)stream>>>(
      (float4*)w, (const float4*)g, (float4*)m, (float4*)v, n4, lr, beta1, beta2, eps, weight_decay, (float)(1.0 / bc1),
      (float)(1.0 / sqrt(bc2)), nullptr, nullptr, 0);
  return tmp::check_launch("adamw_kernel");
}

extern "C" int tmp_grad_nonfinite(const float* g, long long n, int32_t* state, void* stream) {
  TMP_REQUIRE(g && state && n >= 0 && n % 4 == 0, "grad_nonfinite: bad argument");
  if (n == 0) return TMP_OK;
  grad_nonfinite_kernel<<<adamw_grid(n / 4), 256, 0, (cudaStream_t)stream>>>((const float4*)g, n / 4, state);
  return tmp::check_launch("grad_nonfinite_kernel");
}

extern "C" int tmp_adamw_step_dev(float* w, const float* g, float* m, float* v, long long n, const float* lr_dev,
                                  float beta1, float beta2, float eps, float weight_decay, int32_t* step_dev,
                                  int count_skip, void* stream) {
  TMP_REQUIRE(w && g && m && v && lr_dev && step_dev && n >= 0 && n % 4 == 0, "adamw_step_dev: bad argument");
  if (n == 0) return TMP_OK;
  const long long n4 = n / 4;
  adamw_kernel<true><<<adamw_grid(n4), 256, 0, (cudaStream_t)stream>>>(
      (float4*)w, (const float4*)g, (float4*)m, (float4*)v, n4, 0.f, beta1, beta2, eps, weight_decay, 1.f, 1.f, lr_dev,
      step_dev, count_skip);
  return tmp::check_launch("adamw_kernel");
}
