// precise.cu -- the fp32 ("precise") mode of the hot path: north-star's FP32/TF32 parity mode (logits within 1e-3 of the
// reference, every parameter gradient cosine >= 0.999 END TO END, i.e. also under the head's BatchNorm-centred upstream
// gradient where any 16-bit storage plan -- the reference's own fp16 autocast included -- is ill-conditioned).
// Every activation / gradient tensor is stored in fp32 (rowops.cu / embed.cu are templated on the storage format);
// this file adds the two pieces that need their own kernels:
//
//   * bf16x3 operand splitting for the tensor-core GEMMs. x = hi + mid + lo with three bf16 terms (8 + 8 + 8 mantissa
//     bits, full fp32 exponent range, so no scaling / underflow hazard as an fp16 split would have). A product of two
//     split numbers keeps the six terms of order <= 2^-16:  hi*hi + hi*mid + mid*hi + hi*lo + lo*hi + mid*mid.
//     Instead of a new GEMM kernel, the six terms are laid out ALONG THE REDUCTION DIMENSION: A [M,K] fp32 becomes
//     A6 [M,6K] bf16 = [hi|hi|mid|hi|lo|mid] and B [N,K] becomes B6 [N,6K] = [hi|mid|hi|lo|hi|mid], and the existing
//     tcgen05 kernel (gemm_tn, fp32 accumulation in TMEM) computes A6 . B6^T = A . B^T to ~2^-17 relative. For the weight
//     gradient (reduction over tokens) the same six blocks are stacked along the rows: dY6 [6M,N], X6 [6M,K].
//   * attention forward / backward on the CUDA cores in fp32 (reference attention.py:24-49: S = QK^T/8, key-padding mask
//     by kv_len, softmax, PV; no output projection). One warp per query row (forward, dQ) or per key row (dK, dV), one
//     lane per key (query) of a 32-wide tile staged in shared memory: no atomics, deterministic.
// This mode is for parity, not speed (~10x the 16-bit step).
#include "common.cuh"
#include "rowwise.cuh"

using namespace tc05;

namespace {

// ------------------------------------------------------------------------------------------------
// bf16x3 split
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split3(float x, uint16_t& hi, uint16_t& mid, uint16_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(h);          // exact in fp32
  const __nv_bfloat16 m = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(m);         // exact
  const __nv_bfloat16 l = __float2bfloat16_rn(r2);
  hi = *reinterpret_cast<const uint16_t*>(&h);
  mid = *reinterpret_cast<const uint16_t*>(&m);
  lo = *reinterpret_cast<const uint16_t*>(&l);
}

// which term (0 = hi, 1 = mid, 2 = lo) goes into block j of the six, for the A-side and the B-side operand
__device__ __constant__ int kPatA[6] = {0, 0, 1, 0, 2, 1};
__device__ __constant__ int kPatB[6] = {0, 1, 0, 2, 0, 1};

// src [R, C] fp32 (row stride ld) -> dst bf16. stack_rows == 0: dst [R, 6C], block j at columns [jC, (j+1)C);
// stack_rows == 1: dst [6R, C], block j at rows [jR, (j+1)R). One thread = 4 consecutive columns.
__global__ void __launch_bounds__(256) split_bf16x3_kernel(const float* __restrict__ src, long long ld, long long R, int C,
                                                          int side_b, int stack_rows, uint16_t* __restrict__ dst) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4 = C / 4;
  if (idx >= R * c4) return;
  const long long r = idx / c4;
  const int c = (int)(idx % c4) * 4;
  const float4 v = *reinterpret_cast<const float4*>(src + r * ld + c);
  uint16_t t[3][4];
  split3(v.x, t[0][0], t[1][0], t[2][0]);
  split3(v.y, t[0][1], t[1][1], t[2][1]);
  split3(v.z, t[0][2], t[1][2], t[2][2]);
  split3(v.w, t[0][3], t[1][3], t[2][3]);
  const int* pat = side_b ? kPatB : kPatA;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int k = pat[j];
    const uint2 w = make_uint2((uint32_t)t[k][0] | ((uint32_t)t[k][1] << 16), (uint32_t)t[k][2] | ((uint32_t)t[k][3] << 16));
    uint16_t* d = stack_rows ? dst + ((long long)j * R + r) * C + c : dst + r * (6LL * C) + (long long)j * C + c;
    *reinterpret_cast<uint2*>(d) = w;
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 attention on the CUDA cores. qkv [B*T, 768] fp32 (Q | K | V, head h at columns h*64), kv_len [B] or null.
// ------------------------------------------------------------------------------------------------
constexpr int HD = 64;
constexpr int TK = 32;           // keys (queries) per shared-memory tile = one per lane
constexpr int kWarps = 8;        // rows per block
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// stage rows [r0, r0+32) of one 64-wide column block of a [*, ld] fp32 matrix into s[32][65]; rows >= r_end are zero
__device__ __forceinline__ void stage_tile(float (*s)[HD + 1], const float* __restrict__ base, long long ld, int r0, int r_end) {
  for (int i = threadIdx.x; i < TK * (HD / 4); i += blockDim.x) {
    const int r = i / (HD / 4), c = (i % (HD / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < r_end) v = *reinterpret_cast<const float4*>(base + (long long)(r0 + r) * ld + c);
    s[r][c] = v.x; s[r][c + 1] = v.y; s[r][c + 2] = v.z; s[r][c + 3] = v.w;
  }
}

// O[b, q, h*64..] = softmax_k(Q K^T / 8, k < len) V ; lse2 = log2-domain logsumexp (same convention as attn_fwd_tc05.cu)
__global__ void __launch_bounds__(kWarps * 32) attn_fwd_f32_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ kv_len,
                                                                  int T, int H, float* __restrict__ O, int ld_o,
                                                                  float* __restrict__ lse2, int T_lse) {
  __shared__ float sK[TK][HD + 1];
  __shared__ float sV[TK][HD + 1];
  __shared__ float sQ[kWarps][HD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q = blockIdx.x * kWarps + warp;
  const int len = kv_len ? min(kv_len[b], T) : T;
  const long long row_base = (long long)b * T;
  const bool live = q < len;                   // pad query rows: zeros (never consumed)
  if (q < T) {
    const float* qr = qkv + (row_base + q) * 768 + h * HD;
    sQ[warp][lane] = qr[lane];
    sQ[warp][lane + 32] = qr[lane + 32];
  }
  float m = -INFINITY, l = 0.f;
  float acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  const float c = kLog2e / 8.0f;
  const int n_tiles = (len + TK - 1) / TK;     // uniform per block (same sample)
  for (int t = 0; t < n_tiles; ++t) {
    __syncthreads();
    stage_tile(sK, qkv + row_base * 768 + 256 + h * HD, 768, t * TK, len);
    stage_tile(sV, qkv + row_base * 768 + 512 + h * HD, 768, t * TK, len);
    __syncthreads();
    if (!live) continue;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) s = fmaf(sQ[warp][d], sK[lane][d], s);
    s = (t * TK + lane < len) ? s * c : -INFINITY;
    const float m_new = fmaxf(m, warp_max(s));   // finite: the tile holds at least one valid key
    if (m_new > m) {
      const float a = exp2f(m - m_new);          // 0 on the first tile
      l *= a;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] *= a;
      m = m_new;
    }
    const float p = exp2f(s - m);
    l += p;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = fmaf(p, sV[lane][d], acc[d]);
  }
  if (q >= T) return;
  float* o = O + (row_base + q) * ld_o + h * HD;
  if (!live) {
    o[lane] = 0.f; o[lane + 32] = 0.f;
    if (lane == 0) lse2[((long long)b * H + h) * T_lse + q] = 0.f;
    return;
  }
  l = rw::warp_sum(l);
  const float inv = 1.f / l;
  float o0 = 0.f, o1 = 0.f;
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    const float v = rw::warp_sum(acc[d]);
    if (d == lane) o0 = v;
    if (d == lane + 32) o1 = v;
  }
  o[lane] = o0 * inv;
  o[lane + 32] = o1 * inv;
  if (lane == 0) lse2[((long long)b * H + h) * T_lse + q] = m + log2f(l);
}

// dQ (one warp per query row) + delta[b,h,q] = sum_d dO.O
__global__ void __launch_bounds__(kWarps * 32) attn_bwd_dq_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ Of,
                                                                     const float* __restrict__ dO, int ld_o,
                                                                     const int32_t* __restrict__ kv_len, int T, int H,
                                                                     const float* __restrict__ lse2, float* __restrict__ delta,
                                                                     int T_lse, float* __restrict__ dQKV) {
  __shared__ float sK[TK][HD + 1];
  __shared__ float sV[TK][HD + 1];
  __shared__ float sQ[kWarps][HD];
  __shared__ float sDO[kWarps][HD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q = blockIdx.x * kWarps + warp;
  const int len = kv_len ? min(kv_len[b], T) : T;
  const long long row_base = (long long)b * T;
  const bool live = q < len;
  float dl = 0.f, ls = 0.f;
  if (q < T) {
    const float* qr = qkv + (row_base + q) * 768 + h * HD;
    const float* gr = dO + (row_base + q) * ld_o + h * HD;
    const float* orow = Of + (row_base + q) * ld_o + h * HD;
    const float g0 = gr[lane], g1 = gr[lane + 32];
    sQ[warp][lane] = qr[lane]; sQ[warp][lane + 32] = qr[lane + 32];
    sDO[warp][lane] = g0; sDO[warp][lane + 32] = g1;
    dl = rw::warp_sum(g0 * orow[lane] + g1 * orow[lane + 32]);
    ls = lse2[((long long)b * H + h) * T_lse + q];
    if (lane == 0) delta[((long long)b * H + h) * T_lse + q] = live ? dl : 0.f;
  }
  float acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  const float c = kLog2e / 8.0f;
  const int n_tiles = (len + TK - 1) / TK;
  for (int t = 0; t < n_tiles; ++t) {
    __syncthreads();
    stage_tile(sK, qkv + row_base * 768 + 256 + h * HD, 768, t * TK, len);
    stage_tile(sV, qkv + row_base * 768 + 512 + h * HD, 768, t * TK, len);
    __syncthreads();
    if (!live) continue;
    float s = 0.f, dp = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      s = fmaf(sQ[warp][d], sK[lane][d], s);
      dp = fmaf(sDO[warp][d], sV[lane][d], dp);
    }
    const float p = (t * TK + lane < len) ? exp2f(s * c - ls) : 0.f;
    const float ds = p * (dp - dl) * 0.125f;     // dS / sqrt(d)
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = fmaf(ds, sK[lane][d], acc[d]);
  }
  if (q >= T) return;
  float o0 = 0.f, o1 = 0.f;
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    const float v = rw::warp_sum(acc[d]);
    if (d == lane) o0 = v;
    if (d == lane + 32) o1 = v;
  }
  float* dq = dQKV + (row_base + q) * 768 + h * HD;
  dq[lane] = live ? o0 : 0.f;
  dq[lane + 32] = live ? o1 : 0.f;
}

// dK, dV (one warp per key row; one lane per query of a 32-query tile)
__global__ void __launch_bounds__(kWarps * 32) attn_bwd_dkv_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ dO,
                                                                      int ld_o, const int32_t* __restrict__ kv_len, int T,
                                                                      int H, const float* __restrict__ lse2,
                                                                      const float* __restrict__ delta, int T_lse,
                                                                      float* __restrict__ dQKV) {
  __shared__ float sQ[TK][HD + 1];
  __shared__ float sG[TK][HD + 1];
  __shared__ float sKr[kWarps][HD];
  __shared__ float sVr[kWarps][HD];
  __shared__ float sL[TK], sD[TK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int k = blockIdx.x * kWarps + warp;
  const int len = kv_len ? min(kv_len[b], T) : T;
  const long long row_base = (long long)b * T;
  const bool live = k < len;                   // masked / padding keys: dK = dV = 0
  if (k < T) {
    const float* kr = qkv + (row_base + k) * 768 + 256 + h * HD;
    const float* vr = qkv + (row_base + k) * 768 + 512 + h * HD;
    sKr[warp][lane] = kr[lane]; sKr[warp][lane + 32] = kr[lane + 32];
    sVr[warp][lane] = vr[lane]; sVr[warp][lane + 32] = vr[lane + 32];
  }
  float ak[HD], av[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) ak[d] = av[d] = 0.f;
  const float c = kLog2e / 8.0f;
  const int n_tiles = (len + TK - 1) / TK;     // live query rows only (rows >= len are padding: dO == 0)
  const long long stat = ((long long)b * H + h) * T_lse;
  for (int t = 0; t < n_tiles; ++t) {
    __syncthreads();
    stage_tile(sQ, qkv + row_base * 768 + h * HD, 768, t * TK, len);
    stage_tile(sG, dO + row_base * ld_o + h * HD, ld_o, t * TK, len);
    if (threadIdx.x < TK) {
      const int qi = t * TK + threadIdx.x;
      sL[threadIdx.x] = qi < len ? lse2[stat + qi] : 0.f;
      sD[threadIdx.x] = qi < len ? delta[stat + qi] : 0.f;
    }
    __syncthreads();
    if (!live) continue;
    float s = 0.f, dp = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      s = fmaf(sQ[lane][d], sKr[warp][d], s);
      dp = fmaf(sG[lane][d], sVr[warp][d], dp);
    }
    const float p = (t * TK + lane < len) ? exp2f(s * c - sL[lane]) : 0.f;
    const float ds = p * (dp - sD[lane]) * 0.125f;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      av[d] = fmaf(p, sG[lane][d], av[d]);
      ak[d] = fmaf(ds, sQ[lane][d], ak[d]);
    }
  }
  if (k >= T) return;
  float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    const float a = rw::warp_sum(ak[d]);
    const float v = rw::warp_sum(av[d]);
    if (d == lane) { k0 = a; v0 = v; }
    if (d == lane + 32) { k1 = a; v1 = v; }
  }
  float* dk = dQKV + (row_base + k) * 768 + 256 + h * HD;
  float* dv = dQKV + (row_base + k) * 768 + 512 + h * HD;
  dk[lane] = live ? k0 : 0.f; dk[lane + 32] = live ? k1 : 0.f;
  dv[lane] = live ? v0 : 0.f; dv[lane + 32] = live ? v1 : 0.f;
}

}  // namespace

// src [R, C] fp32 with row stride ld (elements) -> dst bf16: [R, 6C] (stack_rows = 0: K-side concatenation for gemm_tn) or
// [6R, C] (stack_rows = 1: token-side stacking for gemm_wgrad). side_b selects the B-operand term order.
extern "C" int tmp_split_bf16x3(const float* src, long long ld, long long R, int C, int side_b, int stack_rows, void* dst,
                                void* stream) {
  TMP_REQUIRE(src && dst && R > 0 && C > 0 && C % 4 == 0 && ld % 4 == 0 && ld >= C, "split_bf16x3: bad argument");
  const long long n = R * (C / 4);
  split_bf16x3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, ld, R, C, side_b, stack_rows,
                                                                                      (uint16_t*)dst);
  return tmp::check_launch("split_bf16x3_kernel");
}

extern "C" int tmp_attn_fwd_f32(const float* qkv, const int32_t* kv_len, int B, int T, int H, float* O, int ld_o,
                                float* lse2, int T_lse, void* stream) {
  TMP_REQUIRE(qkv && O && lse2 && B > 0 && T > 0 && H == 4 && T_lse >= T && ld_o % 4 == 0, "attn_fwd_f32: bad argument");
  dim3 grid((T + kWarps - 1) / kWarps, H, B);
  attn_fwd_f32_kernel<<<grid, kWarps * 32, 0, (cudaStream_t)stream>>>(qkv, kv_len, T, H, O, ld_o, lse2, T_lse);
  return tmp::check_launch("attn_fwd_f32_kernel");
}

// dQKV [B*T, 768] fp32 receives dQ | dK | dV; delta [B, H, T_lse] is a workspace
extern "C" int tmp_attn_bwd_f32(const float* qkv, const float* O, const float* dO, int ld_o, const int32_t* kv_len, int B,
                                int T, int H, const float* lse2, int T_lse, float* delta, float* dQKV, void* stream) {
  TMP_REQUIRE(qkv && O && dO && lse2 && delta && dQKV && B > 0 && T > 0 && H == 4 && T_lse >= T && ld_o % 4 == 0,
              "attn_bwd_f32: bad argument");
  dim3 grid((T + kWarps - 1) / kWarps, H, B);
  attn_bwd_dq_f32_kernel<<<grid, kWarps * 32, 0, (cudaStream_t)stream>>>(qkv, O, dO, ld_o, kv_len, T, H, lse2, delta, T_lse,
                                                                         dQKV);
  int rc = tmp::check_launch("attn_bwd_dq_f32_kernel");
  if (rc) return rc;
  attn_bwd_dkv_f32_kernel<<<grid, kWarps * 32, 0, (cudaStream_t)stream>>>(qkv, dO, ld_o, kv_len, T, H, lse2, delta, T_lse,
                                                                          dQKV);
  return tmp::check_launch("attn_bwd_dkv_f32_kernel");
}
