// swin.cu -- image-encoder feed: the memory-bound glue of the frozen Swin-T forward
// (reference builder/models/src/swin_transformer.py, a patched torchvision copy; called under torch.no_grad from
// tri_mbt_vsltcls.py:205-209). The matmuls (QKV / proj / MLP / patch-merging reduction) run on the tcgen05 GEMM of
// gemm_tc05.cu; this file holds everything between them, fused so that every activation is read and written once:
//   * tmp_swin_patch_embed_ln    : Conv2d(1,96,k=4,s=4) + permute + LayerNorm(96)            (swin_transformer.py:541-551)
//   * tmp_swin_ln_window         : LayerNorm + cyclic shift (torch.roll) + window partition    (:447-449, :140-155)
//   * tmp_swin_window_attn       : per (window, head) softmax(q k^T / sqrt(32) + rel-pos bias + shift mask) v   (:157-199)
//   * tmp_swin_unwindow_add_ln   : window reverse + reverse shift + residual add + the block's second LayerNorm (:205-214, :448-449)
//   * tmp_swin_merge_ln          : PatchMerging gather (x0|x1|x2|x3) + LayerNorm(4C)            (:34-46, :75-86)
// Activations are fp16 [tokens, Cp] with Cp = channel count padded to a GEMM-friendly multiple of 128 (96 -> 128,
// 192 -> 256); pad channels are kept exactly zero. One warp per token, lanes own 8-channel (16 B) chunks.
#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int WS = 7;            // window size (swin_t: window_size=[7,7])
constexpr int WT = WS * WS;      // 49 tokens per window
constexpr int HDIM = 32;         // head dim of every Swin-T stage (C / heads)

__device__ __forceinline__ void unpack8(const uint4 u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[t]));
    v[2 * t] = f.x;
    v[2 * t + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]), pack_f16x2(v[6], v[7]));
}

// ---- token-parallel glue: a token (C channels) is spread over LPT lanes, 8-channel (16 B) chunks interleaved
// (chunk c = sub + LPT*i), so every lane is busy at every stage width and a warp handles 32/LPT tokens at once:
//   LayerNorm(C)  : LPT = C/24 (4, 8, 16, 32 lanes for C = 96, 192, 384, 768), 3 chunks per lane
//   LayerNorm(4C) : LPT = C/12 (8, 16, 32 lanes for C = 96, 192, 384),         6 chunks per lane
// Stage geometry (C, H) is a template parameter: all index arithmetic divides by constants.
template <int LPT>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPT / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// nn.LayerNorm(eps=1e-5) of one token held as NI chunks per lane over LPT lanes; gamma / beta come from shared memory.
template <int NI, int LPT>
__device__ __forceinline__ void group_layernorm(float (&v)[NI][8], int sub, const float* __restrict__ sG,
                                                const float* __restrict__ sB) {
  constexpr float inv_c = 1.f / (float)(NI * LPT * 8);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[i][k];
  const float mean = group_sum<LPT>(s) * inv_c;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[i][k] -= mean;
      q = fmaf(v[i][k], v[i][k], q);
    }
  const float rstd = rsqrtf(group_sum<LPT>(q) * inv_c + 1e-5f);
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int c = sub + LPT * i;
    const float4 g0 = *reinterpret_cast<const float4*>(sG + c * 8), g1 = *reinterpret_cast<const float4*>(sG + c * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(sB + c * 8), b1 = *reinterpret_cast<const float4*>(sB + c * 8 + 4);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) v[i][k] = fmaf(v[i][k] * rstd, gg[k], bb[k]);
  }
}

// Dead-image skipping: `live` (uint8 per image, or null) marks the images whose features have a consumer --
// empty slots (img_time == 10, tri_mbt_vsltcls.py:229-231) and the images of img-missing samples are masked keys of a
// de-selected stream, so their encoder work is skipped. Images are independent in every kernel of the encoder; a warp skips
// its tokens when every image they belong to is dead. Skipped rows keep stale (finite or not) data; the LAST kernel of the
// encoder writes zeros for dead images (`zero_dead`), because a masked key still meets 0 * V in the attention's P.V product.
__device__ __forceinline__ bool all_dead(const uint8_t* __restrict__ live, int first_img, int last_img) {
  if (!live) return false;
  for (int n = first_img; n <= last_img; ++n)
    if (__ldg(live + n)) return false;
  return true;
}

// ------------------------------------------------------------------------------------------------
// patch embedding: img fp32 [N,224,224] -> tokens [N*56*56, Cp] fp16 = LN(conv4x4s4(img)).
// 4 lanes per token: lane `sub` fetches row `sub` of the 4x4 patch (one float4), the group exchanges rows by shuffle,
// each lane produces 24 of the 96 channels (weights from shared memory, 16 B loads).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patch_embed_ln_kernel(const float* __restrict__ img, int n_tok,
                                                             const float* __restrict__ Wt /*[16][96]*/,
                                                             const float* __restrict__ bconv, const float* __restrict__ g,
                                                             const float* __restrict__ b, __half* __restrict__ out, int Cp,
                                                             const uint8_t* __restrict__ live) {
  constexpr int LPT = 4, TPW = 8;
  __shared__ __align__(16) float sW[16 * 96];
  __shared__ __align__(16) float sBc[96], sG[96], sBe[96];
  for (int i = threadIdx.x; i < 16 * 96; i += blockDim.x) sW[i] = Wt[i];
  for (int i = threadIdx.x; i < 96; i += blockDim.x) { sBc[i] = bconv[i]; sG[i] = g[i]; sBe[i] = b[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane & (LPT - 1), grp = lane / LPT;
  const int stride = gridDim.x * (blockDim.x >> 5) * TPW;
  for (int base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * TPW; base < n_tok; base += stride) {
    if (all_dead(live, base / 3136, min(base + TPW - 1, n_tok - 1) / 3136)) continue;
    const bool valid = base + grp < n_tok;
    const int tok = valid ? base + grp : n_tok - 1;
    const int n = tok / 3136, r = tok % 3136, ty = r / 56, tx = r % 56;
    const float4 mine = __ldg(reinterpret_cast<const float4*>(img + ((size_t)n * 224 + ty * 4 + sub) * 224 + tx * 4));
    float pix[16];    // row-major p = dy*4 + dx: the Conv2d weight order [out,1,4,4]
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int src = (lane & ~(LPT - 1)) + d;
      pix[d * 4 + 0] = __shfl_sync(0xffffffffu, mine.x, src);
      pix[d * 4 + 1] = __shfl_sync(0xffffffffu, mine.y, src);
      pix[d * 4 + 2] = __shfl_sync(0xffffffffu, mine.z, src);
      pix[d * 4 + 3] = __shfl_sync(0xffffffffu, mine.w, src);
    }
    float v[3][8];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = sub + LPT * i;
      const float4 b0 = *reinterpret_cast<const float4*>(sBc + c * 8), b1 = *reinterpret_cast<const float4*>(sBc + c * 8 + 4);
      v[i][0] = b0.x; v[i][1] = b0.y; v[i][2] = b0.z; v[i][3] = b0.w;
      v[i][4] = b1.x; v[i][5] = b1.y; v[i][6] = b1.z; v[i][7] = b1.w;
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        const float4 w0 = *reinterpret_cast<const float4*>(sW + p * 96 + c * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(sW + p * 96 + c * 8 + 4);
        v[i][0] = fmaf(pix[p], w0.x, v[i][0]); v[i][1] = fmaf(pix[p], w0.y, v[i][1]);
        v[i][2] = fmaf(pix[p], w0.z, v[i][2]); v[i][3] = fmaf(pix[p], w0.w, v[i][3]);
        v[i][4] = fmaf(pix[p], w1.x, v[i][4]); v[i][5] = fmaf(pix[p], w1.y, v[i][5]);
        v[i][6] = fmaf(pix[p], w1.z, v[i][6]); v[i][7] = fmaf(pix[p], w1.w, v[i][7]);
      }
    }
    group_layernorm<3, LPT>(v, sub, sG, sBe);
    if (valid) {
      __half* o = out + (size_t)tok * Cp;
#pragma unroll
      for (int i = 0; i < 3; ++i) *reinterpret_cast<uint4*>(o + (sub + LPT * i) * 8) = pack8(v[i]);
      for (int c = 12 + sub; c < (Cp >> 3); c += LPT) *reinterpret_cast<uint4*>(o + c * 8) = make_uint4(0, 0, 0, 0);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm + cyclic shift + window partition:  out[window_row] = LN(x[natural row])
// ------------------------------------------------------------------------------------------------
template <int C, int H>
__global__ void __launch_bounds__(256) ln_window_kernel(const __half* __restrict__ x, const float* __restrict__ g,
                                                        const float* __restrict__ b, int n_img, int Cp, int shift,
                                                        __half* __restrict__ out, const uint8_t* __restrict__ live,
                                                        int zero_dead) {
  constexpr int W = H, LPT = C / 24, TPW = 32 / LPT, PER = H * W, NWW = W / WS;
  __shared__ __align__(16) float sG[C], sB[C];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sG[i] = g[i]; sB[i] = b[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane & (LPT - 1), grp = lane / LPT;
  const int n_tok = n_img * PER;
  const int stride = gridDim.x * (blockDim.x >> 5) * TPW;
  // iterate over OUTPUT rows (window order): the writes of a warp / of neighbouring warps are contiguous
  for (int base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * TPW; base < n_tok; base += stride) {
    const bool valid = base + grp < n_tok;
    const int tok = valid ? base + grp : n_tok - 1;
    if (all_dead(live, base / PER, min(base + TPW - 1, n_tok - 1) / PER)) {
      if (zero_dead && valid) {
        __half* dst = out + (size_t)tok * Cp;
        for (int c = sub; c < (Cp >> 3); c += LPT) *reinterpret_cast<uint4*>(dst + c * 8) = make_uint4(0, 0, 0, 0);
      }
      continue;
    }
    const int n = tok / PER, rw = tok % PER;
    const int win = rw / WT, t = rw % WT;
    const int ys = (win / NWW) * WS + t / WS, xs = (win % NWW) * WS + t % WS;     // position in the shifted map
    int y = ys + shift, xx = xs + shift;                                          // source position (roll by -shift)
    if (y >= H) y -= H;
    if (xx >= W) xx -= W;
    const __half* src = x + ((size_t)n * PER + (size_t)y * W + xx) * Cp;
    float v[3][8];
#pragma unroll
    for (int i = 0; i < 3; ++i) unpack8(__ldg(reinterpret_cast<const uint4*>(src + (sub + LPT * i) * 8)), v[i]);
    group_layernorm<3, LPT>(v, sub, sG, sB);
    if (valid) {
      __half* dst = out + (size_t)tok * Cp;
      const bool dead = zero_dead && live && !__ldg(live + n);     // a dead image sharing the warp with a live one
#pragma unroll
      for (int i = 0; i < 3; ++i)
        *reinterpret_cast<uint4*>(dst + (sub + LPT * i) * 8) = dead ? make_uint4(0, 0, 0, 0) : pack8(v[i]);
      for (int c = C / 8 + sub; c < (Cp >> 3); c += LPT) *reinterpret_cast<uint4*>(dst + c * 8) = make_uint4(0, 0, 0, 0);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// window reverse + reverse shift + residual add + LayerNorm:
//   x[natural] += y[window_row];  hn[natural] = LN(x[natural])      (hn may be null: residual add only)
// ------------------------------------------------------------------------------------------------
template <int C, int H>
__global__ void __launch_bounds__(256) unwindow_add_ln_kernel(const __half* __restrict__ y, __half* __restrict__ x,
                                                              const float* __restrict__ g, const float* __restrict__ b,
                                                              int n_img, int Cp, int shift, __half* __restrict__ hn,
                                                              const uint8_t* __restrict__ live) {
  constexpr int W = H, LPT = C / 24, TPW = 32 / LPT, PER = H * W, NWW = W / WS, NWH = H / WS;
  __shared__ __align__(16) float sG[C], sB[C];
  if (hn)
    for (int i = threadIdx.x; i < C; i += blockDim.x) { sG[i] = g[i]; sB[i] = b[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane & (LPT - 1), grp = lane / LPT;
  const int n_tok = n_img * PER;
  const int stride = gridDim.x * (blockDim.x >> 5) * TPW;
  for (int base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * TPW; base < n_tok; base += stride) {
    if (all_dead(live, base / PER, min(base + TPW - 1, n_tok - 1) / PER)) continue;
    const bool valid = base + grp < n_tok;
    const int tok = valid ? base + grp : n_tok - 1;
    const int n = tok / PER, r = tok % PER, yy = r / W, xx = r % W;
    int ys = yy - shift, xs = xx - shift;                                         // where this token sits after roll(-shift)
    if (ys < 0) ys += H;
    if (xs < 0) xs += W;
    const size_t wrow = ((size_t)(n * NWH + ys / WS) * NWW + xs / WS) * WT + (ys % WS) * WS + xs % WS;
    const __half* ysrc = y + wrow * Cp;
    __half* xr = x + (size_t)tok * Cp;
    float v[3][8];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = sub + LPT * i;
      float a[8];
      unpack8(*reinterpret_cast<const uint4*>(xr + c * 8), v[i]);
      unpack8(__ldg(reinterpret_cast<const uint4*>(ysrc + c * 8)), a);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[i][k] += a[k];
      const uint4 packed = pack8(v[i]);
      if (valid) *reinterpret_cast<uint4*>(xr + c * 8) = packed;
      unpack8(packed, v[i]);            // normalise what the residual stream actually holds
    }
    if (hn) {
      group_layernorm<3, LPT>(v, sub, sG, sB);
      if (valid) {
        __half* dst = hn + (size_t)tok * Cp;
#pragma unroll
        for (int i = 0; i < 3; ++i) *reinterpret_cast<uint4*>(dst + (sub + LPT * i) * 8) = pack8(v[i]);
        for (int c = C / 8 + sub; c < (Cp >> 3); c += LPT) *reinterpret_cast<uint4*>(dst + c * 8) = make_uint4(0, 0, 0, 0);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// patch merging gather + LayerNorm(4C): out[(n,i,j), :] = LN([x(2i,2j) | x(2i+1,2j) | x(2i,2j+1) | x(2i+1,2j+1)])
// ------------------------------------------------------------------------------------------------
template <int C, int H>
__global__ void __launch_bounds__(256) merge_ln_kernel(const __half* __restrict__ x, const float* __restrict__ g,
                                                       const float* __restrict__ b, int n_img, int Cp,
                                                       __half* __restrict__ out, const uint8_t* __restrict__ live) {
  constexpr int W = H, H2 = H / 2, W2 = W / 2, C4 = 4 * C, LPT = C / 12, TPW = 32 / LPT, CCH = C / 8;
  __shared__ __align__(16) float sG[C4], sB[C4];
  for (int i = threadIdx.x; i < C4; i += blockDim.x) { sG[i] = g[i]; sB[i] = b[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane & (LPT - 1), grp = lane / LPT;
  int off[6];      // element offset of this lane's chunk i relative to the token's top-left source pixel
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c = sub + LPT * i, part = c / CCH, cc = c % CCH;   // part 0..3 = x0 (0,0), x1 (1,0), x2 (0,1), x3 (1,1)
    off[i] = ((part & 1) * W + (part >> 1)) * Cp + cc * 8;
  }
  const int n_tok = n_img * H2 * W2;
  const int stride = gridDim.x * (blockDim.x >> 5) * TPW;
  for (int base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * TPW; base < n_tok; base += stride) {
    if (all_dead(live, base / (H2 * W2), min(base + TPW - 1, n_tok - 1) / (H2 * W2))) continue;
    const bool valid = base + grp < n_tok;
    const int tok = valid ? base + grp : n_tok - 1;
    const int n = tok / (H2 * W2), r = tok % (H2 * W2), i2 = r / W2, j2 = r % W2;
    const __half* src = x + (((size_t)n * H + 2 * i2) * W + 2 * j2) * Cp;
    float v[6][8];
#pragma unroll
    for (int i = 0; i < 6; ++i) unpack8(__ldg(reinterpret_cast<const uint4*>(src + off[i])), v[i]);
    group_layernorm<6, LPT>(v, sub, sG, sB);
    if (valid) {
      __half* dst = out + (size_t)tok * C4;
#pragma unroll
      for (int i = 0; i < 6; ++i) *reinterpret_cast<uint4*>(dst + (sub + LPT * i) * 8) = pack8(v[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// window attention. One warp = one (window, head): 49 tokens, d = 32, mma.sync m16n8k16 / m16n8k8 (fp16 in, fp32 acc).
// (A 49x49x32 problem per warp is far below a tcgen05 tile; the op is bound by reading q,k,v and writing o once.)
//   S = (q * 32^-0.5) k^T + rel_bias[head] + shift_mask ; P = softmax(S) ; O = P v
// qkv rows are in window order: 49 consecutive rows per window; q at cols [h*32), k at [C + h*32), v at [2C + h*32).
// A CTA (8 warps) serves ONE head and walks over windows: the head's relative-position bias is laid out once per CTA
// in shared memory in mma-fragment order (one 16 B load per accumulator quad, pad keys = -inf, log2 domain); K and V
// go global -> shared with cp.async (80 B row stride: conflict-free ldmatrix), V is consumed through ldmatrix.trans
// (no transposed copy), Q fragments come straight from global; keys are padded to 56 = 3 k16 steps + 1 k8 step.
// The shift mask (swin_transformer.py:183-196) only exists in the last window row / column and is skipped elsewhere.
// ------------------------------------------------------------------------------------------------
constexpr int kAttnWarps = 8;
constexpr int kRowStride = HDIM + 8;     // halfs; 80 B rows
constexpr int kKVRows = 56;
constexpr int kBiasFloats = 4 * 7 * 32 * 4;
constexpr int kAttnSmem = kBiasFloats * 4 + kAttnWarps * (2 * kKVRows * kRowStride * 2 + 64);

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnWarps * 32, 2)
window_attn_kernel(const __half* __restrict__ qkv, int ld_qkv, const float* __restrict__ rel_bias, int n_win_total, int H,
                   int W, int C, int shift, __half* __restrict__ out, int ld_out, const uint8_t* __restrict__ live) {
  extern __shared__ __align__(16) uint8_t attn_smem[];
  float* sBias = reinterpret_cast<float*>(attn_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __half* sK = reinterpret_cast<__half*>(attn_smem + kBiasFloats * 4) + warp * (2 * kKVRows * kRowStride);
  __half* sV = sK + kKVRows * kRowStride;
  uint8_t* sCode = attn_smem + kBiasFloats * 4 + kAttnWarps * (2 * kKVRows * kRowStride * 2) + warp * 64;
  const int head = blockIdx.y;
  constexpr float kL2e = 1.4426950408889634f;
  {
    // bias table in accumulator-fragment order: entry ((mt*7+nt)*32 + lane) = {r0c0, r0c1, r1c0, r1c1}
    const float* bias = rel_bias + (size_t)head * WT * WT;
    for (int idx = threadIdx.x; idx < kBiasFloats; idx += blockDim.x) {
      const int e = idx & 3, ln = (idx >> 2) & 31, f = idx >> 7, mt = f / 7, nt = f % 7;
      const int row = mt * 16 + (ln >> 2) + (e >> 1) * 8, col = nt * 8 + (ln & 3) * 2 + (e & 1);
      sBias[idx] = col < WT ? (row < WT ? __ldg(bias + row * WT + col) * kL2e : 0.f) : -1e30f;
    }
    // padding keys 49..55 stay zero for the whole kernel (cp.async only ever writes rows < 49)
    for (int i = lane; i < (kKVRows - WT) * kRowStride / 2; i += 32) {
      reinterpret_cast<uint32_t*>(sK + WT * kRowStride)[i] = 0u;
      reinterpret_cast<uint32_t*>(sV + WT * kRowStride)[i] = 0u;
    }
    sCode[lane] = 0; sCode[lane + 32] = 0;
  }
  __syncthreads();
  const int nWw = W / WS, nWh = H / WS;
  const int qr = lane >> 2, qc = (lane & 3) * 2;
  const float qs2 = 0.17677669529663687f * kL2e;    // 32^-0.5 (swin_transformer.py:177) in the log2 domain
  const float mask2 = -100.f * kL2e;                // attn_mask fill value (swin_transformer.py:195)
  const int win_per_img = (H / WS) * (W / WS);
  for (int win = blockIdx.x * kAttnWarps + warp; win < n_win_total; win += gridDim.x * kAttnWarps) {
    if (live && !__ldg(live + win / win_per_img)) continue;       // dead image (see all_dead)
    const __half* base = qkv + (size_t)win * WT * ld_qkv + head * HDIM;
    for (int i = lane; i < WT * 4; i += 32) {
      const int r = i >> 2, c8 = (i & 3) * 8;
      const __half* row = base + (size_t)r * ld_qkv + c8;
      cp_async16(sK + r * kRowStride + c8, row + C);
      cp_async16(sV + r * kRowStride + c8, row + 2 * C);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // Q fragments for all four 16-row steps (padding rows >= 49 alias row 0; their results are never stored)
    uint32_t aq[4][2][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int r0 = mt * 16 + qr, r1 = r0 + 8;
      const __half* q0 = base + (size_t)(r0 < WT ? r0 : 0) * ld_qkv;
      const __half* q1 = base + (size_t)(r1 < WT ? r1 : 0) * ld_qkv;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        aq[mt][ks][0] = __ldg(reinterpret_cast<const uint32_t*>(q0 + ks * 16 + qc));
        aq[mt][ks][1] = __ldg(reinterpret_cast<const uint32_t*>(q1 + ks * 16 + qc));
        aq[mt][ks][2] = __ldg(reinterpret_cast<const uint32_t*>(q0 + ks * 16 + 8 + qc));
        aq[mt][ks][3] = __ldg(reinterpret_cast<const uint32_t*>(q1 + ks * 16 + 8 + qc));
      }
    }
    // shift-mask region codes: rows / columns of the shifted map fall in 3 bands; only the last window row / column
    // mixes bands
    const int wimg = win % (nWh * nWw), wy = wimg / nWw, wx = wimg % nWw;
    const bool masked = shift > 0 && (wy == nWh - 1 || wx == nWw - 1);
    if (masked) {
      for (int t = lane; t < WT; t += 32) {
        const int iy = t / WS, ix = t % WS;
        const int hb = wy == nWh - 1 ? (iy < WS - shift ? 1 : 2) : 0;
        const int wb = wx == nWw - 1 ? (ix < WS - shift ? 1 : 2) : 0;
        sCode[t] = (uint8_t)(hb * 3 + wb);
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    __half* obase = out + (size_t)win * WT * ld_out + head * HDIM;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int r0 = mt * 16 + qr, r1 = r0 + 8;
      float s[7][4];
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) {
        uint32_t kb[4];
        ldsm_x4(kb, sK + (nt * 8 + (lane & 7)) * kRowStride + (lane >> 3) * 8);
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        mma16816(s[nt], aq[mt][0], kb[0], kb[1]);
        mma16816(s[nt], aq[mt][1], kb[2], kb[3]);
      }
      const float4* bt = reinterpret_cast<const float4*>(sBias) + (mt * 7) * 32 + lane;
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) {
        const float4 bb = bt[nt * 32];
        s[nt][0] = fmaf(s[nt][0], qs2, bb.x); s[nt][1] = fmaf(s[nt][1], qs2, bb.y);
        s[nt][2] = fmaf(s[nt][2], qs2, bb.z); s[nt][3] = fmaf(s[nt][3], qs2, bb.w);
      }
      if (masked) {
        const int c0 = sCode[r0], c1 = sCode[r1];
#pragma unroll
        for (int nt = 0; nt < 7; ++nt) {
          const int ca = sCode[nt * 8 + qc], cb = sCode[nt * 8 + qc + 1];
          if (ca != c0) s[nt][0] += mask2;
          if (cb != c0) s[nt][1] += mask2;
          if (ca != c1) s[nt][2] += mask2;
          if (cb != c1) s[nt][3] += mask2;
        }
      }
      float m0 = fmaxf(s[0][0], s[0][1]), m1 = fmaxf(s[0][2], s[0][3]);
#pragma unroll
      for (int nt = 1; nt < 7; ++nt) {
        m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
        m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
      }
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) {
        s[nt][0] = ex2f(s[nt][0] - m0); s[nt][1] = ex2f(s[nt][1] - m0);
        s[nt][2] = ex2f(s[nt][2] - m1); s[nt][3] = ex2f(s[nt][3] - m1);
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      // O = P V (unnormalised P <= 1 in fp16; 1/l applied to the fp32 result)
      float o[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) {
        uint32_t ap[4];
        ap[0] = pack_f16x2(s[2 * kk][0], s[2 * kk][1]);
        ap[1] = pack_f16x2(s[2 * kk][2], s[2 * kk][3]);
        ap[2] = pack_f16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        ap[3] = pack_f16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t vb[4];
          ldsm_x4_t(vb, sV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kRowStride + (np * 2 + (lane >> 4)) * 8);
          mma16816(o[np * 2], ap, vb[0], vb[1]);
          mma16816(o[np * 2 + 1], ap, vb[2], vb[3]);
        }
      }
      {
        uint32_t vb[4];
        ldsm_x4_t(vb, sV + (48 + (lane & 7)) * kRowStride + (lane >> 3) * 8);
        const uint32_t a0 = pack_f16x2(s[6][0], s[6][1]), a1 = pack_f16x2(s[6][2], s[6][3]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma1688(o[nt], a0, a1, vb[nt]);
      }
      const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        if (r0 < WT)
          *reinterpret_cast<uint32_t*>(obase + (size_t)r0 * ld_out + nt * 8 + qc) = pack_f16x2(o[nt][0] * i0, o[nt][1] * i0);
        if (r1 < WT)
          *reinterpret_cast<uint32_t*>(obase + (size_t)r1 * ld_out + nt * 8 + qc) = pack_f16x2(o[nt][2] * i1, o[nt][3] * i1);
      }
    }
    __syncwarp();      // every lane is done with sK / sV / sCode before the next window's cp.async overwrites them
  }
}

// tokens handled per CTA iteration = 8 warps x tokens-per-warp; grid capped at 16 CTAs per SM (grid-stride loop)
int token_grid(long long n_tok, int tok_per_warp) {
  long long blocks = (n_tok + 8 * tok_per_warp - 1) / (8 * tok_per_warp);
  const long long cap = (long long)tmp::num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// the four Swin-T stages (swin_transformer.py:598-606: embed_dim 96, depths [2,2,6,2], 224x224 input, patch 4)
int stage_index(int H, int W, int C) {
  if (H != W) return -1;
  if (C == 96 && H == 56) return 0;
  if (C == 192 && H == 28) return 1;
  if (C == 384 && H == 14) return 2;
  if (C == 768 && H == 7) return 3;
  return -1;
}

}  // namespace

// `live` (every tmp_swin_* call): optional uint8 [n_img], 0 = the image's features have no consumer, skip its work
extern "C" int tmp_swin_patch_embed_ln(const float* img, int n_img, const float* Wt, const float* bconv, const float* g,
                                       const float* b, void* out, int Cp, const uint8_t* live, void* stream) {
  TMP_REQUIRE(img && Wt && bconv && g && b && out && n_img > 0 && n_img <= 600000 && Cp >= 96 && Cp % 8 == 0,
              "swin_patch_embed_ln: bad argument");
  const int n_tok = n_img * 3136;
  patch_embed_ln_kernel<<<token_grid(n_tok, 8), 256, 0, (cudaStream_t)stream>>>(img, n_tok, Wt, bconv, g, b, (__half*)out, Cp, live);
  return tmp::check_launch("patch_embed_ln_kernel");
}

#define SWIN_STAGE_DISPATCH(si, CALL)                  \
  switch (si) {                                        \
    case 0: { CALL(96, 56); break; }                   \
    case 1: { CALL(192, 28); break; }                  \
    case 2: { CALL(384, 14); break; }                  \
    default: { CALL(768, 7); break; }                  \
  }

// zero_dead: dead images' output rows are written as zeros (the last kernel of the encoder) instead of being skipped
extern "C" int tmp_swin_ln_window(const void* x, const float* g, const float* b, int n_img, int H, int W, int C, int Cp,
                                  int shift, void* out, const uint8_t* live, int zero_dead, void* stream) {
  const int si = stage_index(H, W, C);
  TMP_REQUIRE(x && g && b && out && n_img > 0 && n_img <= 600000 && si >= 0 && Cp >= C && Cp % 8 == 0 && shift >= 0 &&
                  shift < WS, "swin_ln_window: bad argument (H=%d W=%d C=%d Cp=%d shift=%d; Swin-T stages only)", H, W, C,
              Cp, shift);
  if (H <= WS) shift = 0;   // window covers the whole map: torchvision disables the shift (swin_transformer.py:141-145)
#define CALL(C_, H_)                                                                                              \
  ln_window_kernel<C_, H_><<<token_grid((long long)n_img * H_ * H_, 32 / (C_ / 24)), 256, 0, (cudaStream_t)stream>>>( \
      (const __half*)x, g, b, n_img, Cp, shift, (__half*)out, live, zero_dead)
  SWIN_STAGE_DISPATCH(si, CALL)
#undef CALL
  return tmp::check_launch("ln_window_kernel");
}

extern "C" int tmp_swin_unwindow_add_ln(const void* y, void* x, const float* g, const float* b, int n_img, int H, int W,
                                        int C, int Cp, int shift, void* hn, const uint8_t* live, void* stream) {
  const int si = stage_index(H, W, C);
  TMP_REQUIRE(y && x && n_img > 0 && n_img <= 600000 && si >= 0 && Cp >= C && Cp % 8 == 0 && shift >= 0 && shift < WS &&
                  (!hn || (g && b)), "swin_unwindow_add_ln: bad argument");
  if (H <= WS) shift = 0;
#define CALL(C_, H_)                                                                                                    \
  unwindow_add_ln_kernel<C_, H_><<<token_grid((long long)n_img * H_ * H_, 32 / (C_ / 24)), 256, 0, (cudaStream_t)stream>>>( \
      (const __half*)y, (__half*)x, g, b, n_img, Cp, shift, (__half*)hn, live)
  SWIN_STAGE_DISPATCH(si, CALL)
#undef CALL
  return tmp::check_launch("unwindow_add_ln_kernel");
}

extern "C" int tmp_swin_merge_ln(const void* x, const float* g, const float* b, int n_img, int H, int W, int C, int Cp,
                                 void* out, const uint8_t* live, void* stream) {
  const int si = stage_index(H, W, C);
  TMP_REQUIRE(x && g && b && out && n_img > 0 && n_img <= 600000 && si >= 0 && si < 3 && Cp >= C && Cp % 8 == 0,
              "swin_merge_ln: bad argument");
#define CALL(C_, H_)                                                                                                       \
  merge_ln_kernel<C_, H_><<<token_grid((long long)n_img * (H_ / 2) * (H_ / 2), 32 / (C_ / 12)), 256, 0, (cudaStream_t)stream>>>( \
      (const __half*)x, g, b, n_img, Cp, (__half*)out, live)
  switch (si) {
    case 0: { CALL(96, 56); break; }
    case 1: { CALL(192, 28); break; }
    default: { CALL(384, 14); break; }
  }
#undef CALL
  return tmp::check_launch("merge_ln_kernel");
}

extern "C" int tmp_swin_window_attn(const void* qkv, int ld_qkv, const float* rel_bias, int n_img, int H, int W, int C,
                                    int heads, int shift, void* out, int ld_out, const uint8_t* live, void* stream) {
  TMP_REQUIRE(qkv && rel_bias && out && n_img > 0 && H % WS == 0 && W % WS == 0 && heads > 0 && heads <= 65535 &&
                  C == heads * HDIM && ld_qkv >= 3 * C && ld_qkv % 8 == 0 && ld_out >= C && ld_out % 8 == 0 && shift >= 0 &&
                  shift < WS, "swin_window_attn: bad argument");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e != cudaSuccess) {
      tmp::set_error("cudaFuncSetAttribute(window_attn): %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  const long long n_win = (long long)n_img * (H / WS) * (W / WS);
  TMP_REQUIRE(n_win < (1ll << 30), "swin_window_attn: too many windows");
  // CTAs of one head walk over the windows; ~4 CTAs per SM in flight-order keeps the tail short
  long long gx = (n_win + kAttnWarps - 1) / kAttnWarps;
  const long long cap = ((long long)tmp::num_sms() * 4 + heads - 1) / heads;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)heads);
  window_attn_kernel<<<grid, kAttnWarps * 32, kAttnSmem, (cudaStream_t)stream>>>(
      (const __half*)qkv, ld_qkv, rel_bias, (int)n_win, H, W, C, H > WS ? shift : 0, (__half*)out, ld_out, live);
  return tmp::check_launch("window_attn_kernel");
}
