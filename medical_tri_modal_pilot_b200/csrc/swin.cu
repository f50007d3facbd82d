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
constexpr int kMaxChunks = 6;    // 8-channel chunks per lane: 4C = 1536 -> 192 chunks -> 6 per lane

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
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

// nn.LayerNorm(C, eps=1e-5) of one token held as `nch` 8-channel chunks spread over the lanes of a warp
// (chunk c = lane + 32*i). v is normalised in place; chunks >= nch must hold zeros and are left untouched.
template <int MAXI>
__device__ __forceinline__ void warp_layernorm(float (&v)[MAXI][8], int nch, int lane, int C, const float* __restrict__ g,
                                               const float* __restrict__ b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXI; ++i)
    if (lane + 32 * i < nch) {
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[i][k];
    }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXI; ++i)
    if (lane + 32 * i < nch) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[i][k] -= mean;
        q += v[i][k] * v[i][k];
      }
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    const int c = lane + 32 * i;
    if (c < nch) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + c * 8)), g1 = __ldg(reinterpret_cast<const float4*>(g + c * 8) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c * 8)), b1 = __ldg(reinterpret_cast<const float4*>(b + c * 8) + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) v[i][k] = fmaf(v[i][k] * rstd, gg[k], bb[k]);
    }
  }
}

// row of token (n, y, x) of a [N, H, W] map in WINDOW order after a cyclic shift by `shift`
// (torch.roll(x, -shift) then partition: swin_transformer.py:146-155): position (y, x) of the SHIFTED map.
__device__ __forceinline__ long long window_row(int n, int ys, int xs, int H, int W) {
  const int nWw = W / WS, nWh = H / WS;
  const int wy = ys / WS, iy = ys % WS, wx = xs / WS, ix = xs % WS;
  return (((long long)n * nWh + wy) * nWw + wx) * WT + iy * WS + ix;
}

// ------------------------------------------------------------------------------------------------
// patch embedding: img fp32 [N,224,224] -> tokens [N*56*56, Cp] fp16 = LN(conv4x4s4(img))
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patch_embed_ln_kernel(const float* __restrict__ img, long long n_tok,
                                                             const float* __restrict__ Wt /*[16][96]*/,
                                                             const float* __restrict__ bconv, const float* __restrict__ g,
                                                             const float* __restrict__ b, __half* __restrict__ out, int Cp) {
  __shared__ float sW[16 * 96];
  __shared__ float sB[96], sG[96], sBe[96];
  for (int i = threadIdx.x; i < 16 * 96; i += blockDim.x) sW[i] = Wt[i];
  for (int i = threadIdx.x; i < 96; i += blockDim.x) { sB[i] = bconv[i]; sG[i] = g[i]; sBe[i] = b[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tok < n_tok; tok += warps) {
    const int n = (int)(tok / 3136), r = (int)(tok % 3136), ty = r / 56, tx = r % 56;
    // lanes 0..15 fetch the 4x4 patch (row-major: p = dy*4 + dx, the Conv2d weight order [out,1,4,4])
    float pix = 0.f;
    if (lane < 16) pix = __ldg(img + ((size_t)n * 224 + ty * 4 + (lane >> 2)) * 224 + tx * 4 + (lane & 3));
    float acc[3] = {sB[lane], sB[lane + 32], sB[lane + 64]};
#pragma unroll
    for (int p = 0; p < 16; ++p) {
      const float pv = __shfl_sync(0xffffffffu, pix, p);
      acc[0] = fmaf(pv, sW[p * 96 + lane], acc[0]);
      acc[1] = fmaf(pv, sW[p * 96 + lane + 32], acc[1]);
      acc[2] = fmaf(pv, sW[p * 96 + lane + 64], acc[2]);
    }
    const float mean = warp_sum(acc[0] + acc[1] + acc[2]) * (1.f / 96.f);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) { acc[k] -= mean; q += acc[k] * acc[k]; }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / 96.f) + 1e-5f);
    __half* o = out + tok * Cp;
#pragma unroll
    for (int k = 0; k < 3; ++k) o[lane + 32 * k] = __float2half_rn(fmaf(acc[k] * rstd, sG[lane + 32 * k], sBe[lane + 32 * k]));
    for (int c = 96 + lane; c < Cp; c += 32) o[c] = __float2half_rn(0.f);
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm + cyclic shift + window partition:  out[window_row] = LN(x[natural row])
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ln_window_kernel(const __half* __restrict__ x, const float* __restrict__ g,
                                                        const float* __restrict__ b, int n_img, int H, int W, int C, int Cp,
                                                        int shift, __half* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int nch = C >> 3, pch = Cp >> 3;
  const long long n_tok = (long long)n_img * H * W;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tok < n_tok; tok += warps) {
    // iterate over OUTPUT rows (window order) so that writes of neighbouring warps are contiguous
    const int per_img = H * W;
    const int n = (int)(tok / per_img), rw = (int)(tok % per_img);
    const int nWw = W / WS;
    const int win = rw / WT, t = rw % WT;
    const int ys = (win / nWw) * WS + t / WS, xs = (win % nWw) * WS + t % WS;     // position in the shifted map
    const int y = (ys + shift) % H, xx = (xs + shift) % W;                        // source position (roll by -shift)
    const __half* src = x + ((size_t)n * per_img + (size_t)y * W + xx) * Cp;
    float v[3][8];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = lane + 32 * i;
      if (c < nch) unpack8(*reinterpret_cast<const uint4*>(src + c * 8), v[i]);
      else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[i][k] = 0.f;
      }
    }
    warp_layernorm<3>(v, nch, lane, C, g, b);
    __half* dst = out + tok * Cp;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = lane + 32 * i;
      if (c < pch) *reinterpret_cast<uint4*>(dst + c * 8) = pack8(v[i]);   // pad chunks hold zeros
    }
  }
}

// ------------------------------------------------------------------------------------------------
// window reverse + reverse shift + residual add + LayerNorm:
//   x[natural] += y[window_row];  hn[natural] = LN(x[natural])      (hn may be null: residual add only)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unwindow_add_ln_kernel(const __half* __restrict__ y, __half* __restrict__ x,
                                                              const float* __restrict__ g, const float* __restrict__ b,
                                                              int n_img, int H, int W, int C, int Cp, int shift,
                                                              __half* __restrict__ hn) {
  const int lane = threadIdx.x & 31;
  const int nch = C >> 3, pch = Cp >> 3;
  const int per_img = H * W;
  const long long n_tok = (long long)n_img * per_img;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tok < n_tok; tok += warps) {
    const int n = (int)(tok / per_img), r = (int)(tok % per_img), yy = r / W, xx = r % W;
    const int ys = (yy - shift + H) % H, xs = (xx - shift + W) % W;               // where this token sits after roll(-shift)
    const __half* ysrc = y + window_row(n, ys, xs, H, W) * Cp;
    __half* xr = x + tok * Cp;
    float v[3][8];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = lane + 32 * i;
      if (c < nch) {
        float a[8];
        unpack8(*reinterpret_cast<const uint4*>(xr + c * 8), v[i]);
        unpack8(*reinterpret_cast<const uint4*>(ysrc + c * 8), a);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[i][k] += a[k];
        const uint4 packed = pack8(v[i]);
        *reinterpret_cast<uint4*>(xr + c * 8) = packed;
        unpack8(packed, v[i]);            // normalise what the residual stream actually holds
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[i][k] = 0.f;
      }
    }
    if (hn) {
      warp_layernorm<3>(v, nch, lane, C, g, b);
      __half* dst = hn + tok * Cp;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = lane + 32 * i;
        if (c < pch) *reinterpret_cast<uint4*>(dst + c * 8) = pack8(v[i]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// patch merging gather + LayerNorm(4C): out[(n,i,j), :] = LN([x(2i,2j) | x(2i+1,2j) | x(2i,2j+1) | x(2i+1,2j+1)])
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) merge_ln_kernel(const __half* __restrict__ x, const float* __restrict__ g,
                                                       const float* __restrict__ b, int n_img, int H, int W, int C, int Cp,
                                                       __half* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int H2 = H / 2, W2 = W / 2, C4 = 4 * C;
  const int cch = C >> 3, nch = C4 >> 3;
  const long long n_tok = (long long)n_img * H2 * W2;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tok < n_tok; tok += warps) {
    const int n = (int)(tok / (H2 * W2)), r = (int)(tok % (H2 * W2)), i2 = r / W2, j2 = r % W2;
    float v[kMaxChunks][8];
#pragma unroll
    for (int i = 0; i < kMaxChunks; ++i) {
      const int c = lane + 32 * i;
      if (c < nch) {
        const int part = c / cch, cc = c % cch;      // part 0..3 = x0 (0,0), x1 (1,0), x2 (0,1), x3 (1,1)
        const int yy = 2 * i2 + (part & 1), xx = 2 * j2 + (part >> 1);
        unpack8(*reinterpret_cast<const uint4*>(x + (((size_t)n * H + yy) * W + xx) * Cp + cc * 8), v[i]);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[i][k] = 0.f;
      }
    }
    warp_layernorm<kMaxChunks>(v, nch, lane, C4, g, b);
    __half* dst = out + tok * C4;
#pragma unroll
    for (int i = 0; i < kMaxChunks; ++i) {
      const int c = lane + 32 * i;
      if (c < nch) *reinterpret_cast<uint4*>(dst + c * 8) = pack8(v[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// window attention. One warp = one (window, head): 49 tokens, d = 32, mma.sync m16n8k16 (fp16 in, fp32 acc).
// (A 49x49x32 problem per warp is far below a tcgen05 tile; the op is bound by reading q,k,v and writing o once.)
//   S = (q * 32^-0.5) k^T + rel_bias[head] + shift_mask ; P = softmax(S) ; O = P v
// qkv rows are in window order: 49 consecutive rows per window; q at cols [h*32), k at [C + h*32), v at [2C + h*32).
// ------------------------------------------------------------------------------------------------
constexpr int kAttnWarps = 4;
constexpr int kQKStride = HDIM + 8;   // halfs; 80 B rows keep 4-byte fragment loads conflict-free
constexpr int kVtStride = 64 + 8;     // V^T [32][64 keys]

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kAttnWarps * 32) window_attn_kernel(const __half* __restrict__ qkv, int ld_qkv,
                                                                      const float* __restrict__ rel_bias, int n_win_total,
                                                                      int H, int W, int C, int heads, int shift,
                                                                      __half* __restrict__ out, int ld_out) {
  __shared__ __align__(16) __half sK[kAttnWarps][64 * kQKStride];
  __shared__ __align__(16) __half sVt[kAttnWarps][HDIM * kVtStride];
  __shared__ int sReg[kAttnWarps][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pair = (long long)blockIdx.x * kAttnWarps + warp;     // (window, head), head fastest
  if (pair >= (long long)n_win_total * heads) return;
  const int head = (int)(pair % heads);
  const long long win = pair / heads;
  const int nWw = W / WS, nWh = H / WS;
  const int wimg = (int)(win % (nWh * nWw)), wy = wimg / nWw, wx = wimg % nWw;
  __half* K = sK[warp];
  __half* Vt = sVt[warp];
  const __half* base = qkv + (size_t)win * WT * ld_qkv + head * HDIM;
  const float qscale = 0.17677669529663687f;        // 32^-0.5 (swin_transformer.py:177), applied to q k^T
  // zero-fill the padding keys 49..63, then load K and V^T (4 lanes x 16 B per 64 B row)
  for (int i = lane; i < 64 * kQKStride / 2; i += 32) reinterpret_cast<uint32_t*>(K)[i] = 0u;
  for (int i = lane; i < HDIM * kVtStride / 2; i += 32) reinterpret_cast<uint32_t*>(Vt)[i] = 0u;
  __syncwarp();
  for (int i = lane; i < WT * 4; i += 32) {
    const int r = i >> 2, c8 = (i & 3) * 8;
    const __half* row = base + (size_t)r * ld_qkv + c8;
    const uint4 k = *reinterpret_cast<const uint4*>(row + C);
    const uint4 v = *reinterpret_cast<const uint4*>(row + 2 * C);
    *reinterpret_cast<uint4*>(K + r * kQKStride + c8) = k;
    const __half* vh = reinterpret_cast<const __half*>(&v);
#pragma unroll
    for (int t = 0; t < 8; ++t) Vt[(c8 + t) * kVtStride + r] = vh[t];
  }
  // shift-mask region code per token (swin_transformer.py:183-196): rows/cols of the shifted map fall in 3 bands
  for (int t = lane; t < 64; t += 32) {
    int code = 0;
    if (shift > 0 && t < WT) {
      const int ys = wy * WS + t / WS, xs = wx * WS + t % WS;
      const int hb = ys < H - WS ? 0 : (ys < H - shift ? 1 : 2);
      const int wb = xs < W - WS ? 0 : (xs < W - shift ? 1 : 2);
      code = hb * 3 + wb;
    }
    sReg[warp][t] = code;
  }
  __syncwarp();
  const float* bias = rel_bias + (size_t)head * WT * WT;
  const int qr = lane >> 2, qc = (lane & 3) * 2;
  __half* obase = out + (size_t)win * WT * ld_out + head * HDIM;
#pragma unroll 1
  for (int mt = 0; mt < 4; ++mt) {             // 16 query rows per step (rows 49..63 are padding)
    const int r0 = mt * 16 + qr, r1 = r0 + 8;
    // A fragments of Q straight from global (each 64 B query row is read once by the 4 lanes of a quad);
    // padding rows (>= 49) alias row 0, their results are never stored
    const __half* q0 = base + (size_t)(r0 < WT ? r0 : 0) * ld_qkv;
    const __half* q1 = base + (size_t)(r1 < WT ? r1 : 0) * ld_qkv;
    uint32_t aq[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      aq[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(q0 + ks * 16 + qc));
      aq[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(q1 + ks * 16 + qc));
      aq[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(q0 + ks * 16 + 8 + qc));
      aq[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(q1 + ks * 16 + 8 + qc));
    }
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      if (nt < 7) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(K + (nt * 8 + qr) * kQKStride + ks * 16 + qc);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(K + (nt * 8 + qr) * kQKStride + ks * 16 + 8 + qc);
          mma16816(s[nt], aq[ks], b0, b1);
        }
      }
    }
    // bias + mask + softmax over the 49 keys. Thread holds rows r0 (c0,c1) and r1 (c2,c3), columns nt*8 + qc + {0,1}.
    const int reg0 = sReg[warp][r0 & 63], reg1 = sReg[warp][r1 & 63];
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = nt * 8 + qc + e;
        const bool kv_ok = col < WT;
        const int regc = sReg[warp][col & 63];
        float v0 = -INFINITY, v1 = -INFINITY;
        if (kv_ok) {
          if (r0 < WT) v0 = fmaf(s[nt][e], qscale, __ldg(bias + r0 * WT + col) + (regc != reg0 ? -100.f : 0.f));
          if (r1 < WT) v1 = fmaf(s[nt][2 + e], qscale, __ldg(bias + r1 * WT + col) + (regc != reg1 ? -100.f : 0.f));
        }
        s[nt][e] = v0; s[nt][2 + e] = v1;
        m0 = fmaxf(m0, v0); m1 = fmaxf(m1, v1);
      }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    if (m0 == -INFINITY) m0 = 0.f;      // padding query rows
    if (m1 == -INFINITY) m1 = 0.f;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float p0 = __expf(s[nt][e] - m0), p1 = __expf(s[nt][2 + e] - m1);
        s[nt][e] = p0; s[nt][2 + e] = p1;
        l0 += p0; l1 += p1;
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
    // O = P V : k-steps of 16 keys; the A fragment of step kk is built from S n-tiles 2kk and 2kk+1
    float o[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ap[4];
      ap[0] = pack_f16x2(s[2 * kk][0] * i0, s[2 * kk][1] * i0);
      ap[1] = pack_f16x2(s[2 * kk][2] * i1, s[2 * kk][3] * i1);
      ap[2] = pack_f16x2(s[2 * kk + 1][0] * i0, s[2 * kk + 1][1] * i0);
      ap[3] = pack_f16x2(s[2 * kk + 1][2] * i1, s[2 * kk + 1][3] * i1);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(Vt + (nt * 8 + qr) * kVtStride + kk * 16 + qc);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(Vt + (nt * 8 + qr) * kVtStride + kk * 16 + 8 + qc);
        mma16816(o[nt], ap, b0, b1);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (r0 < WT) *reinterpret_cast<uint32_t*>(obase + (size_t)r0 * ld_out + nt * 8 + qc) = pack_f16x2(o[nt][0], o[nt][1]);
      if (r1 < WT) *reinterpret_cast<uint32_t*>(obase + (size_t)r1 * ld_out + nt * 8 + qc) = pack_f16x2(o[nt][2], o[nt][3]);
    }
  }
}

int token_grid(long long n_tok) {
  long long blocks = (n_tok + 7) / 8;
  const long long cap = (long long)tmp::num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

bool stage_ok(int H, int W, int C, int Cp) {
  return H > 0 && W > 0 && H % WS == 0 && W % WS == 0 && C % 32 == 0 && C <= 768 && Cp >= C && Cp % 8 == 0 && Cp <= 768;
}

}  // namespace

extern "C" int tmp_swin_patch_embed_ln(const float* img, int n_img, const float* Wt, const float* bconv, const float* g,
                                       const float* b, void* out, int Cp, void* stream) {
  TMP_REQUIRE(img && Wt && bconv && g && b && out && n_img > 0 && Cp >= 96 && Cp % 8 == 0, "swin_patch_embed_ln: bad argument");
  const long long n_tok = (long long)n_img * 3136;
  patch_embed_ln_kernel<<<token_grid(n_tok), 256, 0, (cudaStream_t)stream>>>(img, n_tok, Wt, bconv, g, b, (__half*)out, Cp);
  return tmp::check_launch("patch_embed_ln_kernel");
}

extern "C" int tmp_swin_ln_window(const void* x, const float* g, const float* b, int n_img, int H, int W, int C, int Cp,
                                  int shift, void* out, void* stream) {
  TMP_REQUIRE(x && g && b && out && n_img > 0 && stage_ok(H, W, C, Cp) && shift >= 0 && shift < WS,
              "swin_ln_window: bad argument (H=%d W=%d C=%d Cp=%d shift=%d)", H, W, C, Cp, shift);
  if (H <= WS) shift = 0;   // window covers the whole map: torchvision disables the shift (swin_transformer.py:141-145)
  ln_window_kernel<<<token_grid((long long)n_img * H * W), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)x, g, b, n_img, H, W, C, Cp, shift, (__half*)out);
  return tmp::check_launch("ln_window_kernel");
}

extern "C" int tmp_swin_unwindow_add_ln(const void* y, void* x, const float* g, const float* b, int n_img, int H, int W,
                                        int C, int Cp, int shift, void* hn, void* stream) {
  TMP_REQUIRE(y && x && n_img > 0 && stage_ok(H, W, C, Cp) && shift >= 0 && shift < WS && (!hn || (g && b)),
              "swin_unwindow_add_ln: bad argument");
  if (H <= WS) shift = 0;
  unwindow_add_ln_kernel<<<token_grid((long long)n_img * H * W), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)y, (__half*)x, g, b, n_img, H, W, C, Cp, shift, (__half*)hn);
  return tmp::check_launch("unwindow_add_ln_kernel");
}

extern "C" int tmp_swin_merge_ln(const void* x, const float* g, const float* b, int n_img, int H, int W, int C, int Cp,
                                 void* out, void* stream) {
  TMP_REQUIRE(x && g && b && out && n_img > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0 && 4 * C <= 8 * 32 * kMaxChunks &&
                  Cp >= C, "swin_merge_ln: bad argument");
  merge_ln_kernel<<<token_grid((long long)n_img * (H / 2) * (W / 2)), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)x, g, b, n_img, H, W, C, Cp, (__half*)out);
  return tmp::check_launch("merge_ln_kernel");
}

extern "C" int tmp_swin_window_attn(const void* qkv, int ld_qkv, const float* rel_bias, int n_img, int H, int W, int C,
                                    int heads, int shift, void* out, int ld_out, void* stream) {
  TMP_REQUIRE(qkv && rel_bias && out && n_img > 0 && H % WS == 0 && W % WS == 0 && heads > 0 && C == heads * HDIM &&
                  ld_qkv >= 3 * C && ld_qkv % 8 == 0 && ld_out >= C && ld_out % 8 == 0 && shift >= 0 && shift < WS,
              "swin_window_attn: bad argument");
  const long long n_win = (long long)n_img * (H / WS) * (W / WS);
  const long long pairs = n_win * heads;
  window_attn_kernel<<<(unsigned)((pairs + kAttnWarps - 1) / kAttnWarps), kAttnWarps * 32, 0, (cudaStream_t)stream>>>(
      (const __half*)qkv, ld_qkv, rel_bias, (int)n_win, H, W, C, heads, H > WS ? shift : 0, (__half*)out, ld_out);
  return tmp::check_launch("window_attn_kernel");
}
