// head.cu -- classifier head of TRI_MBT_VSLTCLS in training mode, three launches instead of the
// ~45 ATen launches (two SIMT sgemm's among them) that ran between the fused forward and the fused backward with most SMs
// idle. Reference tri_mbt_vsltcls.py:176-177 (demographic branch), :248-255 (head), definitions :72-76, :152-158:
//   c      = LayerNorm_256(cls)                                 nn.LayerNorm, eps 1e-5, biased variance
//   demo   = ReLU(LayerNorm_256(Linear(2,256)([age, gender])))
//   z      = [c | demo]                                          [B, 512]
//   h      = z W1^T + b1                                         Linear(512, 256)
//   a      = ReLU(BatchNorm1d(h))                                batch statistics (biased variance), running statistics
//                                                                updated with momentum (unbiased variance), as nn.BatchNorm1d
//   logit  = a w3^T + b3                                         Linear(256, 1)
// Everything is fp32 (the head stays outside the 16-bit plan). Work split: the BatchNorm statistics are per output
// column of h over the batch, so a CTA owns kCols columns of h for ALL rows (its slice of W1 stays in shared memory; the two
// LayerNorms in front are recomputed by every CTA, 2 x 256 values per row) and the statistics never leave the CTA. What does
// cross CTAs -- the logit (a sum over columns) and, in the backward, the parameter gradients that are sums over rows -- goes
// through per-CTA partials that the LAST CTA to finish adds up in a fixed order: no floating-point atomics, results are
// bit-reproducible from run to run.
#include "common.cuh"

namespace {

constexpr int D = 256;          // model width
constexpr int Z = 2 * D;        // classifier input: [LayerNorm(cls) | demographic embedding]
constexpr int kCols = 8;        // columns of h per CTA (forward, backward 1)
constexpr int kGrid = D / kCols;
constexpr int kThreads = 256;
constexpr int kChunk = 32;      // rows per GEMM pass
constexpr int kLd = Z + 4;      // padded row pitch in shared memory (floats; keeps float4 alignment)
constexpr int kRowsB2 = 4;      // rows per pass in backward 2
constexpr int kRowsPerWarp = kChunk / (kThreads / 32);
static_assert(kRowsPerWarp * (kThreads / 32) == kChunk, "head: chunk rows must split evenly over the warps");

struct HeadParams {
  const float* ln_g; const float* ln_b;            // layer_norms_after_concat
  const float* Wd; const float* bd;                // ie_demo.0 : [256, 2], [256]
  const float* lnd_g; const float* lnd_b;          // ie_demo.1
  const float* W1; const float* b1;                // fc_list.0 : [256, 512], [256]
  const float* bn_g; const float* bn_b;            // fc_list.1
  const float* w3; const float* b3;                // fc_list.3 : [1, 256], [1]
};
struct HeadSaved {
  float* Zs;       // [B, 512]  classifier input
  float* XC;       // [B, 256]  normalised cls (before gamma / beta)
  float* XD;       // [B, 256]  normalised demographic pre-activation
  float* rstd_c;   // [B]
  float* rstd_d;   // [B]
  float* XH;       // [B, 256]  normalised h (BatchNorm xhat)
  float* invstd;   // [256]
};
struct HeadGrads {
  float* ln_g; float* ln_b; float* Wd; float* bd; float* lnd_g; float* lnd_b; float* W1; float* b1; float* bn_g; float* bn_b;
  float* w3; float* b3;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// true in every thread of the LAST block to get here (all blocks' earlier global writes are visible to it); resets the counter
__device__ __forceinline__ bool last_block(unsigned int* counter, unsigned int n_blocks) {
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(counter, 1u);
    s_last = (t == n_blocks - 1) ? 1u : 0u;
    if (s_last) *counter = 0u;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0u;
}

// LayerNorm (eps 1e-5, biased variance) of one 256-wide row held as 8 values per lane; returns rstd, v <- xhat
__device__ __forceinline__ float ln_row(float (&v)[8]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] -= mean; q = fmaf(v[i], v[i], q); }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] *= rstd;
  return rstd;
}

// ------------------------------------------------------------------------------------------------------------------------
// forward. grid = kGrid CTAs; CTA g owns columns [g*kCols, (g+1)*kCols) of h.
// dynamic smem: sW [kCols][kLd] | sZ [kChunk][kLd] | sH [B][kCols]
// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) head_fwd_kernel(const float* __restrict__ cls, const float* __restrict__ age,
                                                            const float* __restrict__ gen, int B, HeadParams p,
                                                            float* __restrict__ run_mean, float* __restrict__ run_var,
                                                            long long* __restrict__ nbt, float momentum, float bn_eps,
                                                            HeadSaved sv, float* __restrict__ part,
                                                            unsigned int* __restrict__ counter, float* __restrict__ logits) {
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;
  float* sZ = sW + kCols * kLd;
  float* sH = sZ + kChunk * kLd;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j0 = blockIdx.x * kCols;
  const bool writer = blockIdx.x == 0;       // one CTA stores the tensors the backward needs

  for (int i = tid; i < kCols * (Z / 4); i += kThreads) {
    const int c = i / (Z / 4), k4 = i % (Z / 4);
    *reinterpret_cast<float4*>(sW + c * kLd + k4 * 4) = __ldg(reinterpret_cast<const float4*>(p.W1 + (size_t)(j0 + c) * Z) + k4);
  }
  float lg[8], lb[8], dg[8], db[8], w0[8], w1[8], bdv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = lane * 8 + i;
    lg[i] = __ldg(p.ln_g + f); lb[i] = __ldg(p.ln_b + f);
    dg[i] = __ldg(p.lnd_g + f); db[i] = __ldg(p.lnd_b + f);
    w0[i] = __ldg(p.Wd + 2 * f); w1[i] = __ldg(p.Wd + 2 * f + 1); bdv[i] = __ldg(p.bd + f);
  }

  __syncthreads();
  for (int r0 = 0; r0 < B; r0 += kChunk) {
    __syncthreads();       // previous chunk's GEMM has read sZ
    // rows of the chunk: LayerNorm(cls) and the demographic branch -> sZ. A warp takes kRowsPerWarp rows; their loads are
    // issued together (the phase is a chain of global-load and shuffle latencies, not of work)
    {
      float v[kRowsPerWarp][8], ag[kRowsPerWarp], gn[kRowsPerWarp];
#pragma unroll
      for (int q = 0; q < kRowsPerWarp; ++q) {
        const int row = r0 + warp + q * (kThreads / 32);
        if (row < B) {
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(cls + (size_t)row * D + lane * 8));
          const float4 a1 = __ldg(reinterpret_cast<const float4*>(cls + (size_t)row * D + lane * 8 + 4));
          v[q][0] = a0.x; v[q][1] = a0.y; v[q][2] = a0.z; v[q][3] = a0.w;
          v[q][4] = a1.x; v[q][5] = a1.y; v[q][6] = a1.z; v[q][7] = a1.w;
          ag[q] = __ldg(age + row);
          gn[q] = __ldg(gen + row);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[q][i] = 0.f;
          ag[q] = gn[q] = 0.f;
        }
      }
#pragma unroll
      for (int q = 0; q < kRowsPerWarp; ++q) {
        const int rr = warp + q * (kThreads / 32);
        const int row = r0 + rr;
        float* zrow = sZ + rr * kLd;
        const float rc = ln_row(v[q]);
        float zc[8], u[8], zd[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) zc[i] = fmaf(v[q][i], lg[i], lb[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = fmaf(w0[i], ag[q], fmaf(w1[i], gn[q], bdv[i]));
        const float rd = ln_row(u);
#pragma unroll
        for (int i = 0; i < 8; ++i) zd[i] = fmaxf(fmaf(u[i], dg[i], db[i]), 0.f);
        if (row >= B) {
#pragma unroll
          for (int i = 0; i < 8; ++i) zc[i] = zd[i] = 0.f;
        }
        *reinterpret_cast<float4*>(zrow + lane * 8) = make_float4(zc[0], zc[1], zc[2], zc[3]);
        *reinterpret_cast<float4*>(zrow + lane * 8 + 4) = make_float4(zc[4], zc[5], zc[6], zc[7]);
        *reinterpret_cast<float4*>(zrow + D + lane * 8) = make_float4(zd[0], zd[1], zd[2], zd[3]);
        *reinterpret_cast<float4*>(zrow + D + lane * 8 + 4) = make_float4(zd[4], zd[5], zd[6], zd[7]);
        if (writer && row < B) {
          float* zs = sv.Zs + (size_t)row * Z;
          *reinterpret_cast<float4*>(zs + lane * 8) = make_float4(zc[0], zc[1], zc[2], zc[3]);
          *reinterpret_cast<float4*>(zs + lane * 8 + 4) = make_float4(zc[4], zc[5], zc[6], zc[7]);
          *reinterpret_cast<float4*>(zs + D + lane * 8) = make_float4(zd[0], zd[1], zd[2], zd[3]);
          *reinterpret_cast<float4*>(zs + D + lane * 8 + 4) = make_float4(zd[4], zd[5], zd[6], zd[7]);
          float* xc = sv.XC + (size_t)row * D + lane * 8;
          *reinterpret_cast<float4*>(xc) = make_float4(v[q][0], v[q][1], v[q][2], v[q][3]);
          *reinterpret_cast<float4*>(xc + 4) = make_float4(v[q][4], v[q][5], v[q][6], v[q][7]);
          float* xd = sv.XD + (size_t)row * D + lane * 8;
          *reinterpret_cast<float4*>(xd) = make_float4(u[0], u[1], u[2], u[3]);
          *reinterpret_cast<float4*>(xd + 4) = make_float4(u[4], u[5], u[6], u[7]);
          if (lane == 0) { sv.rstd_c[row] = rc; sv.rstd_d[row] = rd; }
        }
      }
    }
    __syncthreads();
    // h[r0 + r, j0 + c] = z[r, :] . W1[j0 + c, :] + b1
    {
      const int r = tid >> 3, c = tid & 7;
      const float4* zr = reinterpret_cast<const float4*>(sZ + r * kLd);
      const float4* wr = reinterpret_cast<const float4*>(sW + c * kLd);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
      for (int k = 0; k < Z / 4; ++k) {
        const float4 zv = zr[k], wv = wr[k];
        a0 = fmaf(zv.x, wv.x, a0); a1 = fmaf(zv.y, wv.y, a1); a2 = fmaf(zv.z, wv.z, a2); a3 = fmaf(zv.w, wv.w, a3);
      }
      if (r0 + r < B) sH[(r0 + r) * kCols + c] = (a0 + a1) + (a2 + a3) + __ldg(p.b1 + j0 + c);
    }
  }
  __syncthreads();
  // BatchNorm over the batch: warp w owns column j0 + w
  {
    const int j = j0 + warp;
    float s = 0.f;
    for (int r = lane; r < B; r += 32) s += sH[r * kCols + warp];
    const float mean = warp_sum(s) / (float)B;
    float q = 0.f;
    for (int r = lane; r < B; r += 32) { const float d = sH[r * kCols + warp] - mean; q = fmaf(d, d, q); }
    const float var = warp_sum(q) / (float)B;
    const float inv = rsqrtf(var + bn_eps);
    const float g = __ldg(p.bn_g + j), bb = __ldg(p.bn_b + j), w3 = __ldg(p.w3 + j);
    if (lane == 0) {
      sv.invstd[j] = inv;
      run_mean[j] = fmaf(momentum, mean - run_mean[j], run_mean[j]);
      const float unb = var * ((float)B / (float)(B - 1));
      run_var[j] = fmaf(momentum, unb - run_var[j], run_var[j]);
    }
    for (int r = lane; r < B; r += 32) {
      const float xh = (sH[r * kCols + warp] - mean) * inv;
      sv.XH[(size_t)r * D + j] = xh;
      sH[r * kCols + warp] = fmaxf(fmaf(xh, g, bb), 0.f) * w3;       // this column's term of the logit
    }
  }
  __syncthreads();
  for (int r = tid; r < B; r += kThreads) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kCols; ++c) s += sH[r * kCols + c];
    part[(size_t)blockIdx.x * B + r] = s;
  }
  if (last_block(counter, gridDim.x)) {
    const float b3 = __ldg(p.b3);
    for (int r = tid; r < B; r += kThreads) {
      float v[kGrid];
#pragma unroll
      for (int g = 0; g < kGrid; ++g) v[g] = __ldcg(part + (size_t)g * B + r);     // all loads in flight, then a fixed-order sum
      float s = b3;
#pragma unroll
      for (int g = 0; g < kGrid; ++g) s += v[g];
      logits[r] = s;
    }
    if (tid == 0 && nbt) *nbt += 1;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------------------
// backward 1: through Linear(256,1), ReLU, BatchNorm1d; weight / bias gradients of fc_list.0; dh -> DH [B, 256].
// grid = kGrid, CTA g owns columns [g*kCols, (g+1)*kCols). dynamic smem: sDH [B][kCols]
// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) head_bwd1_kernel(const float* __restrict__ dlogit, int B, HeadParams p,
                                                             HeadSaved sv, HeadGrads g, float* __restrict__ DH) {
  extern __shared__ __align__(16) float smem[];
  float* sDH = smem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j0 = blockIdx.x * kCols;
  {
    const int j = j0 + warp;
    const float gam = __ldg(p.bn_g + j), bet = __ldg(p.bn_b + j), w3 = __ldg(p.w3 + j), inv = sv.invstd[j];
    float s_w3 = 0.f, s_g = 0.f, s_b = 0.f, s_l = 0.f;
    for (int r = lane; r < B; r += 32) {
      const float xh = sv.XH[(size_t)r * D + j];
      const float y = fmaf(xh, gam, bet);
      const float dl = __ldg(dlogit + r);
      const float dy = y > 0.f ? dl * w3 : 0.f;
      s_w3 = fmaf(dl, fmaxf(y, 0.f), s_w3);
      s_g = fmaf(dy, xh, s_g);
      s_b += dy;
      s_l += dl;
    }
    s_w3 = warp_sum(s_w3); s_g = warp_sum(s_g); s_b = warp_sum(s_b); s_l = warp_sum(s_l);
    const float mb = s_b / (float)B, mg = s_g / (float)B, k = gam * inv;
    float s_h = 0.f;
    for (int r = lane; r < B; r += 32) {
      const float xh = sv.XH[(size_t)r * D + j];
      const float y = fmaf(xh, gam, bet);
      const float dy = y > 0.f ? __ldg(dlogit + r) * w3 : 0.f;
      const float dh = k * (dy - mb - xh * mg);
      sDH[r * kCols + warp] = dh;
      DH[(size_t)r * D + j] = dh;
      s_h += dh;
    }
    s_h = warp_sum(s_h);
    if (lane == 0) {
      g.w3[j] = s_w3; g.bn_g[j] = s_g; g.bn_b[j] = s_b; g.b1[j] = s_h;
      if (j == 0) g.b3[0] = s_l;
    }
  }
  __syncthreads();
  // dW1[j0 + c, k] = sum_b dh[b, c] z[b, k]   thread t: k = t and t + 256
  float acc[kCols][2];
#pragma unroll
  for (int c = 0; c < kCols; ++c) acc[c][0] = acc[c][1] = 0.f;
  for (int b0 = 0; b0 < B; b0 += 8) {      // 16 loads in flight per thread: the loop is L2-latency bound otherwise
    float z0[8], z1[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int b = b0 + u;
      z0[u] = b < B ? sv.Zs[(size_t)b * Z + tid] : 0.f;
      z1[u] = b < B ? sv.Zs[(size_t)b * Z + D + tid] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int b = b0 + u;
      if (b < B) {
        const float4 d0 = *reinterpret_cast<const float4*>(sDH + b * kCols);
        const float4 d1 = *reinterpret_cast<const float4*>(sDH + b * kCols + 4);
        const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
        for (int c = 0; c < kCols; ++c) { acc[c][0] = fmaf(d[c], z0[u], acc[c][0]); acc[c][1] = fmaf(d[c], z1[u], acc[c][1]); }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < kCols; ++c) {
    g.W1[(size_t)(j0 + c) * Z + tid] = acc[c][0];
    g.W1[(size_t)(j0 + c) * Z + D + tid] = acc[c][1];
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// backward 2: dz = dh W1; LayerNorm backward of both halves -> dcls, gradients of layer_norms_after_concat and ie_demo.
// CTA g takes row groups g, g + grid, ... of kRowsB2 rows; thread t owns feature t of both halves. Per-CTA partial sums of the
// 7 per-feature parameter gradients -> part [grid][7][256]; the last CTA adds them in order.
// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) head_bwd2_kernel(const float* __restrict__ DH, const float* __restrict__ age,
                                                             const float* __restrict__ gen, int B, HeadParams p,
                                                             HeadSaved sv, HeadGrads g, float* __restrict__ dcls,
                                                             float* __restrict__ part, unsigned int* __restrict__ counter) {
  __shared__ __align__(16) float sD[kRowsB2][D];
  __shared__ __align__(16) float sDZ[2][kRowsB2][Z];     // the two j-parity halves of dz
  __shared__ float sRed[kThreads / 32][kRowsB2][4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float lng = __ldg(p.ln_g + tid), ldg_ = __ldg(p.lnd_g + tid);
  float a_lng = 0.f, a_lnb = 0.f, a_dg = 0.f, a_db = 0.f, a_w0 = 0.f, a_w1 = 0.f, a_bd = 0.f;
  const int n_grp = (B + kRowsB2 - 1) / kRowsB2;
  for (int grp = blockIdx.x; grp < n_grp; grp += gridDim.x) {
    const int b0 = grp * kRowsB2;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRowsB2; ++r) sD[r][tid] = (b0 + r < B) ? DH[(size_t)(b0 + r) * D + tid] : 0.f;
    __syncthreads();
    // dz[r, :] = sum_j dh[r, j] W1[j, :]. Thread = (4 consecutive columns, parity of j): 128 float4 loads per thread instead of
    // 512 scalar ones -- the loop is bound by how many BYTES a thread keeps in flight against the L2 latency (ptxas holds ~12
    // loads in flight: 43 us with scalar loads and a rolled loop, 24 us unrolled).
    {
      const int c4 = (tid & 127) * 4, jh = tid >> 7;
      float a4[kRowsB2][4];
#pragma unroll
      for (int r = 0; r < kRowsB2; ++r) a4[r][0] = a4[r][1] = a4[r][2] = a4[r][3] = 0.f;
#pragma unroll 16
      for (int jj = 0; jj < D / 2; ++jj) {
        const int j = jj * 2 + jh;
        const float4 w = __ldg(reinterpret_cast<const float4*>(p.W1 + (size_t)j * Z + c4));
#pragma unroll
        for (int r = 0; r < kRowsB2; ++r) {
          const float d = sD[r][j];
          a4[r][0] = fmaf(d, w.x, a4[r][0]); a4[r][1] = fmaf(d, w.y, a4[r][1]);
          a4[r][2] = fmaf(d, w.z, a4[r][2]); a4[r][3] = fmaf(d, w.w, a4[r][3]);
        }
      }
#pragma unroll
      for (int r = 0; r < kRowsB2; ++r)
        *reinterpret_cast<float4*>(&sDZ[jh][r][c4]) = make_float4(a4[r][0], a4[r][1], a4[r][2], a4[r][3]);
    }
    __syncthreads();
    float acc[kRowsB2][2];
#pragma unroll
    for (int r = 0; r < kRowsB2; ++r) {
      acc[r][0] = sDZ[0][r][tid] + sDZ[1][r][tid];
      acc[r][1] = sDZ[0][r][D + tid] + sDZ[1][r][D + tid];
    }
    float dch[kRowsB2], ddh[kRowsB2], xc[kRowsB2], xd[kRowsB2];
#pragma unroll
    for (int r = 0; r < kRowsB2; ++r) {
      const int b = b0 + r;
      const bool ok = b < B;
      xc[r] = ok ? sv.XC[(size_t)b * D + tid] : 0.f;
      xd[r] = ok ? sv.XD[(size_t)b * D + tid] : 0.f;
      const float dzc = ok ? acc[r][0] : 0.f;
      const float dzd = (ok && sv.Zs[(size_t)b * Z + D + tid] > 0.f) ? acc[r][1] : 0.f;
      a_lng = fmaf(dzc, xc[r], a_lng); a_lnb += dzc;
      a_dg = fmaf(dzd, xd[r], a_dg); a_db += dzd;
      dch[r] = dzc * lng;
      ddh[r] = dzd * ldg_;
      const float s1 = warp_sum(dch[r]), s2 = warp_sum(dch[r] * xc[r]), s3 = warp_sum(ddh[r]), s4 = warp_sum(ddh[r] * xd[r]);
      if (lane == 0) { sRed[warp][r][0] = s1; sRed[warp][r][1] = s2; sRed[warp][r][2] = s3; sRed[warp][r][3] = s4; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRowsB2; ++r) {
      const int b = b0 + r;
      if (b >= B) continue;
      float s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) { s1 += sRed[w][r][0]; s2 += sRed[w][r][1]; s3 += sRed[w][r][2]; s4 += sRed[w][r][3]; }
      dcls[(size_t)b * D + tid] = sv.rstd_c[b] * (dch[r] - s1 * (1.f / D) - xc[r] * s2 * (1.f / D));
      const float du = sv.rstd_d[b] * (ddh[r] - s3 * (1.f / D) - xd[r] * s4 * (1.f / D));
      a_w0 = fmaf(du, __ldg(age + b), a_w0);
      a_w1 = fmaf(du, __ldg(gen + b), a_w1);
      a_bd += du;
    }
  }
  float* mine = part + (size_t)blockIdx.x * 7 * D;
  mine[0 * D + tid] = a_lng; mine[1 * D + tid] = a_lnb; mine[2 * D + tid] = a_dg; mine[3 * D + tid] = a_db;
  mine[4 * D + tid] = a_w0; mine[5 * D + tid] = a_w1; mine[6 * D + tid] = a_bd;
  if (last_block(counter, gridDim.x)) {
    float s[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (unsigned int b0 = 0; b0 < gridDim.x; b0 += 4) {      // 28 loads in flight; the order of the additions is fixed
      float v[4][7];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int q = 0; q < 7; ++q)
          v[u][q] = (b0 + u < gridDim.x) ? __ldcg(part + ((size_t)(b0 + u) * 7 + q) * D + tid) : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int q = 0; q < 7; ++q) s[q] += v[u][q];
    }
    g.ln_g[tid] = s[0]; g.ln_b[tid] = s[1]; g.lnd_g[tid] = s[2]; g.lnd_b[tid] = s[3];
    g.Wd[2 * tid] = s[4]; g.Wd[2 * tid + 1] = s[5]; g.bd[tid] = s[6];
  }
}

int bwd2_grid(int B) {
  const int n_grp = (B + kRowsB2 - 1) / kRowsB2;
  return n_grp < 64 ? n_grp : 64;
}

}  // namespace

// params: 12 device pointers in HeadParams order; saved: 7 device pointers in HeadSaved order (written here, read by
// tmp_head_bwd). scratch: >= (256 / 8) * B floats; counter: one zero-initialised uint32 (left at zero). run_mean / run_var
// [256] are updated in place, *nbt (int64 num_batches_tracked, may be NULL) is incremented. logits: [B].
extern "C" int tmp_head_fwd(const float* cls, const float* age, const float* gen, int B, const void* const* params,
                            float* run_mean, float* run_var, long long* nbt, float momentum, float bn_eps,
                            void* const* saved, float* scratch, unsigned int* counter, float* logits, void* stream) {
  TMP_REQUIRE(cls && age && gen && params && run_mean && run_var && saved && scratch && counter && logits,
              "head_fwd: null operand");
  TMP_REQUIRE(B >= 2 && B <= 4096, "head_fwd: batch statistics need 2 <= B <= 4096 (B=%d)", B);
  HeadParams p;
  const float** pp = reinterpret_cast<const float**>(&p);
  for (int i = 0; i < 12; ++i) {
    TMP_REQUIRE(params[i], "head_fwd: null parameter %d", i);
    pp[i] = (const float*)params[i];
  }
  HeadSaved sv;
  float** sp = reinterpret_cast<float**>(&sv);
  for (int i = 0; i < 7; ++i) {
    TMP_REQUIRE(saved[i], "head_fwd: null saved tensor %d", i);
    sp[i] = (float*)saved[i];
  }
  const size_t smem = (size_t)(kCols * kLd + kChunk * kLd + (size_t)B * kCols) * sizeof(float);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      tmp::set_error("cudaFuncSetAttribute(head_fwd): %s", cudaGetErrorString(e));
      return (int)e;
    }
    smem_set = smem;
  }
  head_fwd_kernel<<<kGrid, kThreads, smem, (cudaStream_t)stream>>>(cls, age, gen, B, p, run_mean, run_var, nbt, momentum, bn_eps,
                                                                  sv, scratch, counter, logits);
  return tmp::check_launch("head_fwd_kernel");
}

// dlogit [B]; grads: 12 device pointers in HeadParams order (overwritten, not accumulated); dcls [B, 256];
// DH: scratch [B, 256]; scratch: >= 64 * 7 * 256 floats; counter as in tmp_head_fwd. Two launches.
extern "C" int tmp_head_bwd(const float* dlogit, const float* age, const float* gen, int B, const void* const* params,
                            void* const* saved, void* const* grads, float* dcls, float* DH, float* scratch,
                            unsigned int* counter, void* stream) {
  TMP_REQUIRE(dlogit && age && gen && params && saved && grads && dcls && DH && scratch && counter, "head_bwd: null operand");
  TMP_REQUIRE(B >= 2 && B <= 4096, "head_bwd: 2 <= B <= 4096 (B=%d)", B);
  HeadParams p;
  const float** pp = reinterpret_cast<const float**>(&p);
  HeadGrads g;
  float** gp = reinterpret_cast<float**>(&g);
  for (int i = 0; i < 12; ++i) {
    TMP_REQUIRE(params[i] && grads[i], "head_bwd: null parameter / gradient %d", i);
    pp[i] = (const float*)params[i];
    gp[i] = (float*)grads[i];
  }
  HeadSaved sv;
  float** sp = reinterpret_cast<float**>(&sv);
  for (int i = 0; i < 7; ++i) {
    TMP_REQUIRE(saved[i], "head_bwd: null saved tensor %d", i);
    sp[i] = (float*)saved[i];
  }
  const size_t smem = (size_t)B * kCols * sizeof(float);
  static size_t smem_set = 48 * 1024;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(head_bwd1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      tmp::set_error("cudaFuncSetAttribute(head_bwd1): %s", cudaGetErrorString(e));
      return (int)e;
    }
    smem_set = smem;
  }
  head_bwd1_kernel<<<kGrid, kThreads, smem, (cudaStream_t)stream>>>(dlogit, B, p, sv, g, DH);
  int rc = tmp::check_launch("head_bwd1_kernel");
  if (rc) return rc;
  head_bwd2_kernel<<<bwd2_grid(B), kThreads, 0, (cudaStream_t)stream>>>(DH, age, gen, B, p, sv, g, dcls, scratch, counter);
  return tmp::check_launch("head_bwd2_kernel");
}
