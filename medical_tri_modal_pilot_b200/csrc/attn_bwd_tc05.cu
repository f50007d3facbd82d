// attn_bwd_tc05.cu -- backward of the modality-aware attention (reference attention.py:24-49 under autograd)
// as a tcgen05/TMEM kernel. One CTA owns a 128-key tile of one (sample, head) and sweeps the live query tiles:
//
//   S^T  = K_j Q_i^T            dP^T = V_j dO_i^T                      (2 MMAs into TMEM)
//   P^T  = exp2(S^T*c - LSE_i)  dS^T = P^T o (dP^T - delta_i) / 8      (registers -> swizzled smem, fp16)
//   dV_j += P^T dO_i            dK_j += dS^T Q_i       dQ_i = dS K_j   (3 MMAs; dQ -> fp32 atomics)
//
// dS^T is written to shared memory once and read twice: as a K-major A operand (dK) and as an MN-major
// A operand (dQ = dS.K needs the transpose). Q_i / dO_i / K_j are consumed as MN-major B operands straight
// from their natural [row, d] layout. The key-padding mask is kv_len[b] applied in-register (P^T rows of
// masked keys are exactly 0, so dK/dV of pad rows are exactly 0, as in the reference).
#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int kThreads = 192;
constexpr int BT = 128;  // tile rows (keys per CTA, queries per iteration)
constexpr int HD = 64;
constexpr int kStages = 2;
constexpr int kTile = BT * HD * 2;   // 16 KB : [128 rows x 64 bf16]
constexpr int kSq = BT * BT * 2;     // 32 KB : [128 x 128] bf16 as two 64-column sub-tiles

constexpr int kSmemK = 0;
constexpr int kSmemV = kSmemK + kTile;
constexpr int kSmemQ = kSmemV + kTile;
constexpr int kSmemDO = kSmemQ + kStages * kTile;
constexpr int kSmemPT = kSmemDO + kStages * kTile;
constexpr int kSmemDST = kSmemPT + kSq;
constexpr int kSmemStat = kSmemDST + kSq;                 // [2 buffers][lse 128 | delta 128] fp32
constexpr int kSmemBar = kSmemStat + 2 * 2 * BT * 4;
constexpr int kSmemTotal = kSmemBar + 128 + 1024;

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Bars {
  uint64_t kv_full;
  uint64_t qdo_full[kStages], qdo_empty[kStages];
  uint64_t sdp_full;  // MMA -> compute : S^T and dP^T in TMEM
  uint64_t pds_full;  // compute -> MMA : P^T and dS^T in smem
  uint64_t dq_full;   // MMA -> compute : dQ tile in TMEM, all MMAs of this iteration retired
  uint32_t tmem_slot;
};

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const int32_t* __restrict__ kv_len, int T, const float* __restrict__ lse2,
                const float* __restrict__ delta, int T_lse, float* __restrict__ dQ_acc, uint16_t* __restrict__ dQKV,
                float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  Bars* bars = (Bars*)(smem + kSmemBar);
  float* stat = (float*)(smem + kSmemStat);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int jt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int H = gridDim.y;
  const int k0 = jt * BT;
  const int len = kv_len ? min(kv_len[b], T) : T;
  const int row_base = b * T;

  if (k0 >= len) {
    // masked / padding keys: dK = dV = 0
    if (warp >= 2) {
      const int r = (warp - 2) * 32 + lane;
      if (k0 + r < T) {
        uint4* dk = reinterpret_cast<uint4*>(dQKV + (size_t)(row_base + k0 + r) * 768 + 256 + h * HD);
        uint4* dv = reinterpret_cast<uint4*>(dQKV + (size_t)(row_base + k0 + r) * 768 + 512 + h * HD);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          dk[i] = make_uint4(0, 0, 0, 0);
          dv[i] = make_uint4(0, 0, 0, 0);
        }
      }
    }
    return;
  }
  const int n_q = (len + BT - 1) / BT;  // live query tiles (query rows >= len are padding: dO == 0)

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQKV);
    prefetch_tmap(&tmDO);
    mbar_init(&bars->kv_full, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars->qdo_full[s], 1);
      mbar_init(&bars->qdo_empty[s], 1);
    }
    mbar_init(&bars->sdp_full, 1);
    mbar_init(&bars->pds_full, 4);
    mbar_init(&bars->dq_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  const uint32_t tm_ST = tmem_base + 0;
  const uint32_t tm_DPT = tmem_base + 128;
  const uint32_t tm_DV = tmem_base + 256;
  const uint32_t tm_DK = tmem_base + 320;
  const uint32_t tm_DQ = tmem_base + 384;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&bars->kv_full, 2 * kTile);
      tma_load_2d(smem + kSmemK, &tmQKV, &bars->kv_full, 256 + h * HD, row_base + k0);
      tma_load_2d(smem + kSmemV, &tmQKV, &bars->kv_full, 512 + h * HD, row_base + k0);
      int st = 0;
      uint32_t ph = 0;
      for (int i = 0; i < n_q; ++i) {
        mbar_wait(&bars->qdo_empty[st], ph ^ 1);
        mbar_expect_tx(&bars->qdo_full[st], 2 * kTile);
        tma_load_2d(smem + kSmemQ + st * kTile, &tmQKV, &bars->qdo_full[st], h * HD, row_base + i * BT);
        tma_load_2d(smem + kSmemDO + st * kTile, &tmDO, &bars->qdo_full[st], h * HD, row_base + i * BT);
        if (++st == kStages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // all operands fp16 (kind::f16 needs A and B in the same format; gradients are fp16 with a host-side scale)
      constexpr uint32_t idesc_st = make_idesc(BT, BT, 0, 0, FMT_F16, FMT_F16);  // S^T  = K Q^T
      constexpr uint32_t idesc_dp = make_idesc(BT, BT, 0, 0, FMT_F16, FMT_F16);  // dP^T = V dO^T
      constexpr uint32_t idesc_dv = make_idesc(BT, HD, 0, 1, FMT_F16, FMT_F16);  // dV  += P^T dO   (B MN-major)
      constexpr uint32_t idesc_dk = make_idesc(BT, HD, 0, 1, FMT_F16, FMT_F16);  // dK  += dS^T Q   (B MN-major)
      constexpr uint32_t idesc_dq = make_idesc(BT, HD, 1, 1, FMT_F16, FMT_F16);  // dQ   = dS K     (A, B MN-major)
      const uint32_t sK = smem_u32(smem + kSmemK), sV = smem_u32(smem + kSmemV);
      const uint32_t sPT = smem_u32(smem + kSmemPT), sDST = smem_u32(smem + kSmemDST);
      mbar_wait(&bars->kv_full, 0);
      int st = 0;
      uint32_t ph = 0;
      for (int i = 0; i < n_q; ++i) {
        const uint32_t sQ = smem_u32(smem + kSmemQ + st * kTile);
        const uint32_t sDO = smem_u32(smem + kSmemDO + st * kTile);
        mbar_wait(&bars->qdo_full[st], ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_ss(tm_ST, make_sdesc_sw128(sK + k * 32, 16, 1024), make_sdesc_sw128(sQ + k * 32, 16, 1024), idesc_st,
                  k != 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_ss(tm_DPT, make_sdesc_sw128(sV + k * 32, 16, 1024), make_sdesc_sw128(sDO + k * 32, 16, 1024), idesc_dp,
                  k != 0);
        umma_commit(&bars->sdp_full);

        mbar_wait(&bars->pds_full, i & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BT / 16; ++k) {  // dV += P^T dO_i   (reduction over the 128 queries)
          const uint64_t adesc = make_sdesc_sw128(sPT + (k >> 2) * (BT * 128) + (k & 3) * 32, 16, 1024);
          umma_ss(tm_DV, adesc, make_sdesc_sw128(sDO + k * 2048, BT * 128, 1024), idesc_dv, (i | k) != 0);
        }
#pragma unroll
        for (int k = 0; k < BT / 16; ++k) {  // dK += dS^T Q_i
          const uint64_t adesc = make_sdesc_sw128(sDST + (k >> 2) * (BT * 128) + (k & 3) * 32, 16, 1024);
          umma_ss(tm_DK, adesc, make_sdesc_sw128(sQ + k * 2048, BT * 128, 1024), idesc_dk, (i | k) != 0);
        }
#pragma unroll
        for (int k = 0; k < BT / 16; ++k) {  // dQ_i = dS K_j    (reduction over the 128 keys; A = dS^T read MN-major)
          umma_ss(tm_DQ, make_sdesc_sw128(sDST + k * 2048, BT * 128, 1024),
                  make_sdesc_sw128(sK + k * 2048, BT * 128, 1024), idesc_dq, k != 0);
        }
        umma_commit(&bars->dq_full);
        umma_commit(&bars->qdo_empty[st]);
        if (++st == kStages) { st = 0; ph ^= 1; }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // TMEM lane: key row for S^T/dP^T/dK/dV, query row for dQ
    const int tid = (warp - 2) * 32 + lane;
    const bool key_ok = (k0 + r) < len;
    uint8_t* sPT = smem + kSmemPT;
    uint8_t* sDST = smem + kSmemDST;
    const size_t stat_base = ((size_t)b * H + h) * T_lse;
    for (int i = 0; i < n_q; ++i) {
      float* st_lse = stat + (i & 1) * 2 * BT;
      float* st_dl = st_lse + BT;
      st_lse[tid] = lse2[stat_base + i * BT + tid];
      st_dl[tid] = delta[stat_base + i * BT + tid];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&bars->sdp_full, i & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {  // 32 queries per chunk
        uint32_t s[32], dp[32];
        tmem_ld32(tmem_addr(tm_ST, quarter * 32, c * 32), s);
        tmem_ld32(tmem_addr(tm_DPT, quarter * 32, c * 32), dp);
        tmem_ld_wait();
        uint32_t pp[16], dd[16];
#pragma unroll
        for (int t = 0; t < 32; t += 2) {
          float p[2], d[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int qc = c * 32 + t + u;
            const bool ok = key_ok && (i * BT + qc) < len;
            const float pv = ex2_approx(fmaf(__uint_as_float(s[t + u]), scale_log2, -st_lse[qc]));
            p[u] = ok ? pv : 0.f;
            d[u] = ok ? pv * (__uint_as_float(dp[t + u]) - st_dl[qc]) * 0.125f : 0.f;
          }
          pp[t >> 1] = pack_f16x2(p[0], p[1]);
          dd[t >> 1] = pack_f16x2(d[0], d[1]);
        }
        const uint32_t sub = (c >> 1) * (BT * 128);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint32_t off = sub + sw128_offset(r, (c & 1) * 4 + q4);
          *reinterpret_cast<uint4*>(sPT + off) = make_uint4(pp[q4 * 4], pp[q4 * 4 + 1], pp[q4 * 4 + 2], pp[q4 * 4 + 3]);
          *reinterpret_cast<uint4*>(sDST + off) = make_uint4(dd[q4 * 4], dd[q4 * 4 + 1], dd[q4 * 4 + 2], dd[q4 * 4 + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->pds_full);

      mbar_wait(&bars->dq_full, i & 1);
      tc_fence_after();
      const int q = i * BT + r;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_addr(tm_DQ, quarter * 32, c * 32), v);
        tmem_ld_wait();
        if (q < len) {
          float* dst = dQ_acc + (size_t)(row_base + q) * 256 + h * HD + c * 32;
#pragma unroll
          for (int t = 0; t < 32; t += 4)
            red_add_v4(dst + t, __uint_as_float(v[t]), __uint_as_float(v[t + 1]), __uint_as_float(v[t + 2]),
                       __uint_as_float(v[t + 3]));
        }
      }
      tc_fence_before();
    }
    // dK_j, dV_j (all MMAs retired: last dq_full)
    const int kr = k0 + r;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const uint32_t src = which == 0 ? tm_DK : tm_DV;
      uint16_t* dst_row = dQKV + (size_t)(row_base + kr) * 768 + (which == 0 ? 256 : 512) + h * HD;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_addr(src, quarter * 32, c * 32), v);
        tmem_ld_wait();
        if (kr < T) {
          uint4* dst = reinterpret_cast<uint4*>(dst_row + c * 32);
#pragma unroll
          for (int t = 0; t < 4; ++t)
            dst[t] = make_uint4(pack_f16x2(__uint_as_float(v[t * 8 + 0]), __uint_as_float(v[t * 8 + 1])),
                                pack_f16x2(__uint_as_float(v[t * 8 + 2]), __uint_as_float(v[t * 8 + 3])),
                                pack_f16x2(__uint_as_float(v[t * 8 + 4]), __uint_as_float(v[t * 8 + 5])),
                                pack_f16x2(__uint_as_float(v[t * 8 + 6]), __uint_as_float(v[t * 8 + 7])));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
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
extern "C" int tmp_mma_attn_bwd(const void* qkv, const void* O, const void* dO, int ld, const int32_t* kv_len, int B,
                                int T, int H, const float* lse2, int T_lse, float* delta, float* dQ_acc, void* dQKV,
                                void* stream) {
  TMP_REQUIRE(qkv && O && dO && lse2 && delta && dQ_acc && dQKV, "attn_bwd: null operand");
  TMP_REQUIRE(B > 0 && T > 0 && H == 4 && ld == 256, "attn_bwd: need H==4, ld==256 (B=%d T=%d H=%d ld=%d)", B, T, H, ld);
  TMP_REQUIRE(T_lse % BT == 0 && T_lse >= T, "attn_bwd: T_lse must be a multiple of 128 and >= T");
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (e != cudaSuccess) {
      tmp::set_error("cudaFuncSetAttribute(attn_bwd): %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  CUtensorMap tmQKV, tmDO;
  int rc = tmp::encode_tmap_2d_bf16(&tmQKV, qkv, 768, (uint64_t)B * T, 768 * 2, HD, BT);
  if (rc) return rc;
  rc = tmp::encode_tmap_2d_bf16(&tmDO, dO, 256, (uint64_t)B * T, (uint64_t)ld * 2, HD, BT);
  if (rc) return rc;
  {
    const int rows = B * T_lse;
    attn_bwd_delta_kernel<<<(rows + 7) / 8, 256, 0, st>>>((const uint16_t*)O, (const uint16_t*)dO, ld, B, T, H, delta, T_lse);
    rc = tmp::check_launch("attn_bwd_delta_kernel");
    if (rc) return rc;
  }
  cudaError_t e = cudaMemsetAsync(dQ_acc, 0, (size_t)B * T * 256 * sizeof(float), st);
  if (e != cudaSuccess) {
    tmp::set_error("attn_bwd memset: %s", cudaGetErrorString(e));
    return (int)e;
  }
  dim3 grid((T + BT - 1) / BT, H, B);
  attn_bwd_kernel<<<grid, kThreads, kSmemTotal, st>>>(tmQKV, tmDO, kv_len, T, lse2, delta, T_lse, dQ_acc, (uint16_t*)dQKV,
                                                      kLog2e / 8.0f);
  rc = tmp::check_launch("attn_bwd_kernel");
  if (rc) return rc;
  const size_t rows = (size_t)B * T;
  attn_bwd_dq_convert_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(dQ_acc, (uint16_t*)dQKV, rows);
  return tmp::check_launch("attn_bwd_dq_convert_kernel");
}
