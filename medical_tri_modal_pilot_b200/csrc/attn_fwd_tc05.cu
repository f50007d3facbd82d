// attn_fwd_tc05.cu -- modality-aware attention forward (MultiHeadAttention.forward +
// ScaledDotProductAttention.forward, reference attention.py:24-49,65-84) as a flash-style tcgen05/TMEM kernel.
//
//   O[b, q, h*64:(h+1)*64] = softmax_k( Q K^T / 8  masked to k < kv_len[b] ) V          (no out-projection)
//
// Inputs come straight from the fused QKV GEMM: qkv[B*T, 768] fp16 (Q | K | V, head h at columns h*64).
// The reference's bool [B*H,T,T] key-padding mask (attention.py:38, utils.py:116-125) is never materialised:
// kv_len[b] (4 bottleneck keys + CLS + valid tokens) is applied in-register, KV tiles past it are skipped,
// query tiles past it (pad rows, provably dead) are written as zeros.
//
// One CTA = 128 query rows of one (sample, head). warp0 = TMA, warp1 = MMA issuer, warps2-5 = softmax.
// S (128x128 fp32), O (128x64 fp32) and P (128x128 fp16, 64 columns) live in TMEM: P goes registers -> tcgen05.st ->
// A operand of P.V read from TMEM (no shared-memory round trip: the SS form of the N=64 MMA needs 192 B/clk of smem
// reads, above the 128 B/clk the SM has); the 32 KB this frees hold a third K/V ring stage (a stage is only released
// when its MMA retires, so with two stages every tile waited a full TMA round trip). V is consumed as an MN-major B
// operand directly from its natural [kv, d] layout. S(j+1) = Q K^T is issued as soon as S(j) has been copied to
// registers, i.e. it runs under the exponentials of tile j.
#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int kThreads = 192;
constexpr int BQ = 128;   // query rows per CTA
constexpr int BKV = 128;  // keys per iteration
constexpr int HD = 64;    // head dim
constexpr int kKVStages = 3;

constexpr int kQBytes = BQ * HD * 2;       // 16 KB
constexpr int kKBytes = BKV * HD * 2;      // 16 KB
constexpr int kSmemQ = 0;
constexpr int kSmemK = kSmemQ + kQBytes;
constexpr int kSmemV = kSmemK + kKVStages * kKBytes;
constexpr int kSmemBar = kSmemV + kKVStages * kKBytes;
constexpr int kSmemTotal = kSmemBar + 128 + 896;  // barriers + alignment slack

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Bars {
  uint64_t q_full;
  uint64_t k_full[kKVStages], k_empty[kKVStages];
  uint64_t v_full[kKVStages], v_empty[kKVStages];
  uint64_t s_full;   // MMA -> softmax : S tile ready in TMEM
  uint64_t s_free;   // softmax -> MMA : S tile copied to registers (TMEM S reusable: S(j+1) may be issued)
  uint64_t p_full;   // softmax -> MMA : P in TMEM (and O rescaled)
  uint64_t o_done;   // MMA -> softmax : P.V retired (P / O TMEM reusable)
  uint32_t tmem_slot;
};

__global__ void __launch_bounds__(kThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                const int32_t* __restrict__ kv_len, int T, int ld_o,
                uint16_t* __restrict__ O, float* __restrict__ lse2, int T_lse, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  Bars* bars = (Bars*)(smem + kSmemBar);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * BQ;
  const int len = kv_len ? min(kv_len[b], T) : T;
  const int row_base = b * T;  // first row of this sample in the [B*T, 768] qkv matrix

  if (q0 >= len) {
    // dead (padding) query rows: defined zeros, never consumed by live rows
    if (warp >= 2) {
      const int r = (warp - 2) * 32 + lane;
      if (q0 + r < T) {
        uint4* o = reinterpret_cast<uint4*>(O + (size_t)(row_base + q0 + r) * ld_o + h * HD);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = make_uint4(0, 0, 0, 0);
        lse2[((size_t)b * gridDim.y + h) * T_lse + q0 + r] = 0.f;
      }
    }
    return;
  }
  const int n_kv = (len + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQKV);
    prefetch_tmap(&tmO);
    mbar_init(&bars->q_full, 1);
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(&bars->k_full[s], 1);
      mbar_init(&bars->k_empty[s], 1);
      mbar_init(&bars->v_full[s], 1);
      mbar_init(&bars->v_empty[s], 1);
    }
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->s_free, 4);
    mbar_init(&bars->p_full, 4);
    mbar_init(&bars->o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  const uint32_t tmem_S = tmem_base;        // columns [0,128)
  const uint32_t tmem_O = tmem_base + 128;  // columns [128,192)
  const uint32_t tmem_P = tmem_base + 192;  // columns [192,256): 128 fp16 per lane, two per column

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(&bars->q_full, kQBytes);
      tma_load_2d(smem + kSmemQ, &tmQKV, &bars->q_full, h * HD, row_base + q0);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&bars->k_empty[st], ph ^ 1);
        mbar_expect_tx(&bars->k_full[st], kKBytes);
        tma_load_2d(smem + kSmemK + st * kKBytes, &tmQKV, &bars->k_full[st], 256 + h * HD, row_base + j * BKV);
        mbar_wait(&bars->v_empty[st], ph ^ 1);
        mbar_expect_tx(&bars->v_full[st], kKBytes);
        tma_load_2d(smem + kSmemV + st * kKBytes, &tmQKV, &bars->v_full[st], 512 + h * HD, row_base + j * BKV);
        if (++st == kKVStages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc(BQ, BKV, 0, 0, FMT_F16, FMT_F16);  // S = Q K^T, both K-major
      constexpr uint32_t idesc_o = make_idesc(BQ, HD, 0, 1, FMT_F16, FMT_F16);   // O += P V, V MN-major
      const uint32_t sQ = smem_u32(smem + kSmemQ);
      mbar_wait(&bars->q_full, 0);
      // prologue: S(0)
      mbar_wait(&bars->k_full[0], 0);
      tc_fence_after();
      {
        const uint32_t sK = smem_u32(smem + kSmemK);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_ss(tmem_S, make_sdesc_sw128(sQ + k * 32, 16, 1024), make_sdesc_sw128(sK + k * 32, 16, 1024), idesc_s,
                  k != 0);
        umma_commit(&bars->s_full);
        umma_commit(&bars->k_empty[0]);
      }
      int stK = 1 % kKVStages, stV = 0;       // ring stage of K(j+1) / of V(j)
      uint32_t phK = 0, phV = 0;
      for (int j = 0; j < n_kv; ++j) {
        // S(j+1) = Q K(j+1)^T as soon as the softmax warps have copied S(j) out of TMEM: the QK^T MMA runs under
        // the exponentials of tile j instead of after them
        if (j + 1 < n_kv) {
          mbar_wait(&bars->s_free, j & 1);
          mbar_wait(&bars->k_full[stK], phK);
          tc_fence_after();
          const uint32_t sK = smem_u32(smem + kSmemK + stK * kKBytes);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_ss(tmem_S, make_sdesc_sw128(sQ + k * 32, 16, 1024), make_sdesc_sw128(sK + k * 32, 16, 1024),
                    idesc_s, k != 0);
          umma_commit(&bars->s_full);
          umma_commit(&bars->k_empty[stK]);
          if (++stK == kKVStages) { stK = 0; phK ^= 1; }
        }
        // O += P(j) V(j): A = P from TMEM (16 keys = 8 columns per K step), B = V MN-major from smem
        mbar_wait(&bars->p_full, j & 1);
        mbar_wait(&bars->v_full[stV], phV);
        tc_fence_after();
        const uint32_t sV = smem_u32(smem + kSmemV + stV * kKBytes);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          umma_ts(tmem_O, tmem_P + k * 8, make_sdesc_sw128(sV + k * 2048, BKV * 128, 1024), idesc_o, (j | k) != 0);
        umma_commit(&bars->o_done);
        umma_commit(&bars->v_empty[stV]);
        if (++stV == kKVStages) { stV = 0; phV ^= 1; }
      }
    }
  } else {
    // ===================== softmax warps: thread <-> query row =====================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t t_S = tmem_addr(tmem_S, quarter * 32, 0);
    const uint32_t t_O = tmem_addr(tmem_O, quarter * 32, 0);
    const uint32_t t_P = tmem_addr(tmem_P, quarter * 32, 0);
    float m_ref = -INFINITY;  // running reference max (raw score units)
    float l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&bars->s_full, j & 1);
      tc_fence_after();
      uint32_t s[128];
      {
        uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
        uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
        uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
        uint32_t(&s3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
        tmem_ld32(t_S + 0, s0);
        tmem_ld32(t_S + 32, s1);
        tmem_ld32(t_S + 64, s2);
        tmem_ld32(t_S + 96, s3);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_free);
      const int kbase = j * BKV;
      if (kbase + BKV > len) {
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (kbase + c >= len) s[c] = 0xff800000u;  // -inf
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains
#pragma unroll
      for (int c = 0; c < 128; c += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) mx4[u] = fmaxf(mx4[u], __uint_as_float(s[c + u]));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // lazy rescale: keep the old reference max unless the new max exceeds it by > 8 (log2 units)
      float m_new = m_ref;
      const bool bump = (mx - m_ref) * scale_log2 > 8.f;  // also true on the first tile (m_ref = -inf)
      if (bump) m_new = mx;
      const float alpha = ex2_approx((m_ref - m_new) * scale_log2);  // 0 on first tile, 1 if no bump
      const float moff = m_new * scale_log2;
      l *= alpha;
      m_ref = m_new;
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int hblk = 0; hblk < 2; ++hblk) {  // two runs of 64 keys = 32 packed columns each
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 64; c += 2) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(s[hblk * 64 + c]), scale_log2, -moff));
          const float p1 = ex2_approx(fmaf(__uint_as_float(s[hblk * 64 + c + 1]), scale_log2, -moff));
          rs4[c & 3] += p0;
          rs4[(c & 3) + 1] += p1;
          pk[c >> 1] = pack_f16x2(p0, p1);
        }
        // P and O TMEM are free once P.V(j-1) has retired; the first 64 exponentials above do not need them, so the
        // wait sits here, after them
        if (hblk == 0 && j > 0) {
          mbar_wait(&bars->o_done, (j - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, bump)) {
            uint32_t o[32];
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
              tmem_ld32(t_O + c * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st32(t_O + c * 32, o);
            }
            tmem_st_wait();
          }
        }
        tmem_st32(t_P + hblk * 32, pk);
      }
      tmem_st_wait();
      l += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full);
    }
    // epilogue: O / l -> fp16. A full query tile is staged as one swizzled [32 rows x 128 B] box per warp in the Q buffer
    // (dead: the last S MMA retired before this warp read S of the last tile) and leaves through a TMA store; 16 B
    // stores from 32 threads to 32 different rows per instruction (32 partial sectors each) were ~10 % of the kernel.
    mbar_wait(&bars->o_done, (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l;
    const int q = q0 + r;
    const bool full_tile = q0 + BQ <= T;   // the 2-D map cannot clip at the sample boundary: the last tile goes by hand
    uint8_t* sO = smem + kSmemQ + quarter * 4096;
    uint32_t o[32];
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      tmem_ld32(t_O + c * 32, o);
      tmem_ld_wait();
      uint4 u[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        u[i].x = pack_f16x2(__uint_as_float(o[i * 8 + 0]) * inv_l, __uint_as_float(o[i * 8 + 1]) * inv_l);
        u[i].y = pack_f16x2(__uint_as_float(o[i * 8 + 2]) * inv_l, __uint_as_float(o[i * 8 + 3]) * inv_l);
        u[i].z = pack_f16x2(__uint_as_float(o[i * 8 + 4]) * inv_l, __uint_as_float(o[i * 8 + 5]) * inv_l);
        u[i].w = pack_f16x2(__uint_as_float(o[i * 8 + 6]) * inv_l, __uint_as_float(o[i * 8 + 7]) * inv_l);
      }
      if (full_tile) {
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(sO + sw128_offset(lane, c * 4 + i)) = u[i];
      } else if (q < T) {
        uint4* dst = reinterpret_cast<uint4*>(O + (size_t)(row_base + q) * ld_o + h * HD + c * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = u[i];
      }
    }
    if (full_tile) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tmO, sO, h * HD, row_base + q0 + quarter * 32);
        tma_store_commit();
        tma_store_wait_read0();   // the box must stay valid until the copy engine has read it
      }
      __syncwarp();
    }
    if (q < T) lse2[((size_t)b * gridDim.y + h) * T_lse + q] = m_ref * scale_log2 + log2f(l);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

// qkv: [B*T, 768] fp16.  kv_len: [B] int32 or NULL (= unmasked, reference `mask=None`).
// O: [B*T, ld_o] fp16 (head h at columns h*64).  lse2: [B, H, T_lse] fp32, log2-domain logsumexp of the scaled scores.
// q_rows: only the leading q_rows query rows of every sample are computed (rounded up to whole 128-row tiles; T = all).
// The last fused layer of `--mbt-only-vslt 1` only consumes the CLS row (row 4): its O / lse rows past the first tile are
// never read, so they are not produced.
extern "C" int tmp_mma_attn_fwd(const void* qkv, const int32_t* kv_len, int B, int T, int H, void* O, int ld_o,
                                float* lse2, int T_lse, int q_rows, void* stream) {
  TMP_REQUIRE(qkv && O && lse2, "attn_fwd: null operand");
  TMP_REQUIRE(B > 0 && T > 0 && H == 4, "attn_fwd: need B>0,T>0,H==4 (B=%d T=%d H=%d)", B, T, H);
  TMP_REQUIRE(T_lse >= T && ld_o % 8 == 0, "attn_fwd: T_lse >= T and ld_o %% 8 == 0 required");
  TMP_REQUIRE(q_rows > 0 && q_rows <= T, "attn_fwd: q_rows must be in [1, T]");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (e != cudaSuccess) {
      tmp::set_error("cudaFuncSetAttribute(attn_fwd): %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  CUtensorMap tm, tmO;
  int rc = tmp::encode_tmap_2d_h16(&tm, qkv, 768, (uint64_t)B * T, 768 * 2, HD, BQ);
  if (rc) return rc;
  rc = tmp::encode_tmap_2d_h16(&tmO, O, (uint64_t)ld_o, (uint64_t)B * T, (uint64_t)ld_o * 2, 64, 32);   // O boxes [32 rows x 64 cols]
  if (rc) return rc;
  dim3 grid((q_rows + BQ - 1) / BQ, H, B);
  const float scale_log2 = kLog2e / 8.0f;  // 1/sqrt(d_head=64) in log2 units (attention.py:16,35)
  attn_fwd_kernel<<<grid, kThreads, kSmemTotal, (cudaStream_t)stream>>>(tm, tmO, kv_len, T, ld_o, (uint16_t*)O, lse2, T_lse,
                                                                          scale_log2);
  return tmp::check_launch("attn_fwd_kernel");
}
