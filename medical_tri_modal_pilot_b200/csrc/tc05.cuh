// tc05.cuh -- thin inline-PTX layer for Blackwell (sm_100a): mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / st) and the UMMA shared-memory + instruction descriptors.
// Everything in this repo that touches the 5th-gen tensor cores goes through these wrappers.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc05 {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
// One lane of a CONVERGED warp. ptxas knows that exactly one thread runs the guarded region, so instructions that take
// uniform-register operands (tcgen05.mma descriptors, TMA, commit) are issued directly; behind `if (lane == 0)` it
// cannot prove that and wraps every such instruction in an ELECT / BRA.U.ANY "waterfall" loop (~19 SASS instructions
// and ~100 cycles per tcgen05.mma, tools/microbench/ub_mma.cu) -- slower than the 32-cycle N=64 MMAs it feeds.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// 1024 B-aligned base inside the dynamic shared-memory window (128B-swizzle atoms need it). Written as an OFFSET from
// the `extern __shared__` symbol: casting through uintptr_t made nvcc lose the address space and emit generic
// LD.E / ST.E (slower, "lg" stalls) for every smem access of the compute warps instead of LDS / STS.
__device__ __forceinline__ uint8_t* align_smem_1024(uint8_t* smem_raw) {
  return smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy writes to smem -> visible to the async proxy (TMA store / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch fails with an error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && (clock64() - t0) > 4000000000LL) {
      printf("tc05: mbarrier timeout block=(%d,%d,%d) thread=%d bar=%u parity=%u\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// TMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// 2D tiled load: c0 = innermost (contiguous) coordinate, c1 = row coordinate.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// 1D bulk copy global -> shared, completion on an mbarrier (bytes % 16 == 0, both addresses 16 B aligned)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation
// ------------------------------------------------------------------------------------------------
// Called by ONE full warp. ncols: power of two in [32, 512]. Result (TMEM base address) lands in *slot (smem).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// tcgen05: descriptors
// ------------------------------------------------------------------------------------------------
// Instruction descriptor, kind::f16, BF16 x BF16 -> FP32 accumulate, dense.
//   bits[4,6)=c_format(1=F32) [7,10)=a_format(1=BF16) [10,13)=b_format(1=BF16)
//   bit15=a_major (0=K,1=MN)  bit16=b_major  [17,23)=N>>3  [24,29)=M>>4
//   a_fmt / b_fmt: 0 = F16, 1 = BF16 (the step uses fp16 for forward AND gradient tensors -- see DESIGN.md "Precision
//   modes"; bf16 blocks only in the bf16x3 split of the fp32 mode).
enum : int { FMT_F16 = 0, FMT_BF16 = 1, FMT_F32 = 2 };   // FMT_F32: storage format of the fp32 ("precise") mode only
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major, int a_fmt, int b_fmt) {
  return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Shared-memory matrix descriptor for 128B-swizzled operand tiles (rows of 128 bytes, 8-row / 1024 B atoms).
//   bits[0,14)=addr>>4  [16,30)=LBO>>4  [32,46)=SBO>>4  [46,48)=version(1 on sm_100)  [61,64)=layout (2=SW128)
// K-major  : tile rows = M/N index, 64 bf16 of K per 128 B row; SBO = 1024 (next 8 rows); LBO unused (1).
// MN-major : tile rows = K index, 64 bf16 of M/N per 128 B row; SBO = 1024 (next 8 K rows);
//            LBO = byte distance to the next 64-wide M/N chunk.
__device__ __forceinline__ uint64_t make_sdesc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// ------------------------------------------------------------------------------------------------
// tcgen05: mma / commit   (issued by ONE thread)
// ------------------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives (count 1) when all tcgen05 ops previously issued by this thread have completed.
// Implies tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05: TMEM <-> registers.  32x32b shape: thread t of the warp <-> TMEM lane (lane_base + t),
// N consecutive 32-bit columns. A warp may only touch lanes [32*(warp_id%4), 32*(warp_id%4)+32).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// TMEM address = (lane << 16) | column
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) {
  return base + (lane << 16) + col;
}

// ------------------------------------------------------------------------------------------------
// misc
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16lo_to_f32(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 16-bit pair <-> float pair, format chosen at compile time (FMT_F16 / FMT_BF16)
template <int FMT>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if (FMT == FMT_F16) return pack_f16x2(lo, hi);
  return pack_bf16x2(lo, hi);
}
template <int FMT>
__device__ __forceinline__ float2 unpack2(uint32_t v) {
  if (FMT == FMT_F16) return __half22float2(*reinterpret_cast<__half2*>(&v));
  return make_float2(bf16lo_to_f32(v), bf16hi_to_f32(v));
}
// runtime-format variants (uniform branch) for the GEMM epilogue
__device__ __forceinline__ uint32_t pack2_rt(float lo, float hi, int fmt) {
  return fmt == FMT_F16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
}
__device__ __forceinline__ float2 unpack2_rt(uint32_t v, int fmt) {
  return fmt == FMT_F16 ? __half22float2(*reinterpret_cast<__half2*>(&v))
                        : make_float2(bf16lo_to_f32(v), bf16hi_to_f32(v));
}

// Packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: one issue slot for two lanes of fp32 math). The softmax loops of
// the attention kernels are issue-bound next to the MUFU pipe, so everything around the exponential is done in pairs.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t f2_add(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// Byte offset of the 16-byte chunk `chunk` (0..7) of row `row` inside a 128B-swizzled tile whose rows are
// 128 bytes (tile base 1024 B aligned). This is the layout TMA SWIZZLE_128B writes and UMMA SW128 reads.
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk) {
  return row * 128u + ((chunk ^ (row & 7u)) << 4);
}

// Stateless dropout RNG: keep-mask bit for element `idx` of tensor `salt` under `seed`.
// 32-bit mix (two rounds of multiply-xorshift); compared against a 16-bit threshold.
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
// Dropout keep-mask: one hash per PAIR of consecutive elements (thr16 = round(p * 65536)).
//   key  = dropout_key(seed, salt)                        (uniform per launch)
//   h    = dropout_bits(key, idx >> 1)                    (IMAD + xorshift)
//   keep(idx) = top 16 bits of h * C_(idx & 1) >= thr16   (one IMAD + one compare per element)
// Two odd multipliers give the two elements of a pair independent top halves; the multiplies run on the FMA pipe,
// which the fused GEMM epilogues leave idle, instead of a second xorshift + mask on the half-rate integer pipe.
// Forward and backward kernels evaluate the same function of (seed, salt, element index), so no mask is stored.
__host__ __device__ __forceinline__ uint32_t dropout_key(uint32_t seed, uint32_t salt) {
  return mix32(seed) ^ (salt * 0x85ebca6bU);
}
// seed actually used by a launch: the scalar `seed` plus an optional DEVICE word (the per-step counter that lets one
// captured CUDA graph draw a fresh mask on every replay).
__device__ __forceinline__ uint32_t effective_seed(uint32_t seed, const uint32_t* seed_dev) {
  return seed_dev ? seed + __ldg(seed_dev) : seed;
}
__host__ __device__ __forceinline__ uint32_t dropout_bits(uint32_t key, uint32_t pair_idx) {
  uint32_t h = pair_idx * 0x9E3779B1U + key;
  h ^= h >> 16;
  return h;
}
__host__ __device__ __forceinline__ bool dropout_keep_lo(uint32_t bits, uint32_t thr16) { return bits * 0x7feb352dU >= (thr16 << 16); }
__host__ __device__ __forceinline__ bool dropout_keep_hi(uint32_t bits, uint32_t thr16) { return bits * 0x846ca68bU >= (thr16 << 16); }
// per-element form (elementwise kernels)
__host__ __device__ __forceinline__ uint32_t dropout_keep(uint32_t seed, uint32_t salt, uint32_t idx, uint32_t thr16) {
  const uint32_t bits = dropout_bits(dropout_key(seed, salt), idx >> 1);
  return (idx & 1u) ? dropout_keep_hi(bits, thr16) : dropout_keep_lo(bits, thr16);
}
// `v[0..n)` holds n consecutive elements starting at even index `idx0`: zero the dropped ones, scale the kept ones.
template <int N>
__device__ __forceinline__ void dropout_apply_run(float (&v)[N], uint32_t key, uint32_t idx0, uint32_t thr16,
                                                  float scale) {
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    const uint32_t bits = dropout_bits(key, (idx0 + j) >> 1);
    v[j] = dropout_keep_lo(bits, thr16) ? v[j] * scale : 0.f;
    v[j + 1] = dropout_keep_hi(bits, thr16) ? v[j + 1] * scale : 0.f;
  }
}

// same mask, kept elements left unscaled (the caller has folded 1/(1-p) into what it feeds in)
template <int N>
__device__ __forceinline__ void dropout_zero_run(float (&v)[N], uint32_t key, uint32_t idx0, uint32_t thr16) {
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    const uint32_t bits = dropout_bits(key, (idx0 + j) >> 1);
    v[j] = dropout_keep_lo(bits, thr16) ? v[j] : 0.f;
    v[j + 1] = dropout_keep_hi(bits, thr16) ? v[j + 1] : 0.f;
  }
}

__device__ __forceinline__ void tma_store_wait_read1() {
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
// TMA reduce-add (fp32) of a smem box into global memory
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}

}  // namespace tc05
