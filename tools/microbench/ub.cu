// Micro-benchmarks that size the attention softmax loops (B200, sm_100a): tcgen05.ld throughput per SM as a function of
// the number of reading warps, MUFU.EX2 throughput for f32 / f16x2 operands, mbarrier try_wait with a suspend-time hint.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ub ub.cu && ./ub
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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

// ---- tcgen05.ld: W warps, each reads 4 x (32 lanes x 32 columns x 4 B = 4 KB) per iteration -----------------------
__global__ void ldtm_kernel(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 3) * 128;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t a[32], b[32], c[32], d[32];
    tmem_ld32(base + 0, a);
    tmem_ld32(base + 32, b);
    tmem_ld32(base + 64, c);
    tmem_ld32(base + 96, d);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc += a[3] ^ b[7] ^ c[11] ^ d[31];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

// ---- MUFU.EX2: 16 independent chains per thread -------------------------------------------------------------------
template <int MODE>
__global__ void ex2_kernel(int iters, long long* cycles, float* sink) {
  float x[16];
  uint32_t hx[16];
  for (int i = 0; i < 16; ++i) { x[i] = -0.001f * (threadIdx.x + i); hx[i] = 0xb800b800u + i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(hx[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(hx[i]));
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += x[i] + __uint_as_float(hx[i]);
  if (s == 123.456f) sink[0] = s;
}

// ---- f16x2 ex2 accuracy over the softmax range [-16, 8] -----------------------------------------------------------
__global__ void ex2_acc_kernel(float* max_rel, float* max_rel_f32) {
  const int n = gridDim.x * blockDim.x;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float x = -16.f + 24.f * (float)i / (float)n;
  const __half2 h = __floats2half2_rn(x, x);
  uint32_t u = *reinterpret_cast<const uint32_t*>(&h);
  asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u));
  const float y = __low2float(*reinterpret_cast<__half2*>(&u));
  const double ref = exp2((double)x);
  float f;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(x));
  const float r = (float)fabs((y - ref) / ref);
  const float r32 = (float)fabs(((double)__half2float(__float2half_rn(f)) - ref) / ref);
  atomicMax((int*)max_rel, __float_as_int(r));
  atomicMax((int*)max_rel_f32, __float_as_int(r32));
}

// ---- mbarrier try_wait with suspendTimeHint: a waiter warp + an arriver after a delay -----------------------------
__global__ void trywait_kernel(int hint_ns, int delay_cycles, long long* out) {
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x < 32) {
    uint32_t ok = 0;
    long long spins = 0;
    while (!ok) {
      if (hint_ns > 0)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0), "r"(hint_ns) : "memory");
      else
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      ++spins;
    }
    if (threadIdx.x == 0) { out[0] = clock64() - t0; out[1] = spins; }
  } else if (threadIdx.x == 32) {
    while (clock64() - t0 < delay_cycles) {}
    const long long ta = clock64() - t0;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
    out[2] = ta;
  }
}

int main() {
  long long* cyc; uint32_t* sink; float* fsink;
  CK(cudaMalloc(&cyc, 1024 * 8)); CK(cudaMalloc(&sink, 64)); CK(cudaMalloc(&fsink, 64));
  long long h[1024];
  const int sms = 148;
  for (int warps : {4, 8, 16}) {
    const int iters = 2000;
    ldtm_kernel<<<sms, warps * 32>>>(iters, cyc, sink);
    CK(cudaDeviceSynchronize());
    ldtm_kernel<<<sms, warps * 32>>>(iters, cyc, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    const double bytes = (double)iters * warps * 4 * 4096;
    printf("{\"bench\": \"tcgen05.ld.32x32b.x32\", \"warps_per_sm\": %d, \"bytes_per_clk_per_sm\": %.1f, \"cycles_per_x32_per_warp\": %.1f}\n",
           warps, bytes / avg, avg / (iters * 4.0));
  }
  for (int mode = 0; mode < 3; ++mode) {
    for (int warps : {4, 8, 16}) {
      const int iters = 4000;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) ex2_kernel<0><<<sms, warps * 32>>>(iters, cyc, fsink);
        if (mode == 1) ex2_kernel<1><<<sms, warps * 32>>>(iters, cyc, fsink);
        if (mode == 2) ex2_kernel<2><<<sms, warps * 32>>>(iters, cyc, fsink);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost));
      double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
      const double insts = (double)iters * 16 * warps * 32;
      printf("{\"bench\": \"ex2.approx %s\", \"warps_per_sm\": %d, \"lane_instr_per_clk_per_sm\": %.2f, \"exp_per_clk_per_sm\": %.2f}\n",
             mode == 0 ? "f32" : mode == 1 ? "f16x2" : "bf16x2", warps, insts / avg, insts / avg * (mode == 0 ? 1 : 2));
    }
  }
  {
    float* mr; CK(cudaMalloc(&mr, 8)); CK(cudaMemset(mr, 0, 8));
    ex2_acc_kernel<<<1024, 256>>>(mr, mr + 1);
    CK(cudaDeviceSynchronize());
    float hm[2]; CK(cudaMemcpy(hm, mr, 8, cudaMemcpyDeviceToHost));
    printf("{\"bench\": \"ex2 accuracy on [-16,8]\", \"f16x2_max_rel_err\": %.3e, \"f32_then_f16_round_max_rel_err\": %.3e}\n", hm[0], hm[1]);
  }
  for (int hint : {0, 1000, 100000}) {
    long long* o; CK(cudaMalloc(&o, 32));
    trywait_kernel<<<1, 64>>>(hint, 20000, o);
    CK(cudaDeviceSynchronize());
    trywait_kernel<<<1, 64>>>(hint, 20000, o);
    CK(cudaDeviceSynchronize());
    long long ho[3]; CK(cudaMemcpy(ho, o, 24, cudaMemcpyDeviceToHost));
    printf("{\"bench\": \"mbarrier.try_wait\", \"suspend_hint_ns\": %d, \"arrive_at_cycle\": %lld, \"waiter_released_at_cycle\": %lld, \"try_wait_calls\": %lld}\n",
           hint, ho[2], ho[0], ho[1]);
  }
  return 0;
}
