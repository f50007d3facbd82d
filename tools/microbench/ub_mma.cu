// Issue-to-retire throughput of the tcgen05.mma shapes / operand layouts the attention and GEMM kernels use (B200).
// One CTA per SM, one issuing thread, N back-to-back MMAs into one accumulator, one commit; operands are whatever is
// in shared memory (throughput does not depend on the values).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../medical_tri_modal_pilot_b200/csrc -o ub_mma ub_mma.cu
#include <stdio.h>
#include "tc05.cuh"
using namespace tc05;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// kind: 0 SS A,B K-major | 1 TS (A in TMEM), B MN-major | 2 SS A K-major, B MN-major | 3 SS A,B MN-major | 4 TS B K-major
template <bool ELECT>
__global__ void __launch_bounds__(128, 1) mma_kernel(int kind, int M, int N, int n_mma, int concurrent_ld, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;  // 1.0h
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (ELECT ? (warp == 1 && elect_one()) : (threadIdx.x == 32)) {
    const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + 32768);
    const int a_mn = (kind == 3), b_mn = (kind == 1 || kind == 2 || kind == 3);
    const uint32_t idesc = make_idesc(M, N, a_mn, b_mn, FMT_F16, FMT_F16);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const int k = i & 3;
      const uint64_t adesc = a_mn ? make_sdesc_sw128(sA + k * 2048, 128 * 128, 1024) : make_sdesc_sw128(sA + k * 32, 16, 1024);
      const uint64_t bdesc = b_mn ? make_sdesc_sw128(sB + k * 2048, 128 * 128, 1024) : make_sdesc_sw128(sB + k * 32, 16, 1024);
      if (kind == 1 || kind == 4) umma_ts(tm, tm + 256 + k * 8, bdesc, idesc, 1);
      else umma_ss(tm, adesc, bdesc, idesc, 1);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  } else if (concurrent_ld && warp >= 0) {
    // other threads idle (concurrent_ld reserved)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* cyc; CK(cudaMalloc(&cyc, 1024 * 8));
  long long h[1024];
  CK(cudaFuncSetAttribute(mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaFuncSetAttribute(mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  struct Case { const char* name; int kind, M, N; } cases[] = {
      {"SS K-major x K-major  M128 N64  (bwd S^T, dP^T)", 0, 128, 64},
      {"TS A=TMEM, B MN-major M128 N64  (bwd dV, fwd PV)", 1, 128, 64},
      {"SS K-major x MN-major M128 N64  (bwd dK)", 2, 128, 64},
      {"SS MN-major x MN-major M128 N64 (bwd dQ)", 3, 128, 64},
      {"SS K-major x K-major  M128 N128 (fwd S)", 0, 128, 128},
      {"SS K-major x K-major  M128 N256 (GEMM)", 0, 128, 256},
      {"SS MN-major x MN-major M128 N256 (wgrad)", 3, 128, 256},
      {"SS MN-major x MN-major M128 N128", 3, 128, 128},
      {"TS A=TMEM, B K-major  M128 N128", 4, 128, 128},
  };
  const int n = 4096;
  for (int elect = 0; elect < 2; ++elect)
  for (auto& c : cases) {
    for (int rep = 0; rep < 2; ++rep) {
      if (elect) mma_kernel<true><<<148, 128, 100 * 1024>>>(c.kind, c.M, c.N, n, 0, cyc);
      else mma_kernel<false><<<148, 128, 100 * 1024>>>(c.kind, c.M, c.N, n, 0, cyc);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double per = avg / n;
    printf("{\"bench\": \"tcgen05.mma kind::f16 K=16\", \"issuer\": \"%s\", \"case\": \"%s\", \"cycles_per_mma\": %.1f, \"flop_per_clk_per_sm\": %.0f, \"ideal_cycles\": %.0f}\n",
           elect ? "elect.sync" : "if (lane == 0)", c.name, per, 2.0 * c.M * c.N * 16 / per, c.M * c.N / 256.0);
  }
  return 0;
}
