// rowwise.cuh -- helpers for the "one warp = one 256-wide token row" kernels (lane l owns channels 8l..8l+7).
#pragma once
#include "tc05.cuh"

namespace rw {

constexpr int D = 256;  // --transformer-dim (reference control/config.py:97); kernels are specialised for it

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void warp_sum2(float& a, float& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
}

__device__ __forceinline__ void load8_f32(const float* __restrict__ p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// 8 consecutive 16-bit elements (16 B) <-> 8 floats; FMT = tc05::FMT_F16 (activations) or FMT_BF16 (gradients)
typedef uint16_t h16;  // storage type of either 16-bit format
template <int FMT>
__device__ __forceinline__ void load8(const h16* __restrict__ p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 f = tc05::unpack2<FMT>(w[t]);
    v[2 * t] = f.x;
    v[2 * t + 1] = f.y;
  }
}
template <int FMT>
__device__ __forceinline__ void store8(h16* __restrict__ p, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(tc05::pack2<FMT>(v[0], v[1]), tc05::pack2<FMT>(v[2], v[3]),
                                            tc05::pack2<FMT>(v[4], v[5]), tc05::pack2<FMT>(v[6], v[7]));
}
// tcgen05 kind::f16 requires both MMA operands in the SAME 16-bit format (mixing fp16 x bf16 is an illegal
// instruction), and every backward GEMM multiplies a gradient by a forward tensor -> one format everywhere.
// fp16 is used (8x finer than bf16; the reference's autocast dtype); gradient tensors carry a static
// power-of-two scale applied by the host at the backward entry (see runtime.py GRAD_SCALE).
constexpr int ACT = tc05::FMT_F16;  // forward activations / weights
constexpr int GRD = tc05::FMT_F16;  // gradient tensors (scaled)
__device__ __forceinline__ void store8_f32(float* __restrict__ p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
// Storage-format-generic row access: ST = tc05::FMT_F16 (the 16-bit plan) or tc05::FMT_F32 (the fp32 "precise" mode, where
// every activation / gradient tensor is stored in fp32). `elem` is an ELEMENT offset from `base`.
template <int ST>
__device__ __forceinline__ void ld8(const void* __restrict__ base, size_t elem, float (&v)[8]) {
  if (ST == tc05::FMT_F32) {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem);
    const float4 a = p[0], b = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    load8<ST == tc05::FMT_F32 ? tc05::FMT_F16 : ST>(static_cast<const h16*>(base) + elem, v);
  }
}
template <int ST>
__device__ __forceinline__ void st8(void* __restrict__ base, size_t elem, const float (&v)[8]) {
  if (ST == tc05::FMT_F32) store8_f32(static_cast<float*>(base) + elem, v);
  else store8<ST == tc05::FMT_F32 ? tc05::FMT_F16 : ST>(static_cast<h16*>(base) + elem, v);
}

// The same 8 elements kept in their storage form (4 registers for a 16-bit tensor): what a software-pipelined loop holds for
// the NEXT row while the current one is processed.
template <int ST>
struct Raw8 {
  uint4 u;
  __device__ __forceinline__ void load(const void* __restrict__ base, size_t elem) {
    u = *reinterpret_cast<const uint4*>(static_cast<const h16*>(base) + elem);
  }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = tc05::unpack2<ST>(w[t]);
      v[2 * t] = f.x;
      v[2 * t + 1] = f.y;
    }
  }
};
template <>
struct Raw8<tc05::FMT_F32> {
  float4 a, b;
  __device__ __forceinline__ void load(const void* __restrict__ base, size_t elem) {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem);
    a = p[0];
    b = p[1];
  }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

}  // namespace rw
