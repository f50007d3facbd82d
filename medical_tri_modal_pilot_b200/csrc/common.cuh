// common.cuh -- host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define TMP_OK 0
#define TMP_ERR_ARG (-1)
#define TMP_ERR_DRIVER (-2)

namespace tmp {

void set_error(const char* fmt, ...);

// cuTensorMapEncodeTiled resolved at run time through the runtime API (no link-time libcuda dependency:
// the library must load on the CPU-only build box).
int encode_tmap_2d_h16(CUtensorMap* map, const void* gaddr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                        uint32_t box_inner, uint32_t box_outer);

// 16-bit elements, 64B swizzle (box_inner = 32 elements = 64 B rows, dense in shared memory: with the 128B swizzle a
// 64 B box row is padded to a 128 B pitch, measured): chunk c (16 B) of row r sits at r*64 + ((c ^ ((r >> 1) & 3)) << 4).
int encode_tmap_2d_h16_sw64(CUtensorMap* map, const void* gaddr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                             uint32_t box_inner, uint32_t box_outer);

// IEEE fp16 elements (element type matters for TMA reduce-add), 64B swizzle, box_inner = 32
int encode_tmap_2d_f16_sw64(CUtensorMap* map, const void* gaddr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                            uint32_t box_inner, uint32_t box_outer);

// fp32 elements, 128B swizzle (box_inner <= 32 floats): used for TMA reduce-add of fp32 accumulators.
int encode_tmap_2d_f32(CUtensorMap* map, const void* gaddr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                       uint32_t box_inner, uint32_t box_outer);

int num_sms();

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return TMP_OK;
}

}  // namespace tmp

#define TMP_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      tmp::set_error(__VA_ARGS__);  \
      return TMP_ERR_ARG;           \
    }                               \
  } while (0)
