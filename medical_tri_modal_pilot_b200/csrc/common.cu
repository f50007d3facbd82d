// common.cu -- error string, driver entry-point lookup for TMA descriptors, device properties.
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace tmp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
    set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    return nullptr;
  }
  fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  return fn;
}

static int encode_tmap_2d(CUtensorMap* map, CUtensorMapDataType dt, const void* gaddr, uint64_t inner, uint64_t outer,
                          uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer,
                          CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B);

int encode_tmap_2d_h16(CUtensorMap* map, const void* gaddr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                        uint32_t box_inner, uint32_t box_outer) {
  return encode_tmap_2d(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, gaddr, inner, outer, row_stride_bytes, box_inner, box_outer);
}
int encode_tmap_2d_h16_sw64(CUtensorMap* map, const void* gaddr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                             uint32_t box_inner, uint32_t box_outer) {
  return encode_tmap_2d(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, gaddr, inner, outer, row_stride_bytes, box_inner, box_outer,
                        CU_TENSOR_MAP_SWIZZLE_64B);
}
// IEEE fp16 element type (matters for TMA reduce-add, where the copy engine does arithmetic on the elements), 64B swizzle
int encode_tmap_2d_f16_sw64(CUtensorMap* map, const void* gaddr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                            uint32_t box_inner, uint32_t box_outer) {
  return encode_tmap_2d(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, gaddr, inner, outer, row_stride_bytes, box_inner, box_outer,
                        CU_TENSOR_MAP_SWIZZLE_64B);
}
int encode_tmap_2d_f32(CUtensorMap* map, const void* gaddr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                       uint32_t box_inner, uint32_t box_outer) {
  return encode_tmap_2d(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, gaddr, inner, outer, row_stride_bytes, box_inner, box_outer);
}

static int encode_tmap_2d(CUtensorMap* map, CUtensorMapDataType dt, const void* gaddr, uint64_t inner, uint64_t outer,
                          uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz) {
  auto fn = get_encode();
  if (!fn) return TMP_ERR_DRIVER;
  if (((uintptr_t)gaddr & 15) || (row_stride_bytes & 15)) {
    set_error("TMA operand not 16-byte aligned (addr=%p stride=%llu)", gaddr, (unsigned long long)row_stride_bytes);
    return TMP_ERR_ARG;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(gaddr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult=%d (inner=%llu outer=%llu stride=%llu box=%ux%u)", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, box_inner,
              box_outer);
    return TMP_ERR_DRIVER;
  }
  return TMP_OK;
}

// SMs that grid sizing may use = physical SMs - reserved. Every persistent / one-wave grid of this library is sized from
// this number with EQUAL work per CTA, so a kernel that finds a few SMs taken (NCCL's all-reduce CTAs during the
// data-parallel backward) runs its last CTAs as a second wave and takes twice as long. With R SMs left to the
// communication kernels (tmp_set_reserved_sms, or env TMP_B200_RESERVE_SMS) both fit side by side.
static int g_reserved_sms = -1;

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  if (g_reserved_sms < 0) {
    const char* e = getenv("TMP_B200_RESERVE_SMS");
    g_reserved_sms = e ? atoi(e) : 0;
    if (g_reserved_sms < 0) g_reserved_sms = 0;
  }
  const int avail = n - g_reserved_sms;
  return avail > 0 ? avail : 1;
}

}  // namespace tmp

extern "C" const char* tmp_last_error(void) { return tmp::g_err; }
extern "C" int tmp_set_reserved_sms(int n) {
  TMP_REQUIRE(n >= 0 && n < 64, "set_reserved_sms: n must be in [0, 64)");
  tmp::g_reserved_sms = n;
  return TMP_OK;
}
extern "C" int tmp_num_sms(void) { return tmp::num_sms(); }
extern "C" int tmp_abi_version(void) { return 5; }   // 5: tmp_set_reserved_sms / tmp_num_sms; 4: tmp_grad_nonfinite, three-word step_dev of tmp_adamw_step_dev (3: q_rows, fused attn_bwd protocol, fp32 mode)
