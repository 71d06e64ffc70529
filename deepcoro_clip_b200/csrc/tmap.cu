// Host helpers: TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency)
// and cached device properties.
#include "common.cuh"

#include <mutex>

namespace b2host {
using namespace b2;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                      uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return B2_ENOSYS;
  if (!base || (reinterpret_cast<uintptr_t>(base) & 15) || (pitch_elems * 2) % 16 || box_rows == 0 ||
      box_rows > 256)
    return B2_EINVAL;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? B2_OK : B2_EINVAL;
}

// Dense [rows, 64] bf16 matrix (row pitch 128 bytes), no swizzle: boxes of `box_rows` rows are copied byte for byte. Used
// for buffers that already hold SWIZZLE_128B operand images (the stored gradient tiles of logits_bwd3.cu / gt_gemm.cu).
int make_tmap_bf16_rows64(CUtensorMap* out, const void* base, uint64_t rows, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return B2_ENOSYS;
  if (!base || (reinterpret_cast<uintptr_t>(base) & 127) || box_rows == 0 || box_rows > 256 || rows == 0) return B2_EINVAL;
  cuuint64_t gdim[2] = {64, rows};
  cuuint64_t gstride[1] = {128};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? B2_OK : B2_EINVAL;
}

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

int sm_count() {
  static int n = 0;     // one device model per process (every GPU of a B200 box has the same SM count)
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, current_device());
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace b2host
