// Host side of the TMA engine: tensor-map construction (cuTensorMapEncodeTiled through the runtime's driver entry point,
// so libcuda is not linked) and the launcher of gemm_tma_kernel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <string>

#include "gemm_tma.cuh"

namespace pf {

typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tensorMapEncodeTiled tma_encoder() {
  static PFN_tensorMapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_tensorMapEncodeTiled)p;
  }
  return fn;
}

// bf16 tensor [rows][ld] (row-major); box = box_rows x 32 elements (64 B), SWIZZLE_64B.  `cols` = logical row length.
inline const char* tma_map_2d(CUtensorMap* m, const void* base, long long cols, long long rows, long long ld, int box_rows) {
  PFN_tensorMapEncodeTiled enc = tma_encoder();
  if (!enc) return "cuTensorMapEncodeTiled entry point not available";
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  if (((uintptr_t)base & 15) || (strides[0] & 15) || box_rows < 1 || box_rows > 256) return "tma_map_2d: alignment / box";
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? nullptr : "cuTensorMapEncodeTiled (2d) failed";
}

// bf16 NHWC tensor [B][H][W][ld]; box = 1 x 18 x 10 x 64 channels (128 B), SWIZZLE_128B: one halo chunk.
inline const char* tma_map_halo(CUtensorMap* m, const void* base, int B, int H, int W, int ld) {
  PFN_tensorMapEncodeTiled enc = tma_encoder();
  if (!enc) return "cuTensorMapEncodeTiled entry point not available";
  cuuint64_t dims[4] = {(cuuint64_t)ld, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)kHtHaloW, (cuuint32_t)kHtHaloH, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  if (((uintptr_t)base & 15) || (strides[0] & 15)) return "tma_map_halo: alignment";
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? nullptr : "cuTensorMapEncodeTiled (halo) failed";
}

// N tile: fewest tiles of <= 256 columns, then the narrowest multiple of 32 covering N (halo mode: powers of two only).
inline int tma_pick_bn(int N, int mode) {
  const int tiles = cdiv(N, 256);
  int bn = cdiv(cdiv(N, tiles), 32) * 32;
  if (mode == MODE_HALO) bn = bn <= 32 ? 32 : (bn <= 64 ? 64 : (bn <= 128 ? 128 : 256));
  return bn;
}

template <int BN, int MODE>
inline cudaError_t gemm_tma_launch_bn(const TmaMaps& maps, const TmaGemmParams& p, int sm_count, cudaStream_t st) {
  using Cfg = TmaCfg<BN, MODE>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tma_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int tiles_x = MODE == MODE_HALO ? cdiv(p.W, kHtTileW) : 0, tiles_y = MODE == MODE_HALO ? cdiv(p.H, kHtTileH) : 0;
  const long long m_tiles = MODE == MODE_GEMM ? cdiv(p.M, 128) : (long long)p.B * tiles_x * tiles_y;
  const long long total = m_tiles * cdiv(p.N, BN) * p.groups;
  const unsigned grid = (unsigned)(total < sm_count ? total : sm_count);
  gemm_tma_kernel<BN, MODE><<<grid, kTmaThreads, Cfg::kSmemBytes, st>>>(maps, p, tiles_x, tiles_y);
  return cudaGetLastError();
}

inline cudaError_t gemm_tma_launch(int mode, const TmaMaps& maps, const TmaGemmParams& p, int bn, int sm_count, cudaStream_t st) {
  if (mode == MODE_GEMM) {
    switch (bn) {
      case 256: return gemm_tma_launch_bn<256, MODE_GEMM>(maps, p, sm_count, st);
      case 224: return gemm_tma_launch_bn<224, MODE_GEMM>(maps, p, sm_count, st);
      case 192: return gemm_tma_launch_bn<192, MODE_GEMM>(maps, p, sm_count, st);
      case 160: return gemm_tma_launch_bn<160, MODE_GEMM>(maps, p, sm_count, st);
      case 128: return gemm_tma_launch_bn<128, MODE_GEMM>(maps, p, sm_count, st);
      case 96: return gemm_tma_launch_bn<96, MODE_GEMM>(maps, p, sm_count, st);
      case 64: return gemm_tma_launch_bn<64, MODE_GEMM>(maps, p, sm_count, st);
      case 32: return gemm_tma_launch_bn<32, MODE_GEMM>(maps, p, sm_count, st);
    }
  } else {
    switch (bn) {
      case 256: return gemm_tma_launch_bn<256, MODE_HALO>(maps, p, sm_count, st);
      case 128: return gemm_tma_launch_bn<128, MODE_HALO>(maps, p, sm_count, st);
      case 64: return gemm_tma_launch_bn<64, MODE_HALO>(maps, p, sm_count, st);
      case 32: return gemm_tma_launch_bn<32, MODE_HALO>(maps, p, sm_count, st);
    }
  }
  return cudaErrorInvalidValue;
}

}  // namespace pf
