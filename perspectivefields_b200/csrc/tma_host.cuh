// Host side of the TMA engine: tensor-map construction (cuTensorMapEncodeTiled through the runtime's driver entry point,
// so libcuda is not linked) and the launcher of gemm_tma_kernel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <string>

#include "gemm_tma.cuh"
#include "gemm2_tma.cuh"
#include "attention_tc.cuh"

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

// bf16 tensor [rows][ld] (row-major); box = box_rows x kb elements (kb = 32: 64 B rows, SWIZZLE_64B; kb = 64: 128 B rows,
// SWIZZLE_128B).  `cols` = logical row length.
inline const char* tma_map_2d(CUtensorMap* m, const void* base, long long cols, long long rows, long long ld, int box_rows, int kb) {
  PFN_tensorMapEncodeTiled enc = tma_encoder();
  if (!enc) return "cuTensorMapEncodeTiled entry point not available";
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kb, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  if (((uintptr_t)base & 15) || (strides[0] & 15) || box_rows < 1 || box_rows > 256) return "tma_map_2d: alignment / box";
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   kb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? nullptr : "cuTensorMapEncodeTiled (2d) failed";
}

// 32-row x 32-column epilogue tile of a row-major [rows][ld] tensor: fp32 (128 B rows, SWIZZLE_128B) or bf16 (64 B, SWIZZLE_64B)
inline const char* tma_map_tile32(CUtensorMap* m, const void* base, long long rows, long long ld, bool is_f32) {
  PFN_tensorMapEncodeTiled enc = tma_encoder();
  if (!enc) return "cuTensorMapEncodeTiled entry point not available";
  const int es_bytes = is_f32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * es_bytes};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t es[2] = {1, 1};
  if (((uintptr_t)base & 15) || (strides[0] & 15)) return "tma_map_tile32: alignment";
  CUresult r = enc(m, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, is_f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? nullptr : "cuTensorMapEncodeTiled (tile32) failed";
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

// GEMM mode with the number of rows known.  The widest tile moves the fewest operand bytes per output column (GEMM mode is bound
// by L2 -> shared-memory operand traffic: about 1.6 clk per operand row and K step of 16 on a full machine, measured -- a 128 x 256
// tile costs 614 clk per K step against 384 clk of MMA), so it wins whenever the tiles fill the machine.  A launch that leaves
// most SMs idle (stage-4 / spatially reduced layers: 25 row tiles) is faster with narrower tiles spread over more SMs: pick the
// width that minimises waves x time per tile in that model.  Results do not depend on the choice (same K order per output).
inline int tma_pick_bn_gemm(long long M, int N, int K, int sm_count) {
  const int base = tma_pick_bn(N, MODE_GEMM);
  static const int policy = getenv("PF_BN_POLICY") ? atoi(getenv("PF_BN_POLICY")) : 1;     // 0: always the widest tile (A/B runs)
  const long long mt = cdivl(M, 128);
  if (!policy || mt * cdiv(N, base) * 4 > 3LL * sm_count) return base;
  int best = base;
  double best_cost = 1e30;
  for (int bn = base; bn >= 32; bn -= 32) {
    const long long tiles = mt * cdiv(N, bn);
    const double waves = (double)cdivl(tiles, sm_count);
    const double mma = 3.0 * (bn / 2 > 32 + bn / 4 ? bn / 2 : 32 + bn / 4), load = 1.6 * (128 + bn);
    const double cost = waves * ((K / 16) * (mma > load ? mma : load) + 3000.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

// K elements per pipeline step: 64 for narrow tiles when K allows it (halo mode: BN <= 128; GEMM mode, whose stages also
// hold the A tile and whose epilogue staging takes 72 KB: BN <= 64), else 32
inline int tma_pick_kb(int bn, int K, int mode) {
  static const int wide = getenv("PF_KB64") ? atoi(getenv("PF_KB64")) : 0;     // experiment: 128-byte TMA rows for wider GEMM-mode tiles
  const int lim = mode == MODE_HALO ? 128 : (wide ? wide : 64);
  return (bn <= lim && K % 64 == 0) ? 64 : 32;
}

struct PredTail { const float* w; const float* b; float* out; int nc, mode; };   // per group, see TmaGemmParams::pred_*

template <int BN, int MODE, int KB>
inline cudaError_t gemm_tma_launch_bn(const TmaMaps& maps, const TmaGemmParams& p, int sm_count, cudaStream_t st, const PredTail* pred) {
  using Cfg = TmaCfg<BN, MODE, KB>;   // (the > 48 KB shared-memory opt-in is per device: gemm_tma_configure_device, at pf_create)
  const int tiles_x = MODE == MODE_HALO ? cdiv(p.W, kHtTileW) : 0, tiles_y = MODE == MODE_HALO ? cdiv(p.H, kHtTileH) : 0;
  const long long m_tiles = MODE == MODE_GEMM ? cdiv(p.M, 128) : (long long)p.B * tiles_x * tiles_y;
  const long long total = m_tiles * cdiv(p.N, BN) * p.groups;
  const unsigned grid = (unsigned)(total < sm_count ? total : sm_count);
  // resident-weight mode (single chunk, one N tile) assumes every tile of a CTA uses the same weights: one group per launch
  // (a fused prediction tail is per group as well: same decomposition)
  if (MODE == MODE_HALO && ((p.Cin == 64 && 9 * (64 / KB) <= Cfg::kStages && (p.groups > 1 || cdiv(p.N, BN) > 1)) || pred)) {
    cudaError_t last = cudaSuccess;
    for (int g = 0; g < p.groups; ++g)
      for (int nt = 0; nt < cdiv(p.N, BN); ++nt) {
        TmaGemmParams q = p;     // fold group g / N tile nt into the offsets of a single-group, single-tile launch
        q.groups = 1;
        q.a_c0 = p.a_c0 + g * p.a_gc;
        q.bias = p.bias ? p.bias + (long long)g * p.bias_gstride : nullptr;
        q.c_coff = p.c_coff + g * p.c_gcoff; q.s_coff = p.s_coff + g * p.s_gcoff;
        q.r_coff = p.r_coff + g * p.r_gcoff; q.r2_coff = p.r2_coff + g * p.r2_gcoff;
        q.b_row0 = g * p.N;
        if (pred) { q.pred_w = pred[g].w; q.pred_b = pred[g].b; q.pred_out = pred[g].out; q.pred_nc = pred[g].nc; q.pred_mode = pred[g].mode; }
        if (cdiv(p.N, BN) > 1) return cudaErrorInvalidValue;   // (not needed by the network: conv_fuse_conv1 has one N tile)
        const unsigned gr = (unsigned)(m_tiles < sm_count ? m_tiles : sm_count);
        last = launch_pdl(gemm_tma_kernel<BN, MODE, KB>, dim3(gr), dim3(kTmaThreads), Cfg::kSmemBytes, st, maps, q, tiles_x, tiles_y);
        if (last != cudaSuccess) return last;
      }
    return last;
  }
  return launch_pdl(gemm_tma_kernel<BN, MODE, KB>, dim3(grid), dim3(kTmaThreads), Cfg::kSmemBytes, st, maps, p, tiles_x, tiles_y);
}

// every instantiation the dispatcher below can reach: X(BN, MODE, KB)
#define PF_TMA_VARIANTS(X)                                                                                                  \
  X(256, MODE_GEMM, 32) X(224, MODE_GEMM, 32) X(192, MODE_GEMM, 32) X(160, MODE_GEMM, 32) X(128, MODE_GEMM, 32) X(96, MODE_GEMM, 32) \
  X(64, MODE_GEMM, 32) X(32, MODE_GEMM, 32) X(64, MODE_GEMM, 64) X(32, MODE_GEMM, 64)                                        \
  X(96, MODE_GEMM, 64) X(128, MODE_GEMM, 64) X(160, MODE_GEMM, 64)                                                           \
  X(256, MODE_HALO, 32) X(128, MODE_HALO, 64) X(64, MODE_HALO, 64) X(32, MODE_HALO, 64)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: called once for every device an engine is
// created on (pf_create) -- not behind a process-wide flag.
inline cudaError_t gemm_tma_configure_device() {
  cudaError_t e = cudaSuccess;
#define PF_TMA_CFG(BN_, MODE_, KB_)                                                                                          \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tma_kernel<BN_, MODE_, KB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, TmaCfg<BN_, MODE_, KB_>::kSmemBytes);
  PF_TMA_VARIANTS(PF_TMA_CFG)
#undef PF_TMA_CFG
  return e;
}

// ---- attention core on tcgen05 (attention_tc.cuh)
inline cudaError_t attention_tc_configure_device() {
  return cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtcSmemBytes);
}
inline cudaError_t attention_tc_launch(const AtcMaps& maps, __nv_bfloat16* ohi, __nv_bfloat16* olo, int B, int N, int C, int heads, int sm_count,
                                       cudaStream_t st) {
  const int total = B * heads * cdiv(N, 128);
  int grid = total < sm_count ? total : sm_count;
  const int ipc = cdiv(total, grid);          // contiguous items per CTA: K / V of an (image, head) are loaded once per CTA that touches it
  grid = cdiv(total, ipc);
  return launch_pdl(attention_tc_kernel, dim3(grid), dim3(kAtcThreads), kAtcSmemBytes, st, maps, ohi, olo, B, N, C, heads, total, ipc);
}

// ---- CTA-pair GEMM (gemm2_tma.cuh): cluster 2x1x1, one pair per TPC
#define PF_TMA2_VARIANTS(X) X(256) X(224) X(192) X(160) X(128) X(96) X(64)

struct Gemm2Info { int max_clusters[9]; };   // index BN / 32: concurrently resident pairs on this device (0 = not available)
inline Gemm2Info& gemm2_info(int device) {
  static Gemm2Info info[64];
  return info[device & 63];
}
inline cudaError_t gemm2_configure_device(int device, int sm_count) {
  cudaError_t e = cudaSuccess;
  Gemm2Info& gi = gemm2_info(device);
#define PF_TMA2_CFG(BN_)                                                                                                      \
  if (e == cudaSuccess) {                                                                                                     \
    e = cudaFuncSetAttribute(gemm2_tma_kernel<BN_>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tma2Cfg<BN_>::kSmemBytes);   \
    if (e == cudaSuccess) {                                                                                                   \
      cudaLaunchConfig_t cfg{};                                                                                               \
      cfg.gridDim = dim3(sm_count & ~1); cfg.blockDim = dim3(kTmaThreads); cfg.dynamicSmemBytes = Tma2Cfg<BN_>::kSmemBytes;   \
      cudaLaunchAttribute at[1];                                                                                              \
      at[0].id = cudaLaunchAttributeClusterDimension;                                                                         \
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;                                     \
      cfg.attrs = at; cfg.numAttrs = 1;                                                                                       \
      int n = 0;                                                                                                              \
      if (cudaOccupancyMaxActiveClusters(&n, gemm2_tma_kernel<BN_>, &cfg) != cudaSuccess) { n = 0; (void)cudaGetLastError(); } \
      gi.max_clusters[BN_ / 32] = n;                                                                                          \
    }                                                                                                                         \
  }
  PF_TMA2_VARIANTS(PF_TMA2_CFG)
#undef PF_TMA2_CFG
  return e;
}

template <int BN>
inline cudaError_t gemm2_launch_bn(const TmaMaps& maps, const TmaGemmParams& p, int nclusters, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * nclusters); cfg.blockDim = dim3(kTmaThreads); cfg.dynamicSmemBytes = Tma2Cfg<BN>::kSmemBytes; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, gemm2_tma_kernel<BN>, maps, p);
}
// pair tiles of this problem and the number of clusters to launch (0: use the single-CTA kernel)
inline int gemm2_plan(int device, int M, int N, int bn) {
  const int avail = gemm2_info(device).max_clusters[bn / 32];
  if (avail < 1 || bn < 64) return 0;
  const long long tiles = (long long)cdiv(M, 256) * cdiv(N, bn);
  if (tiles < avail) return 0;              // too few pair tiles to cover the machine once: 128-row tiles spread better
  return avail;
}
inline cudaError_t gemm2_launch(const TmaMaps& maps, const TmaGemmParams& p, int bn, int nclusters, cudaStream_t st) {
#define PF_TMA2_CASE(BN_) if (bn == BN_) return gemm2_launch_bn<BN_>(maps, p, nclusters, st);
  PF_TMA2_VARIANTS(PF_TMA2_CASE)
#undef PF_TMA2_CASE
  return cudaErrorInvalidValue;
}

inline cudaError_t gemm_tma_launch(int mode, const TmaMaps& maps, const TmaGemmParams& p, int bn, int kb, int sm_count, cudaStream_t st,
                                   const PredTail* pred = nullptr) {
#define PF_TMA_CASE(BN_, MODE_, KB_) if (mode == MODE_ && bn == BN_ && kb == KB_) return gemm_tma_launch_bn<BN_, MODE_, KB_>(maps, p, sm_count, st, pred);
  PF_TMA_VARIANTS(PF_TMA_CASE)
#undef PF_TMA_CASE
  return cudaErrorInvalidValue;
}

}  // namespace pf
