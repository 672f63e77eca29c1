// Halo-tile tcgen05 kernel for 3x3 / stride 1 / pad 1 convolutions (the decoder heads: RefineNet units, composed
// linear_c o proc, conv_fuse_conv0/1 -- 77 % of the path's FLOPs).  Same math and parameter block as conv_gemm*.cuh.
//
// The generic implicit-GEMM kernel re-reads (and re-splits) every input pixel once per filter tap: 9x the L2->SM traffic
// and 9x the fp32->bf16x2 conversion work, which bounds the narrow-N layers (conv0: N=64, conv1: N=32) far below the
// tensor pipe.  Here a CTA owns a 16 x 8 output-pixel tile (128 = UMMA M) and stages the 18 x 10 input HALO of one
// 64-channel chunk ONCE in shared memory (bf16 hi + lo planes, 128 B per pixel, SWIZZLE_128B applied on absolute
// address bits).  The A operand of filter tap (ky, kx) is then just a SHIFTED VIEW of that pixel array:
//     start address = plane + (ky*10 + kx) * 128 B,   8-row groups (= 8 pixels of one image row) SBO = 10 * 128 B apart
// (descriptor semantics verified on hardware with tools/tc_probe.cu: base_offset 0, arbitrary 128 B-aligned start and
// SBO work because the swizzle is a function of the absolute shared-memory address).  Weights stream through the same
// 4-stage cp.async ring as conv_gemm_tc.cuh (SWIZZLE_64B, 32-wide K steps: two per tap and chunk).
//
//   per chunk (64 input channels): 9 taps x 4 (K=16) x 3 (lo*hi, hi*lo, hi*hi) tcgen05.mma, A halo double buffered
//   warps 0-3: producers (halo chunk c+1 is converted while chunk c is multiplied; B ring), then epilogue
//   warp 4   : MMA issue;  accumulators: BN fp32 columns of TMEM
#pragma once
#include "conv_gemm_tc.cuh"

namespace pf {

constexpr int kHtTileH = 16, kHtTileW = 8;                 // output tile (rows x cols) = 128 pixels
constexpr int kHtHaloW = kHtTileW + 2, kHtHaloH = kHtTileH + 2;
constexpr int kHtHaloPix = kHtHaloW * kHtHaloH;            // 180
constexpr int kHtPlaneBytes = 23 * 1024;                   // 180 x 128 B rounded up to a 1024 B multiple
constexpr int kHtABufBytes = 2 * kHtPlaneBytes;            // hi + lo
constexpr int kHtRounds = (kHtHaloPix + 7) / 8;            // 128 producer threads convert 8 pixels per round
constexpr int kHtBStages = 4;

template <int BN> struct HtCfg {
  static constexpr int kBBytes = BN * 64;                  // one bf16 plane of a 32-wide K step
  static constexpr int kBStageBytes = 2 * kBBytes;
  static constexpr int kSmemBytes = 2 * kHtABufBytes + kHtBStages * kBStageBytes + 256 + 1024;
  static constexpr int kTmemCols = TcCfg<BN>::kTmemCols;
  static constexpr uint32_t kIdesc = TcCfg<BN>::kIdesc;
};

// K-major SWIZZLE_128B descriptor for the halo view: 128 B rows, 8-row groups kHtHaloW * 128 B apart.
__device__ __forceinline__ uint64_t ht_a_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)((kHtHaloW * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

template <int BN>
__global__ void __launch_bounds__(kTcThreads, 1) conv3x3_tc_kernel(const ConvGemmParams p, int tiles_x, int tiles_y) {
  constexpr int kBBytes = HtCfg<BN>::kBBytes, kBStageBytes = HtCfg<BN>::kBStageBytes;
  constexpr uint32_t kIdesc = HtCfg<BN>::kIdesc;
  constexpr int kTmemCols = HtCfg<BN>::kTmemCols;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (sbase - raw);
  const uint32_t a_base = sbase;                                   // 2 x (hi plane, lo plane)
  const uint32_t b_base = sbase + 2 * kHtABufBytes;                // kHtBStages x (hi, lo)
  const uint32_t bars = b_base + kHtBStages * kBStageBytes;
  auto full_b = [&](int s) { return bars + 8u * s; };
  auto empty_b = [&](int s) { return bars + 8u * (kHtBStages + s); };
  auto full_a = [&](int i) { return bars + 8u * (2 * kHtBStages + i); };
  auto empty_a = [&](int i) { return bars + 8u * (2 * kHtBStages + 2 + i); };
  const uint32_t accum_bar = bars + 8u * (2 * kHtBStages + 4);
  const uint32_t tmem_slot = bars + 8u * (2 * kHtBStages + 5);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int t_id = blockIdx.x;
  const int tx = t_id % tiles_x; t_id /= tiles_x;
  const int ty = t_id % tiles_y;
  const int bimg = t_id / tiles_y;
  const int oy0 = ty * kHtTileH, ox0 = tx * kHtTileW;
  const int n0 = blockIdx.y * BN, g = blockIdx.z;
  const int nchunks = p.Cin / 64;
  const int nsteps = nchunks * 18;   // B steps: 9 taps x 2 halves (32 channels each) per chunk

  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < kHtBStages; ++s) { mbar_init(full_b(s), 128); mbar_init(empty_b(s), 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(full_a(i), 128); mbar_init(empty_a(i), 1); }
      mbar_init(accum_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + (tmem_slot - sbase));

  if (warp < 4) {
    // =========================================================================== producers
    const __nv_bfloat16* __restrict__ Whi = p.Whi + (long long)g * p.w_gstride;
    const __nv_bfloat16* __restrict__ Wlo = p.Wlo + (long long)g * p.w_gstride;
    const int cj = tid & 15;               // float4 index inside the 64-channel chunk
    const int psub = tid >> 4;             // pixel within a round of 8
    const long long img_pix0 = (long long)bimg * p.H * p.W;

    auto ldg_round = [&](int c, int r) -> float4 {
      const int q = r * 8 + psub;
      const int hy = q / kHtHaloW, hx = q - hy * kHtHaloW;
      const int iy = oy0 - 1 + hy, ix = ox0 - 1 + hx;
      const int ci0 = c * 64;
      const float* src = p.A;
      int ld = p.lda, coff = p.a_coff + g * p.a_gcoff + ci0;
      if (p.A2 != nullptr && ci0 >= p.c_split) { src = p.A2; ld = p.lda2; coff = p.a2_coff + ci0 - p.c_split; }
      if (q < kHtHaloPix && (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W)
        return __ldg(reinterpret_cast<const float4*>(src + (img_pix0 + (long long)iy * p.W + ix) * ld + coff + cj * 4));
      return make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto sts_round = [&](int buf, int r, float4 v) {
      const int q = r * 8 + psub;
      if (q >= kHtHaloPix) return;
      if (p.in_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      uint2 h, l;
      split_bf16x2(v.x, v.y, h.x, l.x);
      split_bf16x2(v.z, v.w, h.y, l.y);
      const uint32_t row = a_base + buf * kHtABufBytes + q * 128;           // absolute smem address of the pixel row (hi plane)
      const uint32_t off = (uint32_t)(((cj >> 1) ^ ((row >> 7) & 7)) << 4) + (uint32_t)(cj & 1) * 8u;
      unsigned char* dst = sm + (row - sbase) + off;
      *reinterpret_cast<uint2*>(dst) = h;
      *reinterpret_cast<uint2*>(dst + kHtPlaneBytes) = l;                    // plane bases are 1024 B multiples: same swizzle phase
    };
    auto load_B = [&](int k0, int s) {
      const uint32_t bhi = b_base + s * kBStageBytes;
#pragma unroll
      for (int i = 0; i < BN / 16; ++i) {
        const int q = tid + 128 * i;            // [plane][n][chunk]
        const int plane = q / (BN * 4), n = (q % (BN * 4)) >> 2, c = q & 3;
        const bool ok = n0 + n < p.N;
        const __nv_bfloat16* src = (plane ? Wlo : Whi) + (long long)(ok ? n0 + n : 0) * p.K + k0 + c * 8;
        const uint32_t dst = bhi + plane * kBBytes + n * 64 + ((c ^ ((n >> 1) & 3)) << 4);
        cp_async16(dst, src, ok);
      }
    };

    // halo of chunk 0
    for (int r = 0; r < kHtRounds; r += 4) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (r + j < kHtRounds) ? ldg_round(0, r + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) if (r + j < kHtRounds) sts_round(0, r + j, v[j]);
    }
    fence_proxy_async_smem();
    mbar_arrive(full_a(0));

    for (int t = 0; t < nsteps; ++t) {
      const int c = t / 18, u = t - c * 18;
      const int tap = u >> 1, half = u & 1;
      const int s = t % kHtBStages;
      const uint32_t ph = (t / kHtBStages) & 1;
      // halo of chunk c+1: two conversion rounds per B step, steps 1..12 of chunk c
      const bool do_a = (c + 1 < nchunks) && u >= 1 && u <= 12;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      const int r0 = (u - 1) * 2;
      if (do_a) {
        if (r0 < kHtRounds) va = ldg_round(c + 1, r0);
        if (r0 + 1 < kHtRounds) vb = ldg_round(c + 1, r0 + 1);
        if (u == 1) mbar_wait(empty_a((c + 1) & 1), ((((c + 1) >> 1) & 1) ^ 1));   // chunk c-1 has been multiplied
      }
      mbar_wait(empty_b(s), ph ^ 1);
      load_B(tap * p.Cin + c * 64 + half * 32, s);
      cp_async_commit();
      if (do_a) {
        if (r0 < kHtRounds) sts_round((c + 1) & 1, r0, va);
        if (r0 + 1 < kHtRounds) sts_round((c + 1) & 1, r0 + 1, vb);
        if (u == 12) { fence_proxy_async_smem(); mbar_arrive(full_a((c + 1) & 1)); }
      }
      if (t > 0) {
        cp_async_wait<1>();
        fence_proxy_async_smem();
        mbar_arrive(full_b((t - 1) % kHtBStages));
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    mbar_arrive(full_b((nsteps - 1) % kHtBStages));

    // =========================================================================== epilogue
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int mrow = warp * 32 + lane;                // TMEM lane == tile pixel (y * 8 + x)
    const int oy = oy0 + (mrow >> 3), ox = ox0 + (mrow & 7);
    const bool valid = oy < p.H && ox < p.W;
    const long long m = img_pix0 + (long long)oy * p.W + ox;
    const float* __restrict__ bias = p.bias ? p.bias + (long long)g * p.bias_gstride : nullptr;
    const int c_coff = p.c_coff + g * p.c_gcoff, r_coff = p.r_coff + g * p.r_gcoff, r2_coff = p.r2_coff + g * p.r2_gcoff;
    int cls_off = 0;
    if (p.bias_mode == 2) {
      const int ry = oy == 0 ? 0 : (oy == p.H - 1 ? 2 : 1);
      const int rx = ox == 0 ? 0 : (ox == p.W - 1 ? 2 : 1);
      cls_off = (ry * 3 + rx) * p.N;
    }
#pragma unroll 1
    for (int ch = 0; ch < BN / 32; ++ch) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ch * 32), v);
      const int nb = n0 + ch * 32;
      if (valid && nb < p.N) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int n = nb + q * 4;
          float4 o = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
          if (p.bias_mode) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + cls_off + n));
            o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
          }
          if (p.act == 1) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          else if (p.act == 2) { o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w); }
          if (p.gamma) {
            const float4 gv = __ldg(reinterpret_cast<const float4*>(p.gamma + n));
            o.x *= gv.x; o.y *= gv.y; o.z *= gv.z; o.w *= gv.w;
          }
          if (p.res) {
            float4 r = *reinterpret_cast<const float4*>(p.res + m * p.ldr + r_coff + n);
            if (p.res_relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
          }
          if (p.res2) {
            const float4 r = *reinterpret_cast<const float4*>(p.res2 + m * p.ldr2 + r2_coff + n);
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
          }
          *reinterpret_cast<float4*>(p.C + m * p.ldc + c_coff + n) = o;
        }
      }
    }
  } else {
    // =========================================================================== MMA issuer (warp 4)
    for (int t = 0; t < nsteps; ++t) {
      const int c = t / 18, u = t - c * 18;
      const int tap = u >> 1, half = u & 1;
      const int ky = tap / 3, kx = tap - ky * 3;
      const int s = t % kHtBStages;
      const uint32_t ph = (t / kHtBStages) & 1;
      if (u == 0) mbar_wait(full_a(c & 1), (c >> 1) & 1);
      mbar_wait(full_b(s), ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_hi = a_base + (c & 1) * kHtABufBytes + (ky * kHtHaloW + kx) * 128, a_lo = a_hi + kHtPlaneBytes;
        const uint32_t b_hi = b_base + s * kBStageBytes, b_lo = b_hi + kBBytes;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const uint32_t ka = (half * 2 + kk) * 32;   // K offset inside the 128 B pixel row
          const uint64_t dah = ht_a_desc(a_hi + ka), dal = ht_a_desc(a_lo + ka);
          const uint64_t dbh = tc_smem_desc(b_hi + kk * 32), dbl = tc_smem_desc(b_lo + kk * 32);
          umma_bf16(tmem, dal, dbh, kIdesc, (t | kk) ? 1u : 0u);
          umma_bf16(tmem, dah, dbl, kIdesc, 1u);
          umma_bf16(tmem, dah, dbh, kIdesc, 1u);
        }
        umma_commit(empty_b(s));
        if (u == 17) umma_commit(empty_a(c & 1));
        if (t == nsteps - 1) umma_commit(accum_bar);
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, kTmemCols);
}

inline bool conv3x3_tc_eligible(const ConvGemmParams& p) {
  if (conv_gemm_tc_check(p) != nullptr) return false;
  if (p.KH != 3 || p.KW != 3 || p.stride != 1 || p.pad != 1) return false;
  if (p.Cin % 64) return false;
  if (p.A2 && p.c_split % 64) return false;
  return true;
}

template <int BN>
inline cudaError_t conv3x3_tc_launch_bn(const ConvGemmParams& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, HtCfg<BN>::kSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int tiles_x = cdiv(p.W, kHtTileW), tiles_y = cdiv(p.H, kHtTileH);
  dim3 grid((unsigned)(p.B * tiles_x * tiles_y), (unsigned)cdiv(p.N, BN), (unsigned)p.groups);
  conv3x3_tc_kernel<BN><<<grid, kTcThreads, HtCfg<BN>::kSmemBytes, st>>>(p, tiles_x, tiles_y);
  return cudaGetLastError();
}

inline cudaError_t conv3x3_tc_launch(const ConvGemmParams& p, cudaStream_t st) {
  switch (conv_gemm_tc_bn(p)) {
    case 256: return conv3x3_tc_launch_bn<256>(p, st);
    case 128: return conv3x3_tc_launch_bn<128>(p, st);
    case 64: return conv3x3_tc_launch_bn<64>(p, st);
    case 32: return conv3x3_tc_launch_bn<32>(p, st);
    default: return conv_gemm_tc_launch(p, st);   // other widths: generic tcgen05 kernel
  }
}

}  // namespace pf
