// tcgen05 / TMEM engine for the implicit-GEMM convolution (same parameter block and math as conv_gemm.cuh).
//
//   tile 128 (pixels) x 256 (output channels), K step 32, 4-stage shared-memory ring, fp32 accumulators in TMEM
//   warps 0-3 : producers, then epilogue.  A: global fp32 (im2col gather, optional ReLU) -> bf16 hi/lo split in
//               registers -> st.shared in the canonical K-major SWIZZLE_64B layout;  B: pre-split bf16 weights,
//               cp.async 16 B chunks into the same layout;  fence.proxy.async + mbarrier arrive per stage.
//   warp 4    : TMEM allocation and MMA issue: per stage 2 (K=16) x 3 (lo*hi, hi*lo, hi*hi) tcgen05.mma.kind::f16
//               (bf16 inputs, fp32 accumulate), tcgen05.commit to the stage's "empty" barrier.
//   epilogue  : tcgen05.ld 32x32b.x32 (thread = one output pixel row of the tile), bias / activation / layer-scale /
//               residuals in registers, float4 stores.
//
// Every mbarrier wait is bounded (clock64 watchdog -> __trap) so that a protocol bug aborts the launch instead of
// hanging the device.
#pragma once
#include "conv_gemm.cuh"

namespace pf {

constexpr int kTcBM = 128, kTcBK = 32;
constexpr int kTcABytes = kTcBM * 64;                       // one bf16 plane of the A tile (64 B rows)
constexpr int kTcThreads = 160;
// per-BN configuration: N tile, smem ring depth (BN=256: 4 x 48 KB, one CTA/SM; narrower tiles leave room for 2 CTAs/SM)
template <int BN> struct TcCfg {
  static constexpr int kStages = BN > 128 ? 4 : (BN > 64 ? 3 : 4);
  static constexpr int kBBytes = BN * 64;
  static constexpr int kStageBytes = 2 * kTcABytes + 2 * kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 256 + 1024;
  static constexpr int kTmemCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));  // power of two >= BN
  // Instruction descriptor: fp32 accumulate (bits 4-5 = 1), A/B = bf16 (bits 7-9, 10-12 = 1), both K-major, N >> 3 at 17, M >> 4 at 24.
  static constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
};

// ------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s: protocol bug, abort instead of hanging the GPU
  }
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_slot), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                 "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                 "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_64B: rows of 64 B (32 bf16), 8-row groups 512 B apart.
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major, 1) |
//   [32,46) stride byte offset >> 4 (512 B between 8-row groups) | [46,48) version = 1 (Blackwell) | [61,64) layout = 4 (SW64)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
template <int BN>
__global__ void __launch_bounds__(kTcThreads, 1) conv_gemm_tc_kernel(const ConvGemmParams p) {
  constexpr int kTcBN = BN, kTcStages = TcCfg<BN>::kStages, kTcBBytes = TcCfg<BN>::kBBytes, kTcStageBytes = TcCfg<BN>::kStageBytes;
  constexpr uint32_t kTcIdesc = TcCfg<BN>::kIdesc;
  constexpr int kTmemCols = TcCfg<BN>::kTmemCols;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t sbase = (raw + 1023u) & ~1023u;          // 1024 B aligned tile buffers
  unsigned char* sm = smem_dyn + (sbase - raw);
  const uint32_t bars = sbase + kTcStages * kTcStageBytes;  // full[4], empty[4], accum, tmem slot
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kTcStages + s); };
  const uint32_t accum_bar = bars + 8u * (2 * kTcStages);
  const uint32_t tmem_slot = bars + 8u * (2 * kTcStages + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * kTcBM, n0 = blockIdx.y * kTcBN, g = blockIdx.z;
  const int OHW = p.OH * p.OW;
  const int M = p.B * OHW;
  const int nk = p.K / kTcBK;

  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < kTcStages; ++s) { mbar_init(full_bar(s), 128); mbar_init(empty_bar(s), 1); }
      mbar_init(accum_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + kTcStages * kTcStageBytes + 8 * (2 * kTcStages + 1));

  if (warp < 4) {
    // =========================================================================== producers
    const __nv_bfloat16* __restrict__ Whi = p.Whi + (long long)g * p.w_gstride;
    const __nv_bfloat16* __restrict__ Wlo = p.Wlo + (long long)g * p.w_gstride;
    const int a_cg = tid & 7;
    int a_pix0[8], a_yx0[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int m = m0 + (tid >> 3) + 16 * j;
      if (m < M) {
        const int b = m / OHW, rem = m - b * OHW;
        const int oy = rem / p.OW, ox = rem - oy * p.OW;
        a_pix0[j] = b * p.H * p.W;
        a_yx0[j] = ((oy * p.stride - p.pad) << 16) | ((ox * p.stride - p.pad) & 0xffff);
      } else {
        a_pix0[j] = 0;
        a_yx0[j] = (int)0xC0000000;
      }
    }
    float4 areg[2][8];
    auto load_A = [&](int kc, float4 (&dst)[8]) {
      const int k0 = kc * kTcBK;
      const int tap = k0 / p.Cin, ci0 = k0 - tap * p.Cin;
      const int ky = tap / p.KW, kx = tap - ky * p.KW;
      const float* src = p.A;
      int ld = p.lda, coff = p.a_coff + g * p.a_gcoff + ci0;
      if (p.A2 != nullptr && ci0 >= p.c_split) { src = p.A2; ld = p.lda2; coff = p.a2_coff + ci0 - p.c_split; }
      coff += a_cg * 4;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int iy = (a_yx0[j] >> 16) + ky;
        const int ix = (int)(short)(a_yx0[j] & 0xffff) + kx;
        if ((unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W)
          dst[j] = __ldg(reinterpret_cast<const float4*>(src + (long long)(a_pix0[j] + iy * p.W + ix) * ld + coff));
        else
          dst[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto store_A = [&](int s, const float4 (&srcv)[8]) {
      unsigned char* hi = sm + s * kTcStageBytes;
      unsigned char* lo = hi + kTcABytes;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 v = srcv[j];
        if (p.in_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        uint2 h, l;
        split_bf16x2(v.x, v.y, h.x, l.x);
        split_bf16x2(v.z, v.w, h.y, l.y);
        const int r = (tid >> 3) + 16 * j;
        const int chunk = (a_cg >> 1) ^ ((r >> 1) & 3);           // SWIZZLE_64B: 16 B chunk index ^= address bits [7,9)
        const int off = r * 64 + chunk * 16 + (a_cg & 1) * 8;
        *reinterpret_cast<uint2*>(hi + off) = h;
        *reinterpret_cast<uint2*>(lo + off) = l;
      }
    };
    auto load_B = [&](int kc, int s) {
      const int k0 = kc * kTcBK;
      const uint32_t bhi = sbase + s * kTcStageBytes + 2 * kTcABytes;
#pragma unroll
      for (int i = 0; i < kTcBN / 16; ++i) {
        const int q = tid + 128 * i;            // [plane][n][chunk]
        const int plane = q / (kTcBN * 4), n = (q % (kTcBN * 4)) >> 2, c = q & 3;
        const bool ok = n0 + n < p.N;
        const __nv_bfloat16* src = (plane ? Wlo : Whi) + (long long)(ok ? n0 + n : 0) * p.K + k0 + c * 8;
        const uint32_t dst = bhi + plane * kTcBBytes + n * 64 + ((c ^ ((n >> 1) & 3)) << 4);
        cp_async16(dst, src, ok);
      }
    };

    load_A(0, areg[0]);
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % kTcStages;
      const uint32_t ph = (kc / kTcStages) & 1;
      if (kc + 1 < nk) {
        if (kc & 1) load_A(kc + 1, areg[0]); else load_A(kc + 1, areg[1]);
      }
      mbar_wait(empty_bar(s), ph ^ 1);
      load_B(kc, s);
      cp_async_commit();
      if (kc & 1) store_A(s, areg[1]); else store_A(s, areg[0]);
      if (kc > 0) {
        cp_async_wait<1>();          // B of step kc-1 has landed (its A was stored one iteration ago)
        fence_proxy_async_smem();
        mbar_arrive(full_bar((kc - 1) % kTcStages));
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    mbar_arrive(full_bar((nk - 1) % kTcStages));

    // =========================================================================== epilogue
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int m = m0 + warp * 32 + lane;     // TMEM lane == tile row
    const float* __restrict__ bias = p.bias ? p.bias + (long long)g * p.bias_gstride : nullptr;
    const int c_coff = p.c_coff + g * p.c_gcoff, r_coff = p.r_coff + g * p.r_gcoff, r2_coff = p.r2_coff + g * p.r2_gcoff;
    int cls_off = 0;
    if (p.bias_mode == 2 && m < M) {
      const int rem = m % OHW;
      const int oy = rem / p.OW, ox = rem - oy * p.OW;
      const int ry = oy == 0 ? 0 : (oy == p.OH - 1 ? 2 : 1);
      const int rx = ox == 0 ? 0 : (ox == p.OW - 1 ? 2 : 1);
      cls_off = (ry * 3 + rx) * p.N;
    }
#pragma unroll 1
    for (int ch = 0; ch < kTcBN / 32; ++ch) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ch * 32), v);   // warp-collective: no divergence before this
      const int nb = n0 + ch * 32;
      if (m < M && nb < p.N) {
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
            float4 r = *reinterpret_cast<const float4*>(p.res + (long long)m * p.ldr + r_coff + n);
            if (p.res_relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
          }
          if (p.res2) {
            const float4 r = *reinterpret_cast<const float4*>(p.res2 + (long long)m * p.ldr2 + r2_coff + n);
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
          }
          *reinterpret_cast<float4*>(p.C + (long long)m * p.ldc + c_coff + n) = o;
        }
      }
    }
  } else {
    // =========================================================================== MMA issuer (warp 4)
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % kTcStages;
      const uint32_t ph = (kc / kTcStages) & 1;
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_hi = sbase + s * kTcStageBytes, a_lo = a_hi + kTcABytes;
        const uint32_t b_hi = a_hi + 2 * kTcABytes, b_lo = b_hi + kTcBBytes;
#pragma unroll
        for (int kk = 0; kk < kTcBK / 16; ++kk) {
          const uint64_t dah = tc_smem_desc(a_hi + kk * 32), dal = tc_smem_desc(a_lo + kk * 32);
          const uint64_t dbh = tc_smem_desc(b_hi + kk * 32), dbl = tc_smem_desc(b_lo + kk * 32);
          umma_bf16(tmem, dal, dbh, kTcIdesc, (kc | kk) ? 1u : 0u);
          umma_bf16(tmem, dah, dbl, kTcIdesc, 1u);
          umma_bf16(tmem, dah, dbh, kTcIdesc, 1u);
        }
        umma_commit(empty_bar(s));                 // slot reusable once these MMAs have read it
        if (kc == nk - 1) umma_commit(accum_bar);  // accumulator complete
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, kTmemCols);
}

inline const char* conv_gemm_tc_check(const ConvGemmParams& p) {
  const char* m = conv_gemm_check(p);
  if (m) return m;
  if (p.N % 4) return "conv_gemm_tc: N must be a multiple of 4";
  if (p.ldc % 4 || p.c_coff % 4 || p.c_gcoff % 4) return "conv_gemm_tc: C alignment (float4)";
  if (p.res && (p.ldr % 4 || p.r_coff % 4 || p.r_gcoff % 4)) return "conv_gemm_tc: res alignment (float4)";
  if (p.res2 && (p.ldr2 % 4 || p.r2_coff % 4 || p.r2_gcoff % 4)) return "conv_gemm_tc: res2 alignment (float4)";
  if (p.N % 32) return "conv_gemm_tc: N must be a multiple of 32";
  return nullptr;
}

// N tile for a problem (0 = not eligible for this engine): the fewest tiles of at most 256 columns, then the narrowest
// multiple of 32 that covers N with that many tiles (N = 320 -> 2 x 160, 640 -> 3 x 224, 384 -> 2 x 192, 96 -> 96).
inline int conv_gemm_tc_bn(const ConvGemmParams& p) {
  if (conv_gemm_tc_check(p) != nullptr) return 0;
  const int tiles = cdiv(p.N, 256);
  return cdiv(cdiv(p.N, tiles), 32) * 32;
}
inline bool conv_gemm_tc_eligible(const ConvGemmParams& p) { return conv_gemm_tc_bn(p) != 0; }

template <int BN>
inline cudaError_t conv_gemm_tc_launch_bn(const ConvGemmParams& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::kSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const long long M = (long long)p.B * p.OH * p.OW;
  dim3 grid((unsigned)cdivl(M, kTcBM), (unsigned)cdiv(p.N, BN), (unsigned)p.groups);
  conv_gemm_tc_kernel<BN><<<grid, kTcThreads, TcCfg<BN>::kSmemBytes, st>>>(p);
  return cudaGetLastError();
}

inline cudaError_t conv_gemm_tc_launch(const ConvGemmParams& p, cudaStream_t st) {
  switch (conv_gemm_tc_bn(p)) {
    case 256: return conv_gemm_tc_launch_bn<256>(p, st);
    case 224: return conv_gemm_tc_launch_bn<224>(p, st);
    case 192: return conv_gemm_tc_launch_bn<192>(p, st);
    case 160: return conv_gemm_tc_launch_bn<160>(p, st);
    case 128: return conv_gemm_tc_launch_bn<128>(p, st);
    case 96: return conv_gemm_tc_launch_bn<96>(p, st);
    case 64: return conv_gemm_tc_launch_bn<64>(p, st);
    case 32: return conv_gemm_tc_launch_bn<32>(p, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace pf
