// Implicit-GEMM convolution / linear layer on NHWC fp32 activations, bf16x3 split-precision tensor-core math.
//
//   C[m, n] = epilogue( sum_k A_im2col[m, k] * W[n, k] ),   m = (b, oy, ox),  k = (ky, kx, ci)  (ci fastest)
//
// Replaces every dense contraction of the reference graph: nn.Conv2d (groups=1) and nn.Linear calls in
// mix_transformers.py:49-56,108-141,243-249, decode_head.py:51-54,244-256, gravity_head.py:139-176,
// convnext.py:46-59,140-152 (see DESIGN.md for the layer -> launch table).
//
// Precision: operands are split on the fly into bf16 hi + bf16 lo (16 significant bits) and each product is
// evaluated as lo*hi + hi*lo + hi*hi with fp32 accumulation -- 3 bf16 MMAs, measured 7e-5 max relative
// output error end to end against the fp32 reference (1xTF32 gives 4.5e-3 and fails the 1e-3 bar).
// Weights are pre-split at load time ([N][K] bf16 hi / lo planes); activations stay fp32 in HBM and are split
// while being staged into shared memory, which is also where the ReLU prologue of the RefineNet units is applied.
//
// This is the warp-level (mma.sync / HMMA) engine: 128 x BN x 32 tiles, 8 warps, register-staged A (global fp32
// -> split -> st.shared), cp.async B, double buffered.  The tcgen05/TMEM engine for the 3x3 head convolutions
// lives in conv_gemm_tc.cuh and shares this parameter block.
#pragma once
#include "common.cuh"

namespace pf {

struct ConvGemmParams {
  // ---- A: activations, NHWC fp32.  Channel c of pixel q of source s is at s.ptr[q * s.ld + s.coff + c].
  const float* A;   int lda;  int a_coff;
  const float* A2;  int lda2; int a2_coff; int c_split;  // channels >= c_split are read from A2 (virtual concat)
  int B, H, W, Cin;                                       // input geometry
  int OH, OW, KH, KW, stride, pad;                        // output geometry / filter
  int in_relu;                                            // apply ReLU to A while staging
  // ---- W: pre-split weights, [N][K] bf16, K = KH*KW*Cin ordered (ky, kx, ci)
  const __nv_bfloat16* Whi; const __nv_bfloat16* Wlo;
  int N, K;
  // ---- epilogue:  v = acc + bias;  v = act(v);  v *= gamma;  v += relu?(res);  v += res2
  const float* bias; int bias_mode;   // 0 none, 1: bias[n], 2: bias[cls*N + n], cls = 3x3 border class of the pixel
  int act;                            // 0 none, 1 ReLU, 2 GELU(erf)
  const float* gamma;                 // per-channel layer scale (ConvNeXt) or nullptr
  const float* res;  int ldr,  r_coff;  int res_relu;
  const float* res2; int ldr2, r2_coff;
  float* C; int ldc, c_coff;
  // ---- groups (blockIdx.z): independent GEMMs sharing the geometry (the two decoder heads)
  int groups;
  int a_gcoff;            // A channel offset step per group
  long long w_gstride;    // elements between the groups' weight planes
  int bias_gstride;       // floats between the groups' bias tables
  int c_gcoff, r_gcoff, r2_gcoff;  // channel offset steps for C / res / res2
};

constexpr int kGemmBK = 32;
constexpr int kGemmPitch = 40;  // bf16 elements per smem row (80 B): conflict-free ldmatrix, 16 B aligned rows
constexpr int kGemmThreads = 256;

template <int BM, int BN>
constexpr int conv_gemm_smem_bytes() { return 2 /*stages*/ * 2 /*hi,lo*/ * (BM + BN) * kGemmPitch * 2; }

template <int BM, int BN, int WARPS_M, int WARPS_N>
__global__ void __launch_bounds__(kGemmThreads, 2) conv_gemm_kernel(const ConvGemmParams p) {
  constexpr int BK = kGemmBK, PITCH = kGemmPitch;
  static_assert(WARPS_M * WARPS_N == 8, "8 warps");
  constexpr int WTM = BM / WARPS_M, WTN = BN / WARPS_N;
  constexpr int MT = WTM / 16, NT = WTN / 8;
  static_assert(MT >= 1 && NT >= 2 && NT % 2 == 0, "tile shape");
  constexpr int A_ROWS = BM / 32;                      // rows per thread of the A tile (8 threads x float4 per row)
  constexpr int B_CHUNKS = (BN * 4 * 2) / kGemmThreads;  // 16 B cp.async chunks per thread per k-step (hi + lo)
  static_assert((BN * 4 * 2) % kGemmThreads == 0, "B tile");
  constexpr int STAGE_ELEMS = 2 * (BM + BN) * PITCH;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* smem = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  auto sAhi = [&](int s) { return smem + s * STAGE_ELEMS; };
  auto sAlo = [&](int s) { return smem + s * STAGE_ELEMS + BM * PITCH; };
  auto sBhi = [&](int s) { return smem + s * STAGE_ELEMS + 2 * BM * PITCH; };
  auto sBlo = [&](int s) { return smem + s * STAGE_ELEMS + 2 * BM * PITCH + BN * PITCH; };

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN, g = blockIdx.z;
  const int OHW = p.OH * p.OW;
  const int M = p.B * OHW;
  const int nk = p.K / BK;

  const __nv_bfloat16* __restrict__ Whi = p.Whi + (long long)g * p.w_gstride;
  const __nv_bfloat16* __restrict__ Wlo = p.Wlo + (long long)g * p.w_gstride;

  // ---- per-thread A rows: pixel base and top-left input coordinate of the receptive field
  const int a_cg = tid & 7;
  int a_pix0[A_ROWS], a_yx0[A_ROWS];
#pragma unroll
  for (int i = 0; i < A_ROWS; ++i) {
    const int m = m0 + (tid >> 3) + 32 * i;
    if (m < M) {
      const int b = m / OHW, rem = m - b * OHW;
      const int oy = rem / p.OW, ox = rem - oy * p.OW;
      a_pix0[i] = b * p.H * p.W;
      a_yx0[i] = ((oy * p.stride - p.pad) << 16) | ((ox * p.stride - p.pad) & 0xffff);
    } else {
      a_pix0[i] = 0;
      a_yx0[i] = (int)0xC0000000;  // iy0 = -16384: always out of bounds -> zero rows
    }
  }

  float4 areg[A_ROWS];
  auto load_A = [&](int kc) {
    const int k0 = kc * BK;
    const int tap = k0 / p.Cin, ci0 = k0 - tap * p.Cin;
    const int ky = tap / p.KW, kx = tap - ky * p.KW;
    const float* src = p.A;
    int ld = p.lda, coff = p.a_coff + g * p.a_gcoff + ci0;
    if (p.A2 != nullptr && ci0 >= p.c_split) { src = p.A2; ld = p.lda2; coff = p.a2_coff + ci0 - p.c_split; }
    coff += a_cg * 4;
#pragma unroll
    for (int i = 0; i < A_ROWS; ++i) {
      const int iy = (a_yx0[i] >> 16) + ky;
      const int ix = (int)(short)(a_yx0[i] & 0xffff) + kx;
      const bool ok = (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
      if (ok) {
        const float* ptr = src + (long long)(a_pix0[i] + iy * p.W + ix) * ld + coff;
        areg[i] = __ldg(reinterpret_cast<const float4*>(ptr));
      } else {
        areg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  auto store_A = [&](int s) {
    __nv_bfloat16* hi = sAhi(s);
    __nv_bfloat16* lo = sAlo(s);
#pragma unroll
    for (int i = 0; i < A_ROWS; ++i) {
      float4 v = areg[i];
      if (p.in_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      uint2 h, l;
      split_bf16x2(v.x, v.y, h.x, l.x);
      split_bf16x2(v.z, v.w, h.y, l.y);
      const int off = ((tid >> 3) + 32 * i) * PITCH + a_cg * 4;
      *reinterpret_cast<uint2*>(hi + off) = h;
      *reinterpret_cast<uint2*>(lo + off) = l;
    }
  };
  auto load_B = [&](int kc, int s) {
    const int k0 = kc * BK;
#pragma unroll
    for (int i = 0; i < B_CHUNKS; ++i) {
      const int c = tid + i * kGemmThreads;     // [plane][n][part]
      const int plane = c / (BN * 4);
      const int r = (c - plane * BN * 4) >> 2, part = c & 3;
      const int n = n0 + r;
      const bool ok = n < p.N;
      const __nv_bfloat16* src = (plane ? Wlo : Whi) + (long long)(ok ? n : 0) * p.K + k0 + part * 8;
      __nv_bfloat16* dst = (plane ? sBlo(s) : sBhi(s)) + r * PITCH + part * 8;
      cp_async16(smem_u32(dst), src, ok);
    }
  };

  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;

  // ---- prologue
  load_A(0);
  load_B(0, 0);
  cp_async_commit();
  store_A(0);
  cp_async_wait<0>();
  __syncthreads();

  for (int kc = 0; kc < nk; ++kc) {
    const int cur = kc & 1, nxt = cur ^ 1;
    const bool more = kc + 1 < nk;
    if (more) {
      load_A(kc + 1);
      load_B(kc + 1, nxt);
      cp_async_commit();
    }
    const uint32_t aHi = smem_u32(sAhi(cur)), aLo = smem_u32(sAlo(cur));
    const uint32_t bHi = smem_u32(sBhi(cur)), bLo = smem_u32(sBlo(cur));
#pragma unroll
    for (int ks = 0; ks < BK / 16; ++ks) {
      uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        const int row = wn * WTN + np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int col = ks * 16 + ((lane >> 3) & 1) * 8;
        const uint32_t off = (uint32_t)(row * PITCH + col) * 2u;
        ldmatrix_x4(bHi + off, bh[2 * np][0], bh[2 * np][1], bh[2 * np + 1][0], bh[2 * np + 1][1]);
        ldmatrix_x4(bLo + off, bl[2 * np][0], bl[2 * np][1], bl[2 * np + 1][0], bl[2 * np + 1][1]);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int row = wm * WTM + mt * 16 + (lane & 15);
        const int col = ks * 16 + (lane >> 4) * 8;
        const uint32_t off = (uint32_t)(row * PITCH + col) * 2u;
        uint32_t ah[4], al[4];
        ldmatrix_x4(aHi + off, ah[0], ah[1], ah[2], ah[3]);
        ldmatrix_x4(aLo + off, al[0], al[1], al[2], al[3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          mma_bf16_16816(acc[mt][nt], al, bh[nt][0], bh[nt][1]);
          mma_bf16_16816(acc[mt][nt], ah, bl[nt][0], bl[nt][1]);
          mma_bf16_16816(acc[mt][nt], ah, bh[nt][0], bh[nt][1]);
        }
      }
    }
    if (more) {
      store_A(nxt);
      cp_async_wait<0>();
    }
    __syncthreads();
  }

  // ---- epilogue (registers -> global, float2 per thread per 8-column tile)
  const float* __restrict__ bias = p.bias ? p.bias + (long long)g * p.bias_gstride : nullptr;
  const int c_coff = p.c_coff + g * p.c_gcoff;
  const int r_coff = p.r_coff + g * p.r_gcoff;
  const int r2_coff = p.r2_coff + g * p.r2_gcoff;
  const float* __restrict__ gamma = p.gamma;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m = m0 + wm * WTM + mt * 16 + half * 8 + (lane >> 2);
      if (m >= M) continue;
      int cls_off = 0;
      if (p.bias_mode == 2) {
        const int rem = m % OHW;
        const int oy = rem / p.OW, ox = rem - oy * p.OW;
        const int ry = oy == 0 ? 0 : (oy == p.OH - 1 ? 2 : 1);
        const int rx = ox == 0 ? 0 : (ox == p.OW - 1 ? 2 : 1);
        cls_off = (ry * 3 + rx) * p.N;
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int n = n0 + wn * WTN + nt * 8 + (lane & 3) * 2;
        if (n >= p.N) continue;
        float v0 = acc[mt][nt][half * 2 + 0], v1 = acc[mt][nt][half * 2 + 1];
        if (p.bias_mode) { v0 += __ldg(bias + cls_off + n); v1 += __ldg(bias + cls_off + n + 1); }
        if (p.act == 1) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
        else if (p.act == 2) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); }
        if (gamma) { v0 *= __ldg(gamma + n); v1 *= __ldg(gamma + n + 1); }
        if (p.res) {
          float2 r = *reinterpret_cast<const float2*>(p.res + (long long)m * p.ldr + r_coff + n);
          if (p.res_relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); }
          v0 += r.x; v1 += r.y;
        }
        if (p.res2) {
          const float2 r = *reinterpret_cast<const float2*>(p.res2 + (long long)m * p.ldr2 + r2_coff + n);
          v0 += r.x; v1 += r.y;
        }
        *reinterpret_cast<float2*>(p.C + (long long)m * p.ldc + c_coff + n) = make_float2(v0, v1);
      }
    }
  }
}

// Host-side launch.  Returns cudaError_t of the launch.
inline const char* conv_gemm_check(const ConvGemmParams& p) {
  if (p.Cin % kGemmBK) return "conv_gemm: Cin must be a multiple of 32";
  if (p.K != p.KH * p.KW * p.Cin) return "conv_gemm: K != KH*KW*Cin";
  if (p.N % 2) return "conv_gemm: N must be even";
  if (p.lda % 4 || p.a_coff % 4 || p.a_gcoff % 4) return "conv_gemm: A channel pitch/offset must be multiples of 4";
  if (p.A2 && (p.lda2 % 4 || p.a2_coff % 4 || p.c_split % kGemmBK)) return "conv_gemm: A2 alignment";
  if (p.ldc % 2 || p.c_coff % 2 || p.c_gcoff % 2) return "conv_gemm: C alignment";
  if (p.res && (p.ldr % 2 || p.r_coff % 2 || p.r_gcoff % 2)) return "conv_gemm: res alignment";
  if (p.res2 && (p.ldr2 % 2 || p.r2_coff % 2 || p.r2_gcoff % 2)) return "conv_gemm: res2 alignment";
  if (p.H >= 16384 || p.W >= 16384) return "conv_gemm: spatial size too large";
  if (p.groups < 1) return "conv_gemm: groups";
  return nullptr;
}

template <int BM, int BN, int WM, int WN>
inline cudaError_t conv_gemm_launch_cfg(const ConvGemmParams& p, cudaStream_t st) {
  static bool configured = false;
  constexpr int smem = conv_gemm_smem_bytes<BM, BN>();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<BM, BN, WM, WN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const long long M = (long long)p.B * p.OH * p.OW;
  dim3 grid((unsigned)cdivl(M, BM), (unsigned)cdiv(p.N, BN), (unsigned)p.groups);
  conv_gemm_kernel<BM, BN, WM, WN><<<grid, kGemmThreads, smem, st>>>(p);
  return cudaGetLastError();
}

// tile configuration chosen for a problem: 0 = 128x128, 1 = 128x64, 2 = 128x32
inline int conv_gemm_config(const ConvGemmParams& p) { return p.N > 64 ? 0 : (p.N > 32 ? 1 : 2); }

inline cudaError_t conv_gemm_launch(const ConvGemmParams& p, cudaStream_t st) {
  if (p.N > 64) return conv_gemm_launch_cfg<128, 128, 2, 4>(p, st);
  if (p.N > 32) return conv_gemm_launch_cfg<128, 64, 4, 2>(p, st);
  return conv_gemm_launch_cfg<128, 32, 8, 1>(p, st);
}

}  // namespace pf
