// Spatial-reduction attention core on the tensor cores: softmax(q k^T / 8) v, 100 keys, head_dim 64
// (mix_transformers.py:127-131), split-precision bf16x3 products (lo*hi + hi*lo + hi*hi, fp32 accumulate) for both
// q k^T and p v, fp32 softmax.  Replaces the CUDA-core attention_kernel (layers.cuh) on the forward path; ~12 % of the
// step there.
//
//   block = 4 warps (3 blocks per SM), one (image, head); K and V of the head are split once into bf16 hi/lo planes in shared memory
//   (keys padded 100 -> 112, rows of 128 B, 16 B chunks XOR-swizzled for conflict-free ldmatrix); the block then loops over
//   passes of 64 queries (16 per warp); passes per block are chosen so that the grid is about one resident wave.  Per warp and tile: S = q k^T (mma.sync.m16n8k16, 4 k-steps x 14 key tiles x 3),
//   row softmax in registers (quad shuffles), O = P V (7 k-steps x 8 tiles x 3; P re-used from the S accumulators as the A
//   operand, V through ldmatrix.trans), normalise, store fp32 and/or split planes.
#pragma once
#include "common.cuh"

namespace pf {

constexpr int kAmKeys = 100, kAmKeysPad = 112, kAmD = 64, kAmQTile = 64, kAmThreads = 128;   // 4 warps x 16 queries per pass
constexpr int kAmPlane = kAmKeysPad * kAmD * 2;      // bytes of one bf16 plane (K or V, hi or lo)
constexpr int kAmSmem = 4 * kAmPlane;                // K_hi, K_lo, V_hi, V_lo = 57344 B

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

// SPLIT_IN: q and kv arrive as bf16 hi/lo planes (written by the q / kv GEMM epilogues): K and V are copied into shared memory
// with cp.async (no conversion work), Q fragments are read as bf16 pairs; the 1/8 scale is applied to S in fp32 (a power of
// two: identical to scaling q).  Otherwise q, kv are fp32 (operator entry point, legacy graph) and are split on the fly.
template <bool SPLIT_IN>
__global__ void __launch_bounds__(kAmThreads) attention_mma_kernel(const float* __restrict__ q, const float* __restrict__ kv,
                                                                   const __nv_bfloat16* __restrict__ q_hi, const __nv_bfloat16* __restrict__ q_lo,
                                                                   const __nv_bfloat16* __restrict__ kv_hi, const __nv_bfloat16* __restrict__ kv_lo,
                                                                   float* __restrict__ out,
                                                                   __nv_bfloat16* __restrict__ shi, __nv_bfloat16* __restrict__ slo, int N, int C,
                                                                   int tiles_per_block) {
  pdl_wait();
  pdl_launch();
  extern __shared__ __align__(128) unsigned char sm_raw[];
  const uint32_t sK_hi = smem_u32(sm_raw), sK_lo = sK_hi + kAmPlane, sV_hi = sK_lo + kAmPlane, sV_lo = sV_hi + kAmPlane;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z, h = blockIdx.y;
  // ---- stage K and V of this (image, head): fp32 -> bf16 hi/lo, [key][64] rows of 128 B, chunk (16 B) index ^= key & 7
  if (SPLIT_IN) {
    // 4 planes (K_hi, K_lo, V_hi, V_lo) x 112 keys x 8 chunks of 16 B
    const long long kvo = (long long)b * kAmKeys * 2 * C + h * kAmD;
    for (int i = tid; i < 4 * kAmKeysPad * 8; i += kAmThreads) {
      const int plane = i / (kAmKeysPad * 8), j = i % (kAmKeysPad * 8), key = j >> 3, c = j & 7;
      const uint32_t dst = sK_hi + plane * kAmPlane + (uint32_t)key * 128u + (uint32_t)((c ^ (key & 7)) << 4);
      const bool valid = key < kAmKeys;    // keys 100..111: zero fill (src-size 0)
      const __nv_bfloat16* src = ((plane & 1) ? kv_lo : kv_hi) + kvo + (long long)(valid ? key : 0) * 2 * C + ((plane >> 1) ? C : 0) + c * 8;
      cp_async16(dst, src, valid);
    }
    cp_async_commit();
    cp_async_wait<0>();
  }
  const float* kvb = SPLIT_IN ? nullptr : kv + (long long)b * kAmKeys * 2 * C + h * kAmD;
  for (int i = tid; !SPLIT_IN && i < kAmKeysPad * (kAmD / 4) * 2; i += kAmThreads) {
    const int isv = i >= kAmKeysPad * (kAmD / 4);
    const int j = isv ? i - kAmKeysPad * (kAmD / 4) : i;
    const int key = j / (kAmD / 4), d4 = j % (kAmD / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (key < kAmKeys) v = __ldg(reinterpret_cast<const float4*>(kvb + (long long)key * 2 * C + (isv ? C : 0) + d4 * 4));
    uint2 hh, ll;
    split_bf16x2(v.x, v.y, hh.x, ll.x);
    split_bf16x2(v.z, v.w, hh.y, ll.y);
    const uint32_t off = (uint32_t)key * 128u + (uint32_t)(((d4 >> 1) ^ (key & 7)) << 4) + (uint32_t)(d4 & 1) * 8u;
    unsigned char* base = sm_raw + (isv ? 2 * kAmPlane : 0);
    *reinterpret_cast<uint2*>(base + off) = hh;
    *reinterpret_cast<uint2*>(base + kAmPlane + off) = ll;
  }
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  for (int it = 0; it < tiles_per_block; ++it) {
    const int q0 = (blockIdx.x * tiles_per_block + it) * kAmQTile + warp * 16;   // first query row of this warp
    if (q0 >= N) break;                                                         // warp-uniform
    const int r0 = q0 + g, r1 = q0 + g + 8;
    // ---- Q fragments (pre-scaled by 1/8, an exact power of two), split into hi / lo
    uint32_t qh[4][4], ql[4][4];
    if (SPLIT_IN) {
      const long long o0 = ((long long)b * N + (r0 < N ? r0 : N - 1)) * C + h * kAmD, o1 = ((long long)b * N + (r1 < N ? r1 : N - 1)) * C + h * kAmD;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int c0 = ks * 16 + 2 * t;
        qh[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(q_hi + o0 + c0));     ql[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(q_lo + o0 + c0));
        qh[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(q_hi + o1 + c0));     ql[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(q_lo + o1 + c0));
        qh[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(q_hi + o0 + c0 + 8)); ql[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(q_lo + o0 + c0 + 8));
        qh[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(q_hi + o1 + c0 + 8)); ql[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(q_lo + o1 + c0 + 8));
      }
    } else {
      const float* q0p = q + ((long long)b * N + (r0 < N ? r0 : N - 1)) * C + h * kAmD;
      const float* q1p = q + ((long long)b * N + (r1 < N ? r1 : N - 1)) * C + h * kAmD;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const float2 a0 = __ldg(reinterpret_cast<const float2*>(q0p + ks * 16 + 2 * t));
        const float2 a1 = __ldg(reinterpret_cast<const float2*>(q1p + ks * 16 + 2 * t));
        const float2 a2 = __ldg(reinterpret_cast<const float2*>(q0p + ks * 16 + 8 + 2 * t));
        const float2 a3 = __ldg(reinterpret_cast<const float2*>(q1p + ks * 16 + 8 + 2 * t));
        split_bf16x2(a0.x * 0.125f, a0.y * 0.125f, qh[ks][0], ql[ks][0]);
        split_bf16x2(a1.x * 0.125f, a1.y * 0.125f, qh[ks][1], ql[ks][1]);
        split_bf16x2(a2.x * 0.125f, a2.y * 0.125f, qh[ks][2], ql[ks][2]);
        split_bf16x2(a3.x * 0.125f, a3.y * 0.125f, qh[ks][3], ql[ks][3]);
      }
    }
    // ---- S = Q K^T : 16 x 112
    float s[14][4];
#pragma unroll
    for (int nt = 0; nt < 14; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 7; ++np) {
        // two key tiles (16 keys) x 16 d: matrices (keys 0-7, d 0-7), (keys 0-7, d 8-15), (keys 8-15, d 0-7), (keys 8-15, d 8-15)
        const int key = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int chunk = ks * 2 + ((lane >> 3) & 1);
        const uint32_t off = (uint32_t)key * 128u + (uint32_t)((chunk ^ (key & 7)) << 4);
        uint32_t bh0, bh1, bh2, bh3, bl0, bl1, bl2, bl3;
        ldmatrix_x4(sK_hi + off, bh0, bh1, bh2, bh3);
        ldmatrix_x4(sK_lo + off, bl0, bl1, bl2, bl3);
        mma_bf16_16816(s[2 * np], ql[ks], bh0, bh1);
        mma_bf16_16816(s[2 * np], qh[ks], bl0, bl1);
        mma_bf16_16816(s[2 * np], qh[ks], bh0, bh1);
        mma_bf16_16816(s[2 * np + 1], ql[ks], bh2, bh3);
        mma_bf16_16816(s[2 * np + 1], qh[ks], bl2, bl3);
        mma_bf16_16816(s[2 * np + 1], qh[ks], bh2, bh3);
      }
    }
    // ---- softmax over the 100 valid keys (rows r0: regs 0,1 ; r1: regs 2,3), fp32
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 14; ++nt) {
      const int k0 = nt * 8 + 2 * t;
      if (SPLIT_IN) { s[nt][0] *= 0.125f; s[nt][1] *= 0.125f; s[nt][2] *= 0.125f; s[nt][3] *= 0.125f; }
      if (k0 >= kAmKeys) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
      if (k0 + 1 >= kAmKeys) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
      m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 14; ++nt) {
      s[nt][0] = expf(s[nt][0] - m0); s[nt][1] = expf(s[nt][1] - m0);
      s[nt][2] = expf(s[nt][2] - m1); s[nt][3] = expf(s[nt][3] - m1);
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // ---- O = P V : 16 x 64, P taken from the S accumulators (A fragment of k-step j = key tiles 2j, 2j+1)
    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      uint32_t ph[4], pl[4];
      split_bf16x2(s[2 * j][0], s[2 * j][1], ph[0], pl[0]);
      split_bf16x2(s[2 * j][2], s[2 * j][3], ph[1], pl[1]);
      split_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1], ph[2], pl[2]);
      split_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3], ph[3], pl[3]);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        // V[key][d] rows = k: transposed 8x8 loads: (k 0-7, d 0-7), (k 8-15, d 0-7), (k 0-7, d 8-15), (k 8-15, d 8-15)
        const int key = j * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int chunk = np * 2 + (lane >> 4);
        const uint32_t off = (uint32_t)key * 128u + (uint32_t)((chunk ^ (key & 7)) << 4);
        uint32_t vh0, vh1, vh2, vh3, vl0, vl1, vl2, vl3;
        ldmatrix_x4_trans(sV_hi + off, vh0, vh1, vh2, vh3);
        ldmatrix_x4_trans(sV_lo + off, vl0, vl1, vl2, vl3);
        mma_bf16_16816(o[2 * np], pl, vh0, vh1);
        mma_bf16_16816(o[2 * np], ph, vl0, vl1);
        mma_bf16_16816(o[2 * np], ph, vh0, vh1);
        mma_bf16_16816(o[2 * np + 1], pl, vh2, vh3);
        mma_bf16_16816(o[2 * np + 1], ph, vl2, vl3);
        mma_bf16_16816(o[2 * np + 1], ph, vh2, vh3);
      }
    }
    // ---- normalise and store (row r0: regs 0,1 ; row r1: regs 2,3 ; columns nt*8 + 2t, +1)
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int d = nt * 8 + 2 * t;
      if (r0 < N) {
        const long long oi = ((long long)b * N + r0) * C + h * kAmD + d;
        const float x = o[nt][0] * i0, y = o[nt][1] * i0;
        if (out) *reinterpret_cast<float2*>(out + oi) = make_float2(x, y);
        if (shi) { uint32_t hh, ll; split_bf16x2(x, y, hh, ll); *reinterpret_cast<uint32_t*>(shi + oi) = hh; *reinterpret_cast<uint32_t*>(slo + oi) = ll; }
      }
      if (r1 < N) {
        const long long oi = ((long long)b * N + r1) * C + h * kAmD + d;
        const float x = o[nt][2] * i1, y = o[nt][3] * i1;
        if (out) *reinterpret_cast<float2*>(out + oi) = make_float2(x, y);
        if (shi) { uint32_t hh, ll; split_bf16x2(x, y, hh, ll); *reinterpret_cast<uint32_t*>(shi + oi) = hh; *reinterpret_cast<uint32_t*>(slo + oi) = ll; }
      }
    }
  }
}

inline cudaError_t attention_mma_configure_device() {   // per-device shared-memory opt-in (pf_create)
  cudaError_t e = cudaFuncSetAttribute(attention_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAmSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAmSmem);
  return e;
}

// q / kv: fp32 pointers, or (qs / kvs non-empty) split planes with row pitch C / 2C
inline cudaError_t attention_mma_launch(const float* q, const float* kv, float* out, int B, int N, int C, int heads, cudaStream_t st, SplitT sp = SplitT(),
                                        SplitT qs = SplitT(), SplitT kvs = SplitT()) {
  if ((qs.hi != nullptr) != (kvs.hi != nullptr) || (qs.hi && (qs.ld != C || kvs.ld != 2 * C))) return cudaErrorInvalidValue;
  const int tiles = cdiv(N, kAmQTile);
  // passes per block: the grid should be about one resident wave (148 SMs x 3 blocks); the K/V staging of a block is
  // amortised over tpb * 64 queries
  int tpb = (tiles * heads * B) / 400;
  tpb = tpb < 1 ? 1 : (tpb > tiles ? tiles : tpb);
  dim3 grid(cdiv(tiles, tpb), heads, B);
  if (qs.hi) return launch_pdl(attention_mma_kernel<true>, grid, dim3(kAmThreads), kAmSmem, st, nullptr, nullptr, qs.hi, qs.lo, kvs.hi, kvs.lo, out, sp.hi, sp.lo, N, C, tpb);
  return launch_pdl(attention_mma_kernel<false>, grid, dim3(kAmThreads), kAmSmem, st, q, kv, nullptr, nullptr, nullptr, nullptr, out, sp.hi, sp.lo, N, C, tpb);
}

}  // namespace pf
