// Spatial-reduction attention core on tcgen05 / TMEM: softmax(q k^T / 8) v, 100 keys, head_dim 64 (mix_transformers.py:127-131),
// split-precision bf16x3 products for both contractions, fp32 softmax.  Replaces the warp-level mma.sync kernel
// (attention_mma.cuh, kept as option "attn_tc" = 0).  q / kv arrive as bf16 hi/lo planes from their GEMM epilogues.
//
// Work item = 128 queries of one (image, head).  Per item, all operands K-major / SWIZZLE_128B in shared memory:
//   S[128 x 112] = Q[128 x 64] K[112 x 64]^T         (TMA boxes; keys 100..111 are masked)        -> TMEM, 3 MMAs x 4 K-steps
//   softmax over the row: one THREAD per query row reads its 112 scores from TMEM (tcgen05.ld), no shuffles
//   P (unnormalised, bf16 hi/lo) -> shared memory as the A operand [128 x 112 keys]
//   O[128 x 64] = P V, V consumed as an MN-major B operand (the [key][d] tile exactly as TMA delivers it) -> TMEM, 3 x 7 MMAs
//   O / rowsum -> bf16 hi/lo planes of the attention output.
//
// Pipeline (persistent CTA over a contiguous range of items, so that K / V of an (image, head) are loaded once):
//   warp 0  TMA producer: Q per item (single buffer, freed by the commit of S), K / V per (image, head) (double buffered)
//   warp 1  MMA issuer, order S(0), S(1), PV(0), S(2), PV(1), ...: the tensor core computes S(i+1) while a softmax group works on i
//   warp 2  TMEM allocator (512 columns: two S slots of 128, two O slots of 64)
//   warps 4-7 / 8-11  two softmax + epilogue groups (TMEM lane quarter = warp % 4), alternating items
#pragma once
#include "gemm_tma.cuh"

namespace pf {

constexpr int kAtcKeys = 100, kAtcKeysPad = 112, kAtcD = 64, kAtcThreads = 384;
constexpr int kAtcQPlane = 128 * 128;                    // 128 rows x 128 B
constexpr int kAtcKPlane = kAtcKeysPad * 128;            // 112 rows x 128 B = 14336 (a multiple of 1024)
constexpr int kAtcPPlane = 2 * 128 * 128;                // two 64-key sub-tiles of [128 rows x 128 B]
constexpr int kAtcSmemQ = 0, kAtcSmemK = kAtcSmemQ + 2 * kAtcQPlane;            // K: 2 slots x (hi, lo)
constexpr int kAtcSmemV = kAtcSmemK + 4 * kAtcKPlane, kAtcSmemP = kAtcSmemV + 4 * kAtcKPlane;
constexpr int kAtcSmemBars = kAtcSmemP + 2 * kAtcPPlane;
constexpr int kAtcSmemBytes = kAtcSmemBars + 256 + 1024;

struct AtcMaps { CUtensorMap q_hi, q_lo, kv_hi, kv_lo; };

// instruction descriptors: M = 128, fp32 accumulate, bf16 operands; S: N = 112, both K-major; PV: N = 64, B MN-major (bit 16)
constexpr uint32_t kAtcIdescS = umma_idesc_bf16(kAtcKeysPad);
constexpr uint32_t kAtcIdescPV = umma_idesc_bf16(kAtcD) | (1u << 16);

__global__ void __launch_bounds__(kAtcThreads, 1) attention_tc_kernel(const __grid_constant__ AtcMaps maps, __nv_bfloat16* __restrict__ ohi,
                                                                      __nv_bfloat16* __restrict__ olo, int B, int N, int C, int heads, int total_items,
                                                                      int items_per_cta) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (sbase - raw);
  const uint32_t sQ = sbase + kAtcSmemQ, sK = sbase + kAtcSmemK, sV = sbase + kAtcSmemV, sP = sbase + kAtcSmemP, bars = sbase + kAtcSmemBars;
  const uint32_t q_full = bars, q_free = bars + 8, p_full = bars + 16, p_free = bars + 24;
  auto s_full = [&](int w) { return bars + 32u + 8u * w; };
  auto s_free = [&](int w) { return bars + 48u + 8u * w; };
  auto o_full = [&](int w) { return bars + 64u + 8u * w; };
  auto o_free = [&](int w) { return bars + 80u + 8u * w; };
  auto kv_full = [&](int s) { return bars + 96u + 8u * s; };
  auto kv_free = [&](int s) { return bars + 112u + 8u * s; };
  const uint32_t tmem_slot = bars + 128;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int QT = cdiv(N, 128);
  const int i0 = blockIdx.x * items_per_cta;
  const int i1 = min(i0 + items_per_cta, total_items);
  const int n_items = i1 - i0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.q_hi); tma_prefetch_desc(&maps.q_lo); tma_prefetch_desc(&maps.kv_hi); tma_prefetch_desc(&maps.kv_lo);
    mbar_init(q_full, 1); mbar_init(q_free, 1); mbar_init(p_full, 128); mbar_init(p_free, 1);
    for (int w = 0; w < 2; ++w) {
      mbar_init(s_full(w), 1); mbar_init(s_free(w), 128); mbar_init(o_full(w), 1); mbar_init(o_free(w), 128);
      mbar_init(kv_full(w), 1); mbar_init(kv_free(w), 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + (tmem_slot - sbase));
  pdl_wait();
  pdl_launch();
  if (n_items <= 0) {   // (grids are sized so that this does not happen; keep the teardown uniform)
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 512);
    return;
  }

  // item t -> (image * heads + head, query tile); the (image, head) "group" index of the items this CTA owns
  auto group_of = [&](int t) { return t / QT; };
  const int g0 = group_of(i0);

  if (warp == 0) {
    // ======================================================================= TMA producer
    if (lane == 0) {
      int cur_g = -1;
      for (int i = 0; i < n_items; ++i) {
        const int t = i0 + i, g = group_of(t), qt = t - g * QT;
        const int b = g / heads, h = g - b * heads;
        if (g != cur_g) {
          cur_g = g;
          const int gi = g - g0, slot = gi & 1, u = gi >> 1;
          mbar_wait(kv_free(slot), (u & 1) ^ 1);
          mbar_expect_tx(kv_full(slot), 4 * kAtcKPlane);
          tma_load_2d(sK + (2 * slot) * kAtcKPlane, &maps.kv_hi, kv_full(slot), h * kAtcD, b * kAtcKeys);
          tma_load_2d(sK + (2 * slot + 1) * kAtcKPlane, &maps.kv_lo, kv_full(slot), h * kAtcD, b * kAtcKeys);
          tma_load_2d(sV + (2 * slot) * kAtcKPlane, &maps.kv_hi, kv_full(slot), C + h * kAtcD, b * kAtcKeys);
          tma_load_2d(sV + (2 * slot + 1) * kAtcKPlane, &maps.kv_lo, kv_full(slot), C + h * kAtcD, b * kAtcKeys);
        }
        mbar_wait(q_free, (i & 1) ^ 1);
        mbar_expect_tx(q_full, 2 * kAtcQPlane);
        tma_load_2d(sQ, &maps.q_hi, q_full, h * kAtcD, b * N + qt * 128);
        tma_load_2d(sQ + kAtcQPlane, &maps.q_lo, q_full, h * kAtcD, b * N + qt * 128);
      }
    }
  } else if (warp == 1) {
    // ======================================================================= MMA issuer
    auto issue_pv = [&](int i) {   // O(i) = P(i) V(group of i)
      const int w = i & 1, j = i >> 1;
      const int gi = group_of(i0 + i) - g0, slot = gi & 1;
      mbar_wait(p_full, i & 1);
      mbar_wait(o_free(w), (j & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t acc = tmem + 256u + (uint32_t)(w * 64);
        const uint32_t v_hi = sV + (2 * slot) * kAtcKPlane, v_lo = v_hi + kAtcKPlane;
#pragma unroll
        for (int kk = 0; kk < kAtcKeysPad / 16; ++kk) {
          const uint32_t p_hi = sP + (kk >> 2) * (128 * 128) + (kk & 3) * 32, p_lo = p_hi + kAtcPPlane;
          const uint64_t dph = tma_tile_desc<64>(p_hi), dpl = tma_tile_desc<64>(p_lo);
          // V[key][d] as an MN-major B operand: 16 keys of this K step = two 8-row groups 1024 B apart
          const uint64_t dvh = tma_tile_desc<64>(v_hi + kk * 2048), dvl = tma_tile_desc<64>(v_lo + kk * 2048);
          umma_bf16(acc, dpl, dvh, kAtcIdescPV, kk ? 1u : 0u);
          umma_bf16(acc, dph, dvl, kAtcIdescPV, 1u);
          umma_bf16(acc, dph, dvh, kAtcIdescPV, 1u);
        }
        umma_commit(p_free);
        umma_commit(o_full(w));
        const bool last_of_group = (i == n_items - 1) || (group_of(i0 + i + 1) != group_of(i0 + i));
        if (last_of_group) umma_commit(kv_free(slot));
      }
      __syncwarp();
    };
    for (int i = 0; i < n_items; ++i) {
      const int w = i & 1, j = i >> 1;
      const int gi = group_of(i0 + i) - g0, slot = gi & 1, u = gi >> 1;
      const bool first_of_group = (i == 0) || (group_of(i0 + i - 1) != group_of(i0 + i));
      mbar_wait(q_full, i & 1);
      if (first_of_group) mbar_wait(kv_full(slot), u & 1);
      mbar_wait(s_free(w), (j & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t acc = tmem + (uint32_t)(w * 128);
        const uint32_t k_hi = sK + (2 * slot) * kAtcKPlane, k_lo = k_hi + kAtcKPlane;
        uint64_t dqh = tma_tile_desc<64>(sQ), dql = tma_tile_desc<64>(sQ + kAtcQPlane);
        uint64_t dkh = tma_tile_desc<64>(k_hi), dkl = tma_tile_desc<64>(k_lo);
#pragma unroll
        for (int kk = 0; kk < kAtcD / 16; ++kk) {
          umma_bf16(acc, dql, dkh, kAtcIdescS, kk ? 1u : 0u);
          umma_bf16(acc, dqh, dkl, kAtcIdescS, 1u);
          umma_bf16(acc, dqh, dkh, kAtcIdescS, 1u);
          dqh += 2; dql += 2; dkh += 2; dkl += 2;
        }
        umma_commit(q_free);
        umma_commit(s_full(w));
      }
      __syncwarp();
      if (i >= 1) issue_pv(i - 1);
    }
    issue_pv(n_items - 1);
  } else if (warp >= 4) {
    // ======================================================================= softmax + epilogue groups
    const int w = (warp - 4) >> 2;                 // group 0: warps 4-7, group 1: warps 8-11
    const int q = warp & 3;                        // TMEM lane quarter
    const int r = q * 32 + lane;                   // row of the tile = query
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    constexpr float kScale = 0.125f * 1.4426950408889634f;     // 1/8 and log2(e): p = 2^((s - max) * kScale)
    for (int i = w; i < n_items; i += 2) {
      const int j = i >> 1;
      const int t = i0 + i, g = group_of(t), qt = t - g * QT;
      const int b = g / heads, h = g - b * heads;
      mbar_wait(s_full(w), j & 1);
      tc_fence_after();
      float p[kAtcKeysPad];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + (uint32_t)(w * 128 + c * 32), v);
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int k = c * 32 + e;
          if (k < kAtcKeysPad) {
            const float sv = k < kAtcKeys ? __uint_as_float(v[e]) : -INFINITY;
            p[k] = sv;
            mx = fmaxf(mx, sv);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(s_free(w));                      // this thread has read its row of S
      const float mc = mx * kScale;
      float l = 0.f;
#pragma unroll
      for (int k = 0; k < kAtcKeysPad; ++k) {
        p[k] = k < kAtcKeys ? ex2_approx(fmaf(p[k], kScale, -mc)) : 0.f;
        l += p[k];
      }
      // P -> shared memory (A operand of the second contraction), once the MMAs of the previous item have consumed the buffer
      mbar_wait(p_free, (i & 1) ^ 1);
      unsigned char* pp = sm + (sP - sbase);
#pragma unroll
      for (int k8 = 0; k8 < kAtcKeysPad / 8; ++k8) {
        uint4 hh, ll;
        split_bf16x2(p[8 * k8], p[8 * k8 + 1], hh.x, ll.x); split_bf16x2(p[8 * k8 + 2], p[8 * k8 + 3], hh.y, ll.y);
        split_bf16x2(p[8 * k8 + 4], p[8 * k8 + 5], hh.z, ll.z); split_bf16x2(p[8 * k8 + 6], p[8 * k8 + 7], hh.w, ll.w);
        const int off = (k8 >> 3) * (128 * 128) + r * 128 + (((k8 & 7) ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(pp + off) = hh;
        *reinterpret_cast<uint4*>(pp + kAtcPPlane + off) = ll;
      }
      fence_proxy_async_smem();
      mbar_arrive(p_full);
      // O / l -> global
      mbar_wait(o_full(w), j & 1);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld32(tmem + lane_addr + 256u + (uint32_t)(w * 64), o0);
      tmem_ld32(tmem + lane_addr + 256u + (uint32_t)(w * 64 + 32), o1);
      tc_fence_before();
      mbar_arrive(o_free(w));
      const int qrow = qt * 128 + r;
      if (qrow < N) {
        const float inv = 1.0f / l;
        const long long oi = ((long long)b * N + qrow) * C + h * kAtcD;
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint4 hh, ll;
          split_bf16x2(__uint_as_float(o0[e]) * inv, __uint_as_float(o0[e + 1]) * inv, hh.x, ll.x);
          split_bf16x2(__uint_as_float(o0[e + 2]) * inv, __uint_as_float(o0[e + 3]) * inv, hh.y, ll.y);
          split_bf16x2(__uint_as_float(o0[e + 4]) * inv, __uint_as_float(o0[e + 5]) * inv, hh.z, ll.z);
          split_bf16x2(__uint_as_float(o0[e + 6]) * inv, __uint_as_float(o0[e + 7]) * inv, hh.w, ll.w);
          *reinterpret_cast<uint4*>(ohi + oi + e) = hh;
          *reinterpret_cast<uint4*>(olo + oi + e) = ll;
        }
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint4 hh, ll;
          split_bf16x2(__uint_as_float(o1[e]) * inv, __uint_as_float(o1[e + 1]) * inv, hh.x, ll.x);
          split_bf16x2(__uint_as_float(o1[e + 2]) * inv, __uint_as_float(o1[e + 3]) * inv, hh.y, ll.y);
          split_bf16x2(__uint_as_float(o1[e + 4]) * inv, __uint_as_float(o1[e + 5]) * inv, hh.z, ll.z);
          split_bf16x2(__uint_as_float(o1[e + 6]) * inv, __uint_as_float(o1[e + 7]) * inv, hh.w, ll.w);
          *reinterpret_cast<uint4*>(ohi + oi + 32 + e) = hh;
          *reinterpret_cast<uint4*>(olo + oi + 32 + e) = ll;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

}  // namespace pf
