// CTA-pair version of the GEMM mode of gemm_tma.cuh: C[M,N] = A[M,K] W[N,K]^T on tcgen05.mma.cta_group::2.
//
// Why: with 128 x BN tiles per CTA the GEMM mode is bound by L2 -> shared-memory traffic, not by the tensor pipe -- every K step
// of 32 moves (128 + BN) x 32 x 4 B of bf16 hi/lo operands per SM for 3 x 2 MMAs, 64 B/clk/SM at BN = 256, while the TMA path
// of this part sustains ~42 B/clk/SM (measured on three kernels, profiles/r02_notes.md).  A CTA PAIR (two SMs of one TPC,
// cluster 2x1x1) computes a 256 x BN tile with ONE instruction stream: each CTA loads its own 128 rows of A and only HALF of the
// weight tile (BN/2 rows); tcgen05.mma.cta_group::2 reads both halves (the peer's through distributed shared memory) and
// writes each CTA's 128 accumulator rows into its own TMEM.  Per SM and K step: (128 + BN/2) x 32 x 4 B -- 43 B/clk at BN = 256.
//
// Protocol (same roles as gemm_tma.cuh; rank 0 of the pair = leader):
//   warp 0 (both CTAs): TMA producer.  Waits its own empty[s]; issues cp.async.bulk.tensor ... cta_group::2 into its own shared
//           memory with the LEADER's full[s] as the completion barrier.  Only the leader arms full[s] (expect_tx = both CTAs' bytes).
//   warp 1 (leader only): waits full[s], issues the MMAs (instruction descriptor M = 256), tcgen05.commit.cta_group::2 multicast
//           to empty[s] of BOTH CTAs; after the last K step a multicast commit to tmem_full[acc] of both CTAs.
//   warp 2 (both): tcgen05.alloc.cta_group::2 / dealloc (2 x BN columns: two accumulators).
//   warps 4-11 (both): epilogue of the CTA's own 128 rows (identical to gemm_tma.cuh's TMA-store epilogue); each warp then
//           arrives on the LEADER's tmem_empty[acc] (count 16), remotely from rank 1.
#pragma once
#include "gemm_tma.cuh"

namespace pf {

template <int BN> struct Tma2Cfg {
  static constexpr int KB = 32;
  static constexpr int kAPlane = 128 * KB * 2;
  static constexpr int kBPlane = (BN / 2) * KB * 2;                // this CTA's half of the N tile
  static constexpr int kStage = 2 * kAPlane + 2 * kBPlane;
  static constexpr int kEpiStage = 4 * 16384, kEpiVec = 8 * 2 * 256 * 4;
  static constexpr int kBudget = 225 * 1024 - kEpiStage - kEpiVec;
  static constexpr int kStagesRaw = kBudget / kStage;
  static constexpr int kStages = kStagesRaw > 12 ? 12 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStage + kEpiStage + kEpiVec + 512 + 1024;
  static constexpr int kTmemCols = 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  // instruction descriptor as umma_idesc_bf16, M = 256 (the pair)
  static constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  static_assert(kStages >= 3, "ring too shallow");
};

// ------------------------------------------------------------------------------------------------ cluster / cta_group::2 PTX
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory, completion bytes signalled on the barrier at `bar_cluster` (a shared::cluster address:
// the leader CTA's barrier)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_slot), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (once) on the barrier at this shared-memory offset in every CTA of `mask` when the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(bar), "h"(mask) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(kTmaThreads, 1) gemm2_tma_kernel(const __grid_constant__ TmaMaps maps, const TmaGemmParams p) {
  using Cfg = Tma2Cfg<BN>;
  constexpr int NS = Cfg::kStages, KB = Cfg::KB;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (sbase - raw);
  const uint32_t ring = sbase;
  const uint32_t epi_base = ring + NS * Cfg::kStage;
  const uint32_t vec_base = epi_base + Cfg::kEpiStage;
  const uint32_t bars = vec_base + Cfg::kEpiVec;
  auto full_b = [&](int s) { return bars + 8u * s; };
  auto empty_b = [&](int s) { return bars + 8u * (NS + s); };
  auto tmem_full = [&](int i) { return bars + 8u * (2 * NS + i); };
  auto tmem_empty = [&](int i) { return bars + 8u * (2 * NS + 2 + i); };
  const uint32_t tmem_slot = bars + 8u * (2 * NS + 4);
  auto res_bar = [&](int w, int i) { return bars + 8u * (2 * NS + 5 + 2 * w + i); };

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_tiles = cdiv(p.N, BN);
  const int m_pairs = cdiv(p.M, 256);
  const int total_tiles = m_pairs * n_tiles;
  const int nk = p.K / KB;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a_hi); tma_prefetch_desc(&maps.a_lo); tma_prefetch_desc(&maps.b_hi); tma_prefetch_desc(&maps.b_lo);
    for (int s = 0; s < NS; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tmem_full(i), 1); mbar_init(tmem_empty(i), 16); }
    for (int i = 0; i < 16; ++i) mbar_init(res_bar(i >> 1, i & 1), 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrive / multicast commit / peer TMA signal
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + (tmem_slot - sbase));
  pdl_wait();
  pdl_launch();

  auto decode = [&](int tile, int& mp, int& n0) {   // n fastest: pairs running together share the A panel in L2
    n0 = (tile % n_tiles) * BN;
    mp = tile / n_tiles;
  };

  if (warp == 0) {
    // ======================================================================= TMA producer (both CTAs)
    if (lane == 0) {
      int it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
        int mp, n0;
        decode(tile, mp, n0);
        const int arow = mp * 256 + (int)rank * 128;
        const int brow = p.b_row0 + n0 + (int)rank * (BN / 2);
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int s = it % NS;
          mbar_wait(empty_b(s), ((it / NS) & 1) ^ 1);
          if (leader) mbar_expect_tx(full_b(s), 2 * Cfg::kStage);
          const uint32_t lead_full = mapa_u32(full_b(s), 0);
          const uint32_t st = ring + s * Cfg::kStage;
          const int kcol = kc * KB;
          tma_load_2d_2sm(st, &maps.a_hi, lead_full, p.a_c0 + kcol, arow);
          tma_load_2d_2sm(st + Cfg::kAPlane, &maps.a_lo, lead_full, p.a_c0 + kcol, arow);
          tma_load_2d_2sm(st + 2 * Cfg::kAPlane, &maps.b_hi, lead_full, kcol, brow);
          tma_load_2d_2sm(st + 2 * Cfg::kAPlane + Cfg::kBPlane, &maps.b_lo, lead_full, kcol, brow);
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================================= MMA issuer (leader CTA only)
    if (leader) {
      int it = 0, tl = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += nclusters, ++tl) {
        const int as = tl & 1;
        mbar_wait(tmem_empty(as), ((tl >> 1) & 1) ^ 1);      // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem + (uint32_t)(as * BN);
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int s = it % NS;
          mbar_wait(full_b(s), (it / NS) & 1);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_hi = ring + s * Cfg::kStage;
            const uint32_t b_hi = a_hi + 2 * Cfg::kAPlane;
            uint64_t dah = tma_tile_desc<KB>(a_hi);
            uint64_t dal = dah + (uint64_t)(Cfg::kAPlane >> 4);
            uint64_t dbh = tma_tile_desc<KB>(b_hi);
            uint64_t dbl = dbh + (uint64_t)(Cfg::kBPlane >> 4);
#pragma unroll
            for (int kk = 0; kk < KB / 16; ++kk) {
              umma_bf16_2sm(acc, dal, dbh, Cfg::kIdesc, (kc | kk) ? 1u : 0u);
              umma_bf16_2sm(acc, dah, dbl, Cfg::kIdesc, 1u);
              umma_bf16_2sm(acc, dah, dbh, Cfg::kIdesc, 1u);
              dah += 2; dal += 2; dbh += 2; dbl += 2;
            }
            umma_commit_2sm(empty_b(s), 3);
            if (kc == nk - 1) umma_commit_2sm(tmem_full(as), 3);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ======================================================================= epilogue (both CTAs, own 128 rows): TMA loads / stores
    const bool has_res = p.res != nullptr;
    const int q = warp & 3;
    const int ew = warp - 4;
    const int ch0 = ew >> 2;
    const uint32_t stg = epi_base + ew * 8192;
    unsigned char* stg_p = sm + (stg - sbase);
    float* bias_s = reinterpret_cast<float*>(sm + (vec_base - sbase)) + ew * 512;
    float* gamma_s = bias_s + 256;
    uint32_t cc = 0;
    int tl = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += nclusters, ++tl) {
      int mp, n0;
      decode(tile, mp, n0);
      const int as = tl & 1;
      const int row0 = mp * 256 + (int)rank * 128 + q * 32;
      const int nch = ((p.N - n0 < BN ? p.N - n0 : BN) + 31) / 32;
      for (int j = lane; j < BN; j += 32) {
        const bool ok = n0 + j < p.N;
        bias_s[j] = (p.bias_mode && ok) ? __ldg(p.bias + n0 + j) : 0.f;
        gamma_s[j] = (p.gamma && ok) ? __ldg(p.gamma + n0 + j) : 1.f;
      }
      __syncwarp();
      const bool warp_active = row0 < p.M && ch0 < nch;
      if (has_res && warp_active && lane == 0) {
        bulk_wait_read<0>();
        mbar_expect_tx(res_bar(ew, cc & 1), 4096);
        tma_load_2d(stg + (cc & 1) * 4096, &maps.res, res_bar(ew, cc & 1), p.r_coff + n0 + ch0 * 32, row0);
      }
      mbar_wait(tmem_full(as), (tl >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int ch = ch0; ch < nch; ch += 2) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + ch * 32), v);
        if (!warp_active) continue;
        const int ob = cc & 1;
        if (lane == 0) {
          if (has_res) {
            bulk_wait_read<0>();
            if (ch + 2 < nch) {
              mbar_expect_tx(res_bar(ew, ob ^ 1), 4096);
              tma_load_2d(stg + (ob ^ 1) * 4096, &maps.res, res_bar(ew, ob ^ 1), p.r_coff + n0 + (ch + 2) * 32, row0);
            }
          } else {
            bulk_wait_read<1>();
          }
        }
        __syncwarp();
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bv = *reinterpret_cast<const float4*>(bias_s + ch * 32 + j);
          o[j] = __uint_as_float(v[j]) + bv.x; o[j + 1] = __uint_as_float(v[j + 1]) + bv.y;
          o[j + 2] = __uint_as_float(v[j + 2]) + bv.z; o[j + 3] = __uint_as_float(v[j + 3]) + bv.w;
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = fmaxf(o[j], 0.f);
        } else if (p.act == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = gelu_erf(o[j]);
        }
        if (p.gamma) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 gv = *reinterpret_cast<const float4*>(gamma_s + ch * 32 + j);
            o[j] *= gv.x; o[j + 1] *= gv.y; o[j + 2] *= gv.z; o[j + 3] *= gv.w;
          }
        }
        unsigned char* ob_p = stg_p + ob * 4096;
        if (has_res) {
          mbar_wait(res_bar(ew, ob), (cc >> 1) & 1);
          const unsigned char* rb = ob_p + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 rv = *reinterpret_cast<const float4*>(rb + ((j ^ (lane & 7)) << 4));
            if (p.res_relu) { rv.x = fmaxf(rv.x, 0.f); rv.y = fmaxf(rv.y, 0.f); rv.z = fmaxf(rv.z, 0.f); rv.w = fmaxf(rv.w, 0.f); }
            o[4 * j] += rv.x; o[4 * j + 1] += rv.y; o[4 * j + 2] += rv.z; o[4 * j + 3] += rv.w;
          }
          __syncwarp();
        }
        ++cc;
        if (p.C) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(ob_p + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 h, l;
            float t[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) t[e] = p.split_relu ? fmaxf(o[8 * j + e], 0.f) : o[8 * j + e];
            split_bf16x2(t[0], t[1], h.x, l.x); split_bf16x2(t[2], t[3], h.y, l.y);
            split_bf16x2(t[4], t[5], h.z, l.z); split_bf16x2(t[6], t[7], h.w, l.w);
            const int off = lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
            *reinterpret_cast<uint4*>(ob_p + off) = h;
            *reinterpret_cast<uint4*>(ob_p + 2048 + off) = l;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int col = n0 + ch * 32;
          if (p.C) tma_store_2d(&maps.c, stg + ob * 4096, p.c_coff + col, row0);
          else { tma_store_2d(&maps.s_hi, stg + ob * 4096, p.s_coff + col, row0); tma_store_2d(&maps.s_lo, stg + ob * 4096 + 2048, p.s_coff + col, row0); }
          bulk_commit();
        }
      }
      // this warp has read its part of the accumulator: one arrival per warp on the LEADER's barrier (8 warps x 2 CTAs)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(tmem_empty(as), 0));
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer may still read this CTA's shared memory (operand halves) / signal its barriers
  if (warp == 2) tmem_dealloc_2sm(tmem, Cfg::kTmemCols);
}

}  // namespace pf
