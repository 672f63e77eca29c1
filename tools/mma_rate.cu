// mma_rate.cu -- how long does one tcgen05.mma (kind::f16, bf16, M = 128, K = 16, operands in shared memory, SWIZZLE_128B K-major)
// take as a function of N, when consecutive MMAs accumulate into the SAME TMEM accumulator and when they rotate over 2 / 4
// accumulators?  One CTA per SM, one thread issues `iters` MMAs back to back, one commit, clock64 around.  Operand contents are
// irrelevant (zeros).  Experiment infrastructure (profiles/r02_notes.md), not part of the library.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I perspectivefields_b200/csrc -o tools/mma_rate tools/mma_rate.cu && tools/mma_rate
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace pf;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr, uint32_t sbo = 1024) {   // 128-byte rows, 8-row groups `sbo` bytes apart
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

template <int NACCS, int NOPS>
__global__ void __launch_bounds__(384, 1) rate(int n, int iters, long long* out, int a_sbo, int a_off, int commit_every = 0, int spinners = 0) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn), sbase = (raw + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (sbase - raw);
  // A, B: 64 KB regions each (contents irrelevant); NOPS operand sets 8 KB apart vary the operand addresses
  const uint32_t a0 = sbase, b0 = sbase + 65536, bar = sbase + 2 * 65536, slot = bar + 8, bar2 = bar + 16, bar3 = bar + 24;
  for (int i = threadIdx.x; i < (2 * 65536) / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); mbar_init(bar3, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + (slot - sbase));
  const uint32_t idesc = umma_idesc_bf16(n);
  if (threadIdx.x == 0) {
    uint64_t da[8], db[8];
    uint32_t acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {     // the issue loop below has no address arithmetic: everything is precomputed
      da[u] = desc_sw128(a0 + (u % NOPS) * 8192 + a_off * (u % 3) , a_sbo) + 2 * (u & 3);   // a_off: tap-like shifted start rows
      db[u] = desc_sw128(b0 + (u % NOPS) * 8192) + 2 * (u & 3);
      acc[u] = tmem + (uint32_t)((u % NACCS) * n);
    }
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) umma_bf16(acc[u], da[u], db[u], idesc, 1u);
      if (commit_every && ((i >> 3) + 1) % commit_every == 0) umma_commit(bar2);      // a pipeline-stage release nobody waits for
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
    mbar_arrive(bar3);                                                                // release the spinning warps
  } else if ((int)(threadIdx.x >> 5) >= 1 && (int)(threadIdx.x >> 5) <= spinners) {
    mbar_wait(bar3, 0);            // warps that wait on a barrier for the whole run, like the epilogue warps of the engine (all 32 lanes)
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int NACCS, int NOPS>
static void run(int grid, int n, int smem, long long* d, int a_sbo = 1024, int a_off = 0, int commit_every = 0, int spinners = 0) {
  if (NACCS * n > 512) return;
  const int iters = 4096;
  cudaFuncSetAttribute(rate<NACCS, NOPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate<NACCS, NOPS><<<grid, 384, smem>>>(n, 64, d, a_sbo, a_off, commit_every, spinners);      // warm-up
  rate<NACCS, NOPS><<<grid, 384, smem>>>(n, iters, d, a_sbo, a_off, commit_every, spinners);
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  printf("%d,%d,%d,%d,%d,%d,%d,%d,%.1f\n", n, NACCS, NOPS, grid, a_sbo, a_off, commit_every, spinners, (double)h / iters);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int smem = 2 * 65536 + 2048;
  printf("N,accumulators,operand_sets,grid,a_sbo,a_start_step,commit_every_8mma,spinning_warps,clk_per_mma\n");
  for (int grid : {148})
    for (int n : {32, 64, 96, 128, 160, 192, 256}) {
      run<1, 1>(grid, n, smem, d); run<1, 4>(grid, n, smem, d);
      run<2, 1>(grid, n, smem, d); run<2, 4>(grid, n, smem, d);
      run<4, 1>(grid, n, smem, d); run<4, 4>(grid, n, smem, d);
    }
  // the halo view: 8-row groups 1280 B apart (10 pixels), start row shifted by the filter tap
  for (int n : {64, 128, 256})
    for (int sbo : {1024, 1280, 2048})
      for (int off : {0, 128}) run<2, 1>(148, n, smem, d, sbo, off);
  // a commit every 8 / 16 MMAs (pipeline-stage release), and 0 / 2 / 10 other warps of the CTA blocked in mbarrier.try_wait
  for (int n : {64, 128, 256})
    for (int ce : {0, 1, 2})
      for (int sp : {0, 2, 10}) run<1, 1>(148, n, smem, d, 1024, 0, ce, sp);
  return 0;
}
