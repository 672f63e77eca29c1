// tc_probe.cu -- hardware probe for UMMA shared-memory descriptor semantics (SWIZZLE_128B, K-major) used to design the
// halo-tile implicit-GEMM kernel: can the A operand be a SHIFTED VIEW (start address at an arbitrary 128 B row, 8-row
// groups SBO bytes apart with SBO not a multiple of 1024) of a pixel array written with the absolute-address XOR swizzle?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_probe tools/tc_probe.cu && tools/tc_probe
//
// Pixel array P[256][64 bf16] (128 B rows) at a 1024 B aligned base; 16 B chunk c of row p is stored at chunk
// c ^ ((row_address >> 7) & 7).  B = selector matrix so that D[m][n] = A[m][kbase + n].  Experiment infrastructure.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(int start_row, int sbo_bytes, int base_offset, int fill_mode, int kk, float* out /*[128][16]*/) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (sbase - raw);
  unsigned char* sA = sm;                 // 400 pixel rows x 128 B = 51200
  unsigned char* sB = sm + 51200 + 1024 - (51200 % 1024);   // 1024 aligned
  uint64_t* bar = (uint64_t*)(sB + 4096);
  uint32_t* slot = (uint32_t*)(bar + 1);
  const int tid = threadIdx.x;
  // fill A
  for (int i = tid; i < 400 * 64; i += blockDim.x) {
    const int p = i / 64, k = i % 64;
    const float v = fill_mode == 0 ? (float)(p % 256) : (float)(k + 1);
    const uint32_t row_addr = sbase + p * 128;
    const int chunk = (k / 8) ^ ((row_addr >> 7) & 7);
    *reinterpret_cast<__nv_bfloat16*>(sA + p * 128 + chunk * 16 + (k % 8) * 2) = __float2bfloat16(v);
  }
  // fill B: [16 rows n][64 k], B[n][k] = (k == kk*16 + n) for fill_mode 1;  for fill_mode 0: B[n][k] = (k == 0) (every column reads A[m][0])
  for (int i = tid; i < 16 * 64; i += blockDim.x) {
    const int n = i / 64, k = i % 64;
    const float v = fill_mode == 0 ? (k == kk * 16 ? 1.f : 0.f) : (k == kk * 16 + n ? 1.f : 0.f);
    const uint32_t row_addr = smem_u32(sB) + n * 128;
    const int chunk = (k / 8) ^ ((row_addr >> 7) & 7);
    *reinterpret_cast<__nv_bfloat16*>(sB + n * 128 + chunk * 16 + (k % 8) * 2) = __float2bfloat16(v);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(slot);
  if (tid == 0) {
    const uint32_t a_addr = sbase + start_row * 128 + kk * 32;
    const uint32_t b_addr = smem_u32(sB) + kk * 32;
    auto desc = [](uint32_t addr, uint32_t sbo, uint32_t bo) {
      return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) |
             ((uint64_t)(bo & 7) << 49) | ((uint64_t)2 << 61);   // layout 2 = SWIZZLE_128B
    };
    const uint64_t da = desc(a_addr, sbo_bytes, base_offset), db = desc(b_addr, 1024, 0);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
  }
  // wait
  {
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
      if (clock64() - t0 > 2000000000LL) __trap();
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (tid < 128) {
    const int warp = tid >> 5;
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int n = 0; n < 16; ++n) out[tid * 16 + n] = __uint_as_float(v[n]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(32) : "memory");
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 16 * 4);
  const int smem = 51200 + 2048 + 4096 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> h(128 * 16);
  struct V { int start, sbo; };
  const V vs[] = {{0, 1024}, {3, 1024}, {8, 1024}, {0, 1280}, {3, 1280}, {11, 1280}, {0, 2048}, {3, 2048}, {17, 2048}, {21, 2048}, {0, 1152}, {5, 1152}};
  for (const V& v : vs) {
    const int pitch = v.sbo / 128;
    for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
      const int bo = bo_mode ? (v.start & 7) : 0;
      if (bo_mode && bo == 0) continue;
      // test 1: which pixel row does operand row m read?
      probe<<<1, 128, smem>>>(v.start, v.sbo, bo, 0, 0, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("start %d sbo %d bo %d: CUDA error %s\n", v.start, v.sbo, bo, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
      int bad_rows = 0, first_bad = -1; float got_first = 0.f;
      for (int m = 0; m < 128; ++m) {
        const int expect = (v.start + (m / 8) * pitch + (m % 8)) % 256;
        if ((int)h[m * 16] != expect) { if (first_bad < 0) { first_bad = m; got_first = h[m * 16]; } ++bad_rows; }
      }
      // test 2: are the 64 k values of each row un-swizzled correctly (4 K=16 steps)?
      int bad_k = 0;
      for (int kk = 0; kk < 4; ++kk) {
        probe<<<1, 128, smem>>>(v.start, v.sbo, bo, 1, kk, d);
        cudaDeviceSynchronize();
        cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 16; ++n)
            if ((int)h[m * 16 + n] != kk * 16 + n + 1) ++bad_k;
      }
      printf("start_row %2d  SBO %4d  base_offset %d : wrong rows %3d (first %3d, got %g)  wrong k-values %5d  %s\n", v.start, v.sbo, bo, bad_rows, first_bad,
             got_first, bad_k, (bad_rows == 0 && bad_k == 0) ? "OK" : "MISMATCH");
    }
  }
  return 0;
}
