// Microbenchmark: cycles per tcgen05.mma kind::i8 K-block (4 MMAs of K=32 over one 128-byte SWIZZLE_128B k-block) with the
// operands resident in shared memory (no TMA traffic), for cta_group::1 (128 x N) and cta_group::2 (256 x N).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bw tools/mma_bw.cu ; ./tools/mma_bw
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
  return ok;
}

template <int CG>
__global__ void __launch_bounds__(128, 1) mma_kernel(int n, int iters, int stages, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)((CG == 2 ? 256 : 128) >> 4) << 24);
  const int brows = CG == 2 ? n / 2 : n;
  const int stage_bytes = 16384 + brows * 128;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0 && rank == 0) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages;
      const uint64_t da = desc_sw128(smem_u32(smem + (size_t)s * stage_bytes));
      const uint64_t db = desc_sw128(smem_u32(smem + (size_t)s * stage_bytes + 16384));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (CG == 1)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tmem), "l"(da + 2 * k),
                       "l"(db + 2 * k), "r"(idesc), "r"(1)
                       : "memory");
        else
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tmem), "l"(da + 2 * k),
                       "l"(db + 2 * k), "r"(idesc), "r"(1)
                       : "memory");
      }
    }
    if (CG == 1)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                   "h"((uint16_t)1)
                   : "memory");
    while (!mbar_try(&bar, 0)) {}
    t1 = clock64();
    out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int CG>
void run(int n, int stages, int grid) {
  const int iters = 2000;
  const int brows = CG == 2 ? n / 2 : n;
  size_t smem = (size_t)stages * (16384 + brows * 128) + 1024;
  unsigned long long* out;
  CK(cudaMalloc(&out, grid * sizeof(unsigned long long)));
  CK(cudaMemset(out, 0, grid * sizeof(unsigned long long)));
  CK(cudaFuncSetAttribute(mma_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, mma_kernel<CG>, n, iters, stages, out));
  CK(cudaDeviceSynchronize());
  unsigned long long h[148];
  CK(cudaMemcpy(h, out, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  int cnt = 0;
  for (int i = 0; i < grid; ++i)
    if (h[i]) { avg += h[i]; ++cnt; }
  avg /= cnt;
  const double cyc_kb = avg / iters;
  const double macs = (CG == 2 ? 256.0 : 128.0) * n * 128;
  printf("cta_group::%d  %3d x %3d, %d smem stages, grid %3d: %.1f cycles per k-block (4 MMAs) -> %.0f MAC/clk/SM\n", CG, CG == 2 ? 256 : 128, n,
         stages, grid, cyc_kb, macs / cyc_kb / CG);
  cudaFree(out);
}

int main() {
  for (int n : {64, 128, 160, 192, 256}) run<1>(n, 4, 148);
  for (int n : {64, 128, 160, 192, 256}) run<2>(n, 4, 148);
  run<2>(256, 1, 148);
  run<2>(256, 4, 2);
  run<1>(256, 4, 1);
  return 0;
}
