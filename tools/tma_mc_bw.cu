// Microbenchmark: does TMA multicast raise the bytes per clock an SM can LAND when the CTAs of a cluster want the same box?
// Every CTA of a cluster of `cs` loads 1/cs of each 128 B x box_rows box and multicasts its slice to all cs CTAs, so each
// SM receives the full box while L2 is read once per cluster.  cs = 1 is the unicast reference (same code path).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_mc_bw tools/tma_mc_bw.cu -lcuda
//   ./tools/tma_mc_bw <box_rows> <stages> <src_MB> <cluster size 1|2|4|8>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void remote_arrive(uint64_t* b, uint32_t rank) {
  uint32_t a;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(b)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cta.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void tma2d_mc(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
                   smem_u32(dst)),
               "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tma2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

__global__ void __launch_bounds__(128, 1) mc_kernel(const __grid_constant__ CUtensorMap map, int box_rows, int stages, int iters, int cs,
                                                     int rows_total, int kcols, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(base + (size_t)stages * box_rows * 128);
  uint64_t* empty = full + stages;
  const uint32_t r = cluster_rank();
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], cs); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  const uint32_t bytes = box_rows * 128;
  const int slice = box_rows / cs;
  const int row_blocks = rows_total / box_rows, kblocks = kcols / 128;
  const int cluster_id = blockIdx.x / cs;
  const uint16_t mask = (uint16_t)((1u << cs) - 1);
  long long t0 = clock64();
  if (threadIdx.x == 0) {          // producer: my slice of every box, multicast to the whole cluster
    uint32_t ph = 0;
    int item = cluster_id * 977;
    for (int it = 0; it < iters; ++it) {
      for (int s = 0; s < stages; ++s) {
        if (it > 0) { while (!mbar_try(&empty[s], ph ^ 1)) {} }
        const int rb = item % row_blocks, kb = (item / row_blocks) % kblocks;
        item += 13;
        mbar_expect(&full[s], bytes);
        uint8_t* dst = base + (size_t)s * bytes + (size_t)r * slice * 128;
        if (cs == 1) tma2d(&map, &full[s], dst, kb * 128, rb * box_rows);
        else tma2d_mc(&map, &full[s], dst, kb * 128, rb * box_rows + r * slice, mask);
      }
      ph ^= 1;
    }
  } else if (threadIdx.x == 32) {  // consumer: box landed -> tell every producer of the cluster that slot s is free here
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      for (int s = 0; s < stages; ++s) {
        while (!mbar_try(&full[s], ph)) {}
        for (int q = 0; q < cs; ++q) remote_arrive(&empty[s], q);
      }
      ph ^= 1;
    }
  }
  __syncthreads();
  long long t1 = clock64();
  cluster_sync();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

int main(int argc, char** argv) {
  int box_rows = argc > 1 ? atoi(argv[1]) : 128;
  int stages = argc > 2 ? atoi(argv[2]) : 6;
  int src_mb = argc > 3 ? atoi(argv[3]) : 64;
  int cs = argc > 4 ? atoi(argv[4]) : 1;
  int iters = 200;
  const int kcols = 4096;
  const long long rows = (long long)src_mb * 1024 * 1024 / kcols;
  uint8_t* src;
  CK(cudaMalloc(&src, rows * kcols));
  CK(cudaMemset(src, 1, rows * kcols));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
  CUtensorMap map;
  cuuint64_t gdim[2] = {(cuuint64_t)kcols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)kcols};
  cuuint32_t box[2] = {128, (cuuint32_t)(box_rows / cs)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", r); return 1; }
  int grid = (148 / cs) * cs;
  unsigned long long* cyc;
  CK(cudaMalloc(&cyc, grid * sizeof(unsigned long long)));
  size_t smem = (size_t)stages * box_rows * 128 + 1024 + 256;
  CK(cudaFuncSetAttribute(mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (cs > 8) CK(cudaFuncSetAttribute(mc_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  // the largest grid of whole clusters that is co-resident (GPCs differ in size)
  int max_clusters = 0;
  cfg.gridDim = dim3(grid);
  CK(cudaOccupancyMaxActiveClusters(&max_clusters, mc_kernel, &cfg));
  if (max_clusters * cs < grid) grid = max_clusters * cs;
  cfg.gridDim = dim3(grid);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    CK(cudaLaunchKernelEx(&cfg, mc_kernel, map, box_rows, stages, iters, cs, (int)rows, kcols, cyc));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<unsigned long long> h(grid);
    CK(cudaMemcpy(h.data(), cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    double avg = 0;
    for (auto c : h) avg += c;
    avg /= grid;
    const double bytes_per_cta = (double)iters * stages * box_rows * 128;
    if (rep == 2)
      printf("cluster %d, box %3d rows, %d stages, src %4d MB, grid %3d: %.1f B/clk/SM landed, %.2f TB/s landed aggregate, %.2f TB/s read from L2 (%.1f us)\n",
             cs, box_rows, stages, src_mb, grid, bytes_per_cta / avg, bytes_per_cta * grid / (ms * 1e-3) / 1e12,
             bytes_per_cta * grid / cs / (ms * 1e-3) / 1e12, ms * 1e3);
  }
  return 0;
}
