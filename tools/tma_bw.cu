// Microbenchmark: how many bytes per clock can one SM ingest through TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B boxes)?
// Decides the tiling of the MixLinear kernels.  Build: nvcc -arch=sm_100a -O3 -o tma_bw tools/tma_bw.cu -lcuda
//   ./tma_bw <box_rows> <stages> <nthreads_issuing> <src_MB> <grid> [cluster multicast width]
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
__device__ __forceinline__ void tma2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

// Each CTA streams `iters` boxes (128 B x box_rows) of its own row range through `stages` smem slots.
// `nissue` threads issue (slot s is owned by thread s % nissue); the waiter just re-arms.
__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap map, int box_rows, int stages, int nissue,
                                                         int iters, int rows_total, int kcols, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(base + (size_t)stages * box_rows * 128);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int t = threadIdx.x;
  const uint32_t bytes = box_rows * 128;
  const int row_blocks = rows_total / box_rows;
  const int kblocks = kcols / 128;
  long long t0 = clock64();
  if (t < nissue) {
    // thread t owns slots t, t+nissue, ...; it issues, waits and re-issues (steady-state ingest)
    int nslots = 0;
    for (int s = t; s < stages; s += nissue) ++nslots;
    uint32_t ph = 0;
    int item = blockIdx.x * 977 + t;   // spread CTAs over the source
    for (int it = 0; it < iters; ++it) {
      for (int s = t; s < stages; s += nissue) {
        if (it > 0) { while (!mbar_try(&bars[s], ph ^ 1)) {} }
        const int rb = item % row_blocks, kb = (item / row_blocks) % kblocks;
        item += 13;
        mbar_expect(&bars[s], bytes);
        tma2d(&map, &bars[s], base + (size_t)s * bytes, kb * 128, rb * box_rows);
      }
      ph ^= 1;
    }
    for (int s = t; s < stages; s += nissue) { while (!mbar_try(&bars[s], ph ^ 1)) {} }
    (void)nslots;
  }
  __syncthreads();
  long long t1 = clock64();
  if (t == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

int main(int argc, char** argv) {
  int box_rows = argc > 1 ? atoi(argv[1]) : 128;
  int stages = argc > 2 ? atoi(argv[2]) : 6;
  int nissue = argc > 3 ? atoi(argv[3]) : 1;
  int src_mb = argc > 4 ? atoi(argv[4]) : 64;
  int grid = argc > 5 ? atoi(argv[5]) : 148;
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
  cuuint32_t box[2] = {128, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", r); return 1; }
  unsigned long long* cyc;
  CK(cudaMalloc(&cyc, grid * sizeof(unsigned long long)));
  size_t smem = (size_t)stages * box_rows * 128 + 1024 + 256;
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 3; ++rep) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    stream_kernel<<<grid, 128, smem>>>(map, box_rows, stages, nissue, iters, (int)rows, kcols, cyc);
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
      printf("box %3d rows, %d stages, %d issuers, src %4d MB, grid %3d: %.1f B/clk/SM, %.2f TB/s aggregate (%.1f us)\n", box_rows, stages, nissue,
             src_mb, grid, bytes_per_cta / avg, bytes_per_cta * grid / (ms * 1e-3) / 1e12, ms * 1e3);
  }
  return 0;
}
