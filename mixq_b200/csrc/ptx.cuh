// Thin inline-PTX wrappers for the sm_100a features the MixLinear kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and the
// proxy fences between them.  Nothing here is generic; it is exactly what
// mixq_gemm.cu needs.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace mixq {

// A spin that never ends would hang the whole device; every wait loop in this
// library gives up after this many nanoseconds and traps so the host sees an error.
#ifndef MIXQ_SPIN_TIMEOUT_NS
#define MIXQ_SPIN_TIMEOUT_NS 4000000000ull
#endif

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

// One lane of a fully converged warp.  The single-thread roles (TMA issue, tcgen05.mma issue) run their loops
// WARP-UNIFORMLY and predicate only the issuing instruction with this: inside an `if (lane == 0)` region the compiler
// cannot keep descriptors / addresses in uniform registers and wraps every UTCIMMA / UTMALDG in an ELECT + R2UR.BROADCAST
// loop (~120 issue cycles per MMA measured: the whole GEMM was paced by its issuing thread).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- programmatic dependent launch
// Every kernel of this library is launched with cudaLaunchAttributeProgrammaticStreamSerialization (mixq_api.cu), so its
// CTAs may start while the previous kernel of the stream is still draining.  pdl_launch_dependents() lets the NEXT kernel's
// CTAs queue up behind ours as early as possible; pdl_wait() returns once every earlier kernel has completed and its
// memory is visible — nothing before it may touch a tensor another kernel writes or reads (constants such as the
// quantised weights may be fetched earlier: that is the point).  Both are no-ops for a plain launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
static __device__ __noinline__ void spin_timeout_trap(int what, int a, int b) {
  printf("mixq: device wait timed out (what=%d a=%d b=%d block=%d thread=%d)\n", what, a, b, blockIdx.x,
         threadIdx.x);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int what = 0, int tag = 0) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xff) == 0 && globaltimer_ns() - t0 > MIXQ_SPIN_TIMEOUT_NS) spin_timeout_trap(what, tag, parity);
  }
}

// Long waits by whole warps (epilogue warps waiting for a tile): one lane polls, with back-off, so that the spinning
// does not compete with the tensor core / TMA for the shared-memory and barrier pipes.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int what = 0, int tag = 0) {
  if ((threadIdx.x & 31) == 0) {
    if (!mbar_try_wait(bar, parity)) {
      uint64_t t0 = globaltimer_ns();
      uint32_t spins = 0;
      while (!mbar_try_wait(bar, parity)) {
        __nanosleep(100);
        if ((++spins & 0xff) == 0 && globaltimer_ns() - t0 > MIXQ_SPIN_TIMEOUT_NS) spin_timeout_trap(what, tag, parity);
      }
    }
  }
  __syncwarp();
}

// Short waits by whole warps inside a pipeline: lane 0 polls without back-off.
__device__ __forceinline__ void mbar_wait_lane0(uint64_t* bar, uint32_t parity, int what = 0, int tag = 0) {
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity, what, tag);
  __syncwarp();
}

// ---------------------------------------------------------------- proxies
// generic-proxy writes (st.global / st.shared) that a later async-proxy op (TMA, tcgen05.mma)
// must observe, and the reverse.
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// L2 eviction-priority policies for the cache_hint operand (same encodings CUTLASS uses).
static constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
static constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
static constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* smem_dst, int c0, int c1,
                                            uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "l"(hint)
      : "memory");
}
// Pull one cache line into L2 (LSU path; a TMA prefetch would cost a full TMA issue slot — measured slower).
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p) : "memory");
}
// Pull one box into L2 without landing it anywhere (used once per launch for the weight boxes just beyond the pipeline depth:
// HBM latency is longer than the 4-stage ring covers at start-up; in the steady state the prefetch costs a TMA issue slot per
// box and was measured 15 % slower, see mixq_gemm2.cu).
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1),
               "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// plain (non-tensor) bulk copy shared -> global, tracked by the bulk async-group of the issuing thread
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(gdst)), "r"(smem_src),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32.  One thread issues.
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// fp16 x fp16 -> fp32.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All tcgen05 async ops issued so far by this thread arrive on `bar` when they retire.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp gets TMEM lane (base_lane + t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// 32 lanes x 16 consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// The reverse: 16 consecutive 32-bit columns of registers -> TMEM (split-K: partial sums folded into the accumulator).
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor for a K-major operand tile stored as rows of 128 bytes,
// 128B-swizzled in groups of 8 rows (exactly what a SWIZZLE_128B TMA box of 128 B x rows writes).
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4 (unused for SW128 K-major)
//   bits [32,46) stride byte offset >> 4   (8 rows * 128 B = 1024 B between row groups)
//   bits [46,48) version = 1 (Blackwell)   bits [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;             // LBO: ignored for swizzled K-major, keep 1
  d |= static_cast<uint64_t>(1024 >> 4) << 32;     // SBO
  d |= static_cast<uint64_t>(1) << 46;             // descriptor version
  d |= static_cast<uint64_t>(2) << 61;             // SWIZZLE_128B
  return d;
}

// Instruction descriptor (upper 32 bits of the "idesc" operand), dense, K-major A and B.
//   [4,6) D format: 1 = F32, 2 = S32     [7,10) A format   [10,13) B format
//   (i8 kind: 0 = u8, 1 = s8; f16 kind: 0 = f16, 1 = bf16)
//   [15] A major (0 = K)  [16] B major (0 = K)  [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_i8(int m, int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// ---------------------------------------------------------------- shared memory through 32-bit shared-window addresses
// (a pointer derived from the aligned dynamic-smem base has lost its address space: the compiler emits generic LD.E / ST.E,
// a slower path than LDS / STS)
__device__ __forceinline__ uint4 lds128(uint32_t sa) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sa) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t sa, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sa), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint16_t lds16(uint32_t sa) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(sa) : "memory");
  return v;
}
__device__ __forceinline__ void sts16(uint32_t sa, uint16_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(sa), "h"(v) : "memory"); }

// ---------------------------------------------------------------- packed-nibble sign extension
// Four two's-complement nibbles (the low / high nibbles of the four bytes of w) -> four int8 lanes: n | (n & 8 ? 0xF0 : 0).
// ((w >> 3) & 0x01010101) * 0xF0 puts 0xF0 in the lanes whose nibble has its sign bit set (one bit per lane: no carries), so
// each half costs two logic ops and one IMAD — the SIMD-in-a-word video instructions (__vsub4) are emulated on this
// architecture and made the unpack the pace-setter of the W4 mainloop.
__device__ __forceinline__ uint32_t nib_lo_s8x4(uint32_t w) {
  return ((w >> 3) & 0x01010101u) * 0xF0u + (w & 0x0F0F0F0Fu);
}
__device__ __forceinline__ uint32_t nib_hi_s8x4(uint32_t w) {
  return ((w >> 7) & 0x01010101u) * 0xF0u + ((w >> 4) & 0x0F0F0F0Fu);
}

// ---------------------------------------------------------------- misc
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Sense-reversing grid barrier on one 32-bit word that never needs resetting: block 0 adds
// 0x80000000 - (grid-1), every other block adds 1, so each use flips the top bit exactly once.
// Callers must guarantee co-residency of the whole grid (cooperative launch).
__device__ __forceinline__ void grid_barrier(uint32_t* word, unsigned long long* dbg = nullptr) {
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t add = (blockIdx.x == 0) ? (0x80000000u - (gridDim.x - 1)) : 1u;
    if (dbg) dbg[0] = globaltimer_ns();
    __threadfence();
    if (dbg) dbg[1] = globaltimer_ns();
    uint32_t old = atomicAdd(word, add);
    if (dbg) dbg[2] = globaltimer_ns();
    uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (((old ^ ld_acquire_u32(word)) & 0x80000000u) == 0) {
      if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > MIXQ_SPIN_TIMEOUT_NS) spin_timeout_trap(9, 0, 0);
    }
    if (dbg) dbg[3] = globaltimer_ns();
    __threadfence();
  }
  __syncthreads();
}

}  // namespace mixq

// ================================================================ 2-CTA (cta_group::2) variants
namespace mixq {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// Arrive on a barrier of another CTA of the cluster.  Used to hand TMEM columns back to the MMA warp: the tcgen05.ld's
// are ordered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync, so CTA-scope release is all that is needed — a
// cluster-scope release would also wait for the warp's outstanding global stores of y (measured: ~1 us per arrive).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cta.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Cluster-scope release: everything this thread wrote (to ITS OWN shared memory, read next by the pair's tensor cores) is
// ordered before the arrival.  Used by the nibble-unpack warps, which have no global stores in flight.
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// TMA load into THIS CTA's shared memory, completing on the barrier at `bar_cluster_addr` (the pair leader's).
__device__ __forceinline__ void tma_load_2d_2cta(const CUtensorMap* m, uint32_t bar_cluster_addr, void* smem_dst, int c0,
                                                 int c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
// 3-D box {c0, c1, c2} (the [bytes in a 128 B k-atom, row, k-atom] view of a row-major int8 matrix): one op brings
// several k-atoms of the same rows, landing as [atom][row][128 B].
__device__ __forceinline__ void tma_load_3d_2cta(const CUtensorMap* m, uint32_t bar_cluster_addr, void* smem_dst, int c0, int c1,
                                                 int c2, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
      : "memory");
}
// D[tmem, 256 x N over the CTA pair] (+)= A * B^T; issued by one thread of the pair's leader CTA.
__device__ __forceinline__ void umma_i8_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The four K = 32 steps of one 128-byte k-atom in ONE asm block, for one or two accumulator column chunks.
// The MMA warp runs one such block per k-atom and nothing else in its loop: with the descriptors built in C++ (64-bit adds
// in the uniform datapath, a predicate per MMA) the loop was ~110 mostly dependent instructions per stage and paced narrow
// tiles at ~680 clocks per stage even with every load and every MMA removed (MIXQ_DEBUG_ABLATE, profiles/).
//   lo_a / lo_b: low words of the SWIZZLE_128B K-major descriptors of the atom (start address >> 4 | LBO); the high word is
//   the constant kDescHi; +2 in the low word = +32 bytes = the next K step.  acc = 0: the first step overwrites D.
static constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo_sw128(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma_i8_2cta_atom(uint32_t d1, uint32_t lo_a, uint32_t lo_b, uint32_t idesc1, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 a, b;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 q, 0, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], da, db, %3, p;\n\t"
      "add.u32 a, %1, 2;\n\tadd.u32 b, %2, 2;\n\tmov.b64 da, {a, %5};\n\tmov.b64 db, {b, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], da, db, %3, q;\n\t"
      "add.u32 a, %1, 4;\n\tadd.u32 b, %2, 4;\n\tmov.b64 da, {a, %5};\n\tmov.b64 db, {b, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], da, db, %3, q;\n\t"
      "add.u32 a, %1, 6;\n\tadd.u32 b, %2, 6;\n\tmov.b64 da, {a, %5};\n\tmov.b64 db, {b, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], da, db, %3, q;\n\t}"
      :
      : "r"(d1), "r"(lo_a), "r"(lo_b), "r"(idesc1), "r"(acc), "r"(kDescHi)
      : "memory");
}
// Two column chunks per K step: D1 (+)= A * B[rows 0..], D2 (+)= A * B[rows off2..]  (lo_b2 = lo_b + row offset >> 4).
__device__ __forceinline__ void umma_i8_2cta_atom2(uint32_t d1, uint32_t d2, uint32_t lo_a, uint32_t lo_b, uint32_t lo_b2,
                                                   uint32_t idesc1, uint32_t idesc2, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db, dc;\n\t.reg .b32 a, b, c;\n\t"
      "setp.ne.b32 p, %7, 0;\n\t"
      "setp.eq.b32 q, 0, 0;\n\t"
      "mov.b64 da, {%2, %8};\n\tmov.b64 db, {%3, %8};\n\tmov.b64 dc, {%4, %8};\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], da, db, %5, p;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%1], da, dc, %6, p;\n\t"
      "add.u32 a, %2, 2;\n\tadd.u32 b, %3, 2;\n\tadd.u32 c, %4, 2;\n\tmov.b64 da, {a, %8};\n\tmov.b64 db, {b, %8};\n\tmov.b64 dc, {c, %8};\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], da, db, %5, q;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%1], da, dc, %6, q;\n\t"
      "add.u32 a, %2, 4;\n\tadd.u32 b, %3, 4;\n\tadd.u32 c, %4, 4;\n\tmov.b64 da, {a, %8};\n\tmov.b64 db, {b, %8};\n\tmov.b64 dc, {c, %8};\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], da, db, %5, q;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%1], da, dc, %6, q;\n\t"
      "add.u32 a, %2, 6;\n\tadd.u32 b, %3, 6;\n\tadd.u32 c, %4, 6;\n\tmov.b64 da, {a, %8};\n\tmov.b64 db, {b, %8};\n\tmov.b64 dc, {c, %8};\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], da, db, %5, q;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%1], da, dc, %6, q;\n\t}"
      :
      : "r"(d1), "r"(d2), "r"(lo_a), "r"(lo_b), "r"(lo_b2), "r"(idesc1), "r"(idesc2), "r"(acc), "r"(kDescHi)
      : "memory");
}
// All tcgen05 ops issued so far by this thread arrive (once) on the barrier at this offset in every CTA of `mask`.
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_i8_rt(int m, int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc_f16_rt(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

}  // namespace mixq
