// Per-token activation prologue of MixLinear: [RMSNorm ->] gather the fp16 outlier columns (and zero
// them), row abs-max -> x_scale, symmetric int8 (or int4-range) quantisation, and the outlier scan
// against sigma.
//
// HBM-bound byte work, laid out for latency: a GROUP of G warps owns one row.  The row is brought into
// shared memory by ONE bulk-async copy (cp.async.bulk, mbarrier complete_tx) — x is read from HBM exactly
// once, no register staging — and every pass over it is a small ROLLED loop of 16-byte LDS: this code runs
// once per launch, so straight-line unrolled code would be paced by cold instruction fetch (measured:
// stall_no_inst dominated the register-resident version).  Reductions are warp shuffles plus one named
// barrier per group; q_x is written exactly once with 8-byte coalesced stores.
// Used by the standalone mixlib-style entry points (FindRowScale / layernorm_forward_cuda[_extract_outliers])
// and as phase A of the fused single-launch kernel (row buffers = not-yet-used pipeline stages).
//
// Reference semantics restated (file:line under /root/reference):
//   mixquant/modules/linear.py:187-193  ExtractOutliersAndSetToZeros(ind, x) then FindRowScale(x, x_scale, M, K, bit)
//   mixquant/modules/linear.py:201      x_scale * (2^(bit-1)-1) == row absmax  (threshold identity)
//   mixquant/modules/linear.py:157-161  FindOutliers: columns with any |x| > sigma
//   mixquant/modules/fused/norm.py:24-33 RMSNorm + extract + quantise in one kernel
#pragma once
#include "ptx.cuh"

namespace mixq {

struct RowQuantArgs {
  __half* x;               // [M,K] fp16 row-major; outlier columns are zeroed IN PLACE (reference behaviour)
  const __half* norm_w;    // optional RMSNorm weight [K]; when set, x is the un-normed input (read-only)
  __half* norm_out;        // optional [M,K] normed output (outlier columns zeroed), may be nullptr
  float eps;
  const int32_t* ind;      // [n_ind] outlier column ids
  int n_ind;
  __half* act_out;         // [M, ld_ao] gathered outlier activations
  int ld_ao;
  int8_t* q_x;             // [M,K] int8 (bit 4: values in [-7,7], one per byte); nullptr = RMSNorm only
  __half* x_scale;         // [>=M]
  int M, K, bit;
  // outlier scan (optional): col_over[c] = 1 if any |x[m,c]| > sigma; *over_flag |= 1 if any x_scale > thr
  __half sigma;
  __half thr;              // fp16(sigma / qmax), what the reference compares x_scale against
  uint8_t* col_over;       // [K] or nullptr
  uint32_t* over_flag;     // or nullptr
  unsigned long long* trace;  // tuning aid (mixq_set_trace_buffer) or nullptr
  // work split, filled by the host (pick_row_groups)
  int group_warps;         // G: warps per row
  int ngroups;             // rows in flight per CTA; ngroups * G <= warps per CTA, ngroups * K * 2 bytes of row buffer
};

constexpr int kRowQuantMaxK = 32768;     // one row <= 64 KB of shared memory
constexpr int kRowQuantMaxGroups = 12;

struct RowQuantSmem {
  float slots[2][2][16];                 // [row parity][sum | max][warp in CTA]
  uint64_t bars[kRowQuantMaxGroups];     // one mbarrier per group: the row's bulk copy lands on it
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// global -> shared bulk copy of `bytes` (multiple of 16) completing on an mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

union H8 {
  uint4 u;
  __half2 h2[4];
  __half h[8];
};

// q = clamp(rint(f / xs)) for 8 halves, bit-identical to IEEE fp32 division followed by rint (half-even), without a
// division or a conversion-pipe instruction: t = f * r with r = fl(1/xs) (|t - f/xs| <= 127 * 2^-22 < 3e-5), rounded by
// the 1.5*2^23 magic add (FADD rounds half-even; the low byte of the sum's bit pattern is q in two's complement).
// f and xs are fp16, so f/xs = a/(b*2^s) with 11-bit a, b: a quotient that is not EXACTLY k + 1/2 is at least
// 1/8188 > 1.2e-4 away from it.  Hence |t - rint(t)| > 0.49995 can only mean an exact tie whose product landed a hair
// to either side: the answer is then the even neighbour, which is also what rint(fl(f/xs)) gives (k + 1/2 is exact in
// fp32).  Everything else is far enough from a tie for t and f/xs to round alike.
__device__ __forceinline__ uint2 quant8(uint4 vu, float r, float qmax) {
  H8 v;
  v.u = vu;
  constexpr float kMagic = 12582912.0f;   // 0x4B400000
  uint32_t qi[8];
  float dd[8];
  float worst = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float f = __half2float(v.h[j]);
    const float t = fminf(fmaxf(__fmul_rn(f, r), -qmax), qmax);
    const float u = __fadd_rn(t, kMagic);
    dd[j] = __fadd_rn(t, -__fadd_rn(u, -kMagic));
    worst = fmaxf(worst, fabsf(dd[j]));
    qi[j] = __float_as_uint(u);
  }
  if (worst > 0.49995f) {   // some element is an exact tie: take the even neighbour (rows whose scale has few mantissa bits)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool tie = fabsf(dd[j]) > 0.49995f;
      const uint32_t odd = qi[j] & 1u;
      const uint32_t step = dd[j] > 0.f ? 1u : 0xffffffffu;   // towards the other neighbour
      qi[j] += (tie && odd) ? step : 0u;
    }
  }
  const uint32_t a = __byte_perm(__byte_perm(qi[0], qi[1], 0x0040), __byte_perm(qi[2], qi[3], 0x0040), 0x5410);
  const uint32_t b = __byte_perm(__byte_perm(qi[4], qi[5], 0x0040), __byte_perm(qi[6], qi[7], 0x0040), 0x5410);
  return make_uint2(a, b);
}
__device__ __forceinline__ void scan8(uint4 vu, float sigma, uint8_t* dst) {
  H8 v;
  v.u = vu;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (fabsf(__half2float(v.h[j])) > sigma) dst[j] = 1;
}
__device__ __forceinline__ float sumsq8(uint4 vu, float ss) {
  H8 v;
  v.u = vu;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(v.h2[j]);
    ss = fmaf(f.x, f.x, ss);
    ss = fmaf(f.y, f.y, ss);
  }
  return ss;
}
__device__ __forceinline__ uint4 norm8(uint4 vu, uint4 wu, float rstd) {
  H8 v, w;
  v.u = vu;
  w.u = wu;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    v.h[j] = __float2half_rn(__fmul_rn(__fmul_rn(__half2float(v.h[j]), rstd), __half2float(w.h[j])));
  return v.u;
}
__device__ __forceinline__ __half2 absmax8(uint4 vu, __half2 m) {
  H8 v;
  v.u = vu;
#pragma unroll
  for (int j = 0; j < 4; ++j) m = __hmax2(m, __habs2(v.h2[j]));
  return m;
}

// Reduce `v` over the G consecutive warps of a group (first warp of the group = warp0).
template <bool IS_MAX>
__device__ __forceinline__ float group_reduce(float v, int G, int warp0, int warp, int lane, float* slots, int bar_id) {
  v = IS_MAX ? warp_max(v) : warp_sum(v);
  if (G == 1) return v;
  if (lane == 0) slots[warp] = v;
  named_bar_sync(bar_id, G * 32);
  float r = slots[warp0];
  for (int w = 1; w < G; ++w) r = IS_MAX ? fmaxf(r, slots[warp0 + w]) : r + slots[warp0 + w];
  return r;
}

// One row, all 32*G threads of its group.  row_s = the row in shared memory (fp16 [K]).
__device__ __forceinline__ void process_row(const RowQuantArgs& a, int m, int group, int gl, int iter, RowQuantSmem* sm,
                                            __half* row_s) {
  const int G = a.group_warps;
  const int gsize = G * 32;
  const int lane = gl & 31;
  const int warp = threadIdx.x >> 5;
  const int warp0 = group * G;
  const int K = a.K;
  const int nvec = K >> 3;  // K % 8 == 0 is checked on the host
  const int bar_id = 1 + group;
  float* s_sum = sm->slots[iter & 1][0];
  float* s_max = sm->slots[iter & 1][1];
  const uint32_t row_sa = smem_u32(row_s);          // the row in shared memory, addressed through ld/st.shared
  __half* xrow = a.x + static_cast<size_t>(m) * K;
  const bool norm = a.norm_w != nullptr;

  // RMSNorm statistics: fp32 accumulation over the raw row
  float rstd = 1.f;
  if (norm) {
    float ss = 0.f;
#pragma unroll 2
    for (int i = gl; i < nvec; i += gsize) ss = sumsq8(lds128(row_sa + i * 16), ss);
    ss = group_reduce<false>(ss, G, warp0, warp, lane, s_sum, bar_id);
    rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fdiv_rn(ss, static_cast<float>(K)) + a.eps));
  }
  if (a.q_x == nullptr) {  // plain RMSNorm (norm.py:20-21)
    const uint4* w4 = reinterpret_cast<const uint4*>(a.norm_w);
    uint4* o4 = reinterpret_cast<uint4*>(a.norm_out + static_cast<size_t>(m) * K);
#pragma unroll 2
    for (int i = gl; i < nvec; i += gsize) o4[i] = norm8(lds128(row_sa + i * 16), __ldg(w4 + i), rstd);
    return;
  }

  // gather the outlier columns (normed value = the identical arithmetic on the raw element) and zero them in the
  // shared-memory row — and, without a norm, in the caller's tensor, as the reference does (linear.py:189)
  if (a.n_ind > 0) {
    for (int j = gl; j < a.n_ind; j += gsize) {
      const int c = a.ind[j];
      __half val = __ushort_as_half(lds16(row_sa + c * 2));
      if (norm) val = __float2half_rn(__fmul_rn(__fmul_rn(__half2float(val), rstd), __half2float(a.norm_w[c])));
      else if (a.x != nullptr) xrow[c] = __float2half_rn(0.f);   // (a producer-fused caller may keep no fp16 copy at all)
      a.act_out[static_cast<size_t>(m) * a.ld_ao + j] = val;
      sts16(row_sa + c * 2, 0);
    }
    if (G == 1) __syncwarp();
    else named_bar_sync(bar_id, gsize);
  }

  // [normalise in place ->] row abs-max of what is left -> x_scale
  __half2 am2 = __float2half2_rn(0.f);
  if (norm) {
    const uint4* w4 = reinterpret_cast<const uint4*>(a.norm_w);
    uint4* o4 = a.norm_out ? reinterpret_cast<uint4*>(a.norm_out + static_cast<size_t>(m) * K) : nullptr;
#pragma unroll 2
    for (int i = gl; i < nvec; i += gsize) {
      const uint4 u = norm8(lds128(row_sa + i * 16), __ldg(w4 + i), rstd);
      sts128(row_sa + i * 16, u);   // each thread re-reads only its own vectors below
      if (o4) o4[i] = u;
      am2 = absmax8(u, am2);
    }
  } else {
#pragma unroll 4
    for (int i = gl; i < nvec; i += gsize) am2 = absmax8(lds128(row_sa + i * 16), am2);
  }
  float amax = fmaxf(__low2float(am2), __high2float(am2));
  amax = group_reduce<true>(amax, G, warp0, warp, lane, s_max, bar_id);
  if (a.trace && threadIdx.x == 0 && iter == 0) a.trace[blockIdx.x * 8 + 7] = globaltimer_ns();
  const float qmax = (a.bit == 4) ? 7.f : 127.f;
  const __half xs_h = __float2half_rn(__fdiv_rn(amax, qmax));
  const float xs = __half2float(xs_h);
  if (gl == 0) {
    a.x_scale[m] = xs_h;
    if (a.over_flag != nullptr && __hgt(xs_h, a.thr)) atomicOr(a.over_flag, 1u);
  }
  const float sigma = __half2float(a.sigma);
  const bool scan = (a.col_over != nullptr) && (amax > sigma);

  // quantise (see quant8)
  const float r = (xs > 0.f) ? __fdiv_rn(1.0f, xs) : 0.f;
  uint2* dst = reinterpret_cast<uint2*>(a.q_x + static_cast<size_t>(m) * K);
#pragma unroll 2
  for (int i = gl; i < nvec; i += gsize) {
    const uint4 u = lds128(row_sa + i * 16);
    dst[i] = quant8(u, r, qmax);
    if (scan) scan8(u, sigma, a.col_over + i * 8);
  }
}

// Row loop of one CTA: rows are dealt round-robin over (CTA, group) so that every CTA gets ceil(M / gridDim) rows
// at most.  rowbuf: ngroups * K * 2 bytes of 16-byte aligned shared memory.
// rowquant_init: threads [0, ngroups) initialise the per-group mbarriers; a __syncthreads (the caller's) must follow.
// rowquant_begin: every thread, after that barrier (and after pdl_wait: it reads x); starts each group's first row copy.
__device__ __forceinline__ void rowquant_init(const RowQuantArgs& a, RowQuantSmem* sm) {
  if (threadIdx.x < a.ngroups) mbar_init(&sm->bars[threadIdx.x], 1);
  fence_mbar_init();
}
__device__ __forceinline__ void rowquant_begin(const RowQuantArgs& a, RowQuantSmem* sm, uint8_t* rowbuf) {
  const int G = a.group_warps;
  const int group = (threadIdx.x >> 5) / G;
  const int gl = threadIdx.x - group * G * 32;
  const uint32_t row_bytes = static_cast<uint32_t>(a.K) * 2u;
  const int m = blockIdx.x + gridDim.x * group;
  if (group < a.ngroups && m < a.M && gl == 0) {
    mbar_arrive_expect_tx(&sm->bars[group], row_bytes);
    bulk_load(rowbuf + static_cast<size_t>(group) * row_bytes, a.x + static_cast<size_t>(m) * a.K, row_bytes, &sm->bars[group]);
  }
}
// rowquant_run: every thread of the CTA after rowquant_begin.
__device__ __forceinline__ void rowquant_run(const RowQuantArgs& a, RowQuantSmem* sm, uint8_t* rowbuf) {
  const int G = a.group_warps;
  const int ngroups = a.ngroups;
  const int group = (threadIdx.x >> 5) / G;
  const int gl = threadIdx.x - group * G * 32;
  const uint32_t row_bytes = static_cast<uint32_t>(a.K) * 2u;
  if (a.trace && threadIdx.x == 0) a.trace[blockIdx.x * 8 + 6] = globaltimer_ns();
  if (group >= ngroups) return;   // warps beyond the last whole group sit the prologue out
  __half* row_s = reinterpret_cast<__half*>(rowbuf + static_cast<size_t>(group) * row_bytes);
  uint64_t* bar = &sm->bars[group];
  const int stride = gridDim.x * ngroups;
  int m = blockIdx.x + gridDim.x * group;
  for (int iter = 0; m < a.M; ++iter) {
    mbar_wait(bar, iter & 1, 7, group);
    process_row(a, m, group, gl, iter, sm, row_s);
    m += stride;
    if (m < a.M) {
      // every thread is done reading row_s; generic-proxy writes to it must be ordered before the async-proxy refill
      fence_proxy_async_smem();
      if (G == 1) __syncwarp();
      else named_bar_sync(1 + group, G * 32);
      if (gl == 0) {
        mbar_arrive_expect_tx(bar, row_bytes);
        bulk_load(row_s, a.x + static_cast<size_t>(m) * a.K, row_bytes, bar);
      }
    }
  }
}
__device__ __forceinline__ void rowquant_cta(const RowQuantArgs& a, RowQuantSmem* sm, uint8_t* rowbuf) {
  rowquant_init(a, sm);
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();
  rowquant_begin(a, sm, rowbuf);
  rowquant_run(a, sm, rowbuf);
}

}  // namespace mixq
