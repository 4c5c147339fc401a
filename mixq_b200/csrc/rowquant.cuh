// Per-token activation prologue of MixLinear: [RMSNorm ->] gather the fp16 outlier columns (and zero
// them), row abs-max -> x_scale, symmetric int8 (or int4-range) quantisation, and the outlier scan
// against sigma.
//
// HBM-bound byte work, laid out for latency: a GROUP of G warps (G = 1, 2 or 4, chosen on the host so
// that every row of the batch is in flight at once) owns one row and keeps the whole row in registers —
// x is read exactly once with 16-byte coalesced loads, q_x is written exactly once with 8-byte
// coalesced stores.  Reductions are warp shuffles plus one named barrier per group.  Outlier columns
// are zeroed in registers through a K-bit mask in shared memory that the CTA builds once.
// Used by the standalone mixlib-style entry points (FindRowScale / layernorm_forward_cuda[_extract_outliers])
// and as phase A of the fused single-launch kernel.
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
  // work split, filled by the host (pick_row_groups)
  int group_warps;         // G: warps per row, 1 / 2 / 4
  int nv;                  // 16-byte vectors per lane: 8 / 16 / 32  (>= ceil(K/8 / (32 G)))
};

constexpr int kRowQuantMaxK = 32768;                    // 4 warps x 32 lanes x 32 vectors x 8 halves
constexpr int kRowQuantMaskBytes = kRowQuantMaxK / 8;   // one bit per column
constexpr int kRowQuantSmemBytes = kRowQuantMaskBytes + 512;  // + 2 parities x {sum,max} x 32 floats

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

union H8 {
  uint4 u;
  __half2 h2[4];
  __half h[8];
};

// Every thread of the CTA calls this once before the row loop (contains __syncthreads).
__device__ __forceinline__ void rowquant_build_mask(const RowQuantArgs& a, uint8_t* smem) {
  if (a.n_ind <= 0 || a.q_x == nullptr) return;
  uint32_t* mask = reinterpret_cast<uint32_t*>(smem);
  const int nwords = (a.K + 31) >> 5;
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) mask[i] = 0u;
  __syncthreads();
  for (int j = threadIdx.x; j < a.n_ind; j += blockDim.x) {
    const int c = a.ind[j];
    atomicOr(&mask[c >> 5], 1u << (c & 31));
  }
  __syncthreads();
}

// Reduce `v` over the G warps of a group.  slots: G floats private to (group, parity).
template <bool IS_MAX>
__device__ __forceinline__ float group_reduce(float v, int G, int warp_in_group, int lane, float* slots, int bar_id) {
  v = IS_MAX ? warp_max(v) : warp_sum(v);
  if (G == 1) return v;
  if (lane == 0) slots[warp_in_group] = v;
  named_bar_sync(bar_id, G * 32);
  float r = slots[0];
  for (int w = 1; w < G; ++w) r = IS_MAX ? fmaxf(r, slots[w]) : r + slots[w];
  return r;
}

// All 32*G threads of one group call this with the same m.  `gl` = thread index inside the group.
// `iter` = how many rows this group has already processed (selects the reduction slot parity).
template <int NV>
__device__ __forceinline__ void quantize_row_group(const RowQuantArgs& a, int m, int group, int gl, int iter,
                                                   uint8_t* smem) {
  const int G = a.group_warps;
  const int gsize = G * 32;
  const int lane = gl & 31;
  const int wig = gl >> 5;
  const int K = a.K;
  const int nvec = K >> 3;  // K % 8 == 0 is checked on the host
  const uint8_t* mask = smem;
  float* slots = reinterpret_cast<float*>(smem + kRowQuantMaskBytes) + (iter & 1) * 64 + group * 4;
  const int bar_id = 1 + group;
  __half* xrow = a.x + static_cast<size_t>(m) * K;

  // 1. the row, once
  H8 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = gl + i * gsize;
    v[i].u = (idx < nvec) ? *reinterpret_cast<const uint4*>(xrow + static_cast<size_t>(idx) * 8) : make_uint4(0, 0, 0, 0);
  }

  // 2. RMSNorm in registers: out = fp16((x * rstd) * w), fp32 accumulation
  float rstd = 1.f;
  const bool norm = a.norm_w != nullptr;
  if (norm) {
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(v[i].h2[j]);
        ss = fmaf(f.x, f.x, ss);
        ss = fmaf(f.y, f.y, ss);
      }
    }
    ss = group_reduce<false>(ss, G, wig, lane, slots, bar_id);
    rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fdiv_rn(ss, static_cast<float>(K)) + a.eps));
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = gl + i * gsize;
      if (idx < nvec) {
        H8 w;
        w.u = __ldg(reinterpret_cast<const uint4*>(a.norm_w) + idx);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          v[i].h[j] = __float2half_rn(__fmul_rn(__fmul_rn(__half2float(v[i].h[j]), rstd), __half2float(w.h[j])));
      }
    }
  }
  if (a.q_x == nullptr) {  // plain RMSNorm (norm.py:20-21)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = gl + i * gsize;
      if (idx < nvec) *reinterpret_cast<uint4*>(a.norm_out + static_cast<size_t>(m) * K + static_cast<size_t>(idx) * 8) = v[i].u;
    }
    return;
  }

  // 3. gather the outlier columns (the value lives in another lane's registers: re-derive it from x with
  //    the identical arithmetic) and zero them where the caller can see them
  for (int j = gl; j < a.n_ind; j += gsize) {
    const int c = a.ind[j];
    __half val = xrow[c];
    if (norm) val = __float2half_rn(__fmul_rn(__fmul_rn(__half2float(val), rstd), __half2float(a.norm_w[c])));
    else xrow[c] = __float2half_rn(0.f);
    a.act_out[static_cast<size_t>(m) * a.ld_ao + j] = val;
  }
  if (a.n_ind > 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = gl + i * gsize;
      if (idx < nvec) {
        const uint32_t mb = mask[idx];
        if (mb != 0u) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (mb & (1u << j)) v[i].h[j] = __float2half_rn(0.f);
        }
      }
    }
  }
  if (norm && a.norm_out != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = gl + i * gsize;
      if (idx < nvec) *reinterpret_cast<uint4*>(a.norm_out + static_cast<size_t>(m) * K + static_cast<size_t>(idx) * 8) = v[i].u;
    }
  }

  // 4. row abs-max of what is left -> x_scale
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(__habs2(v[i].h2[j]));
      amax = fmaxf(amax, fmaxf(f.x, f.y));
    }
  }
  amax = group_reduce<true>(amax, G, wig, lane, slots + 32, bar_id);
  const float qmax = (a.bit == 4) ? 7.f : 127.f;
  const __half xs_h = __float2half_rn(__fdiv_rn(amax, qmax));
  const float xs = __half2float(xs_h);
  if (gl == 0) {
    a.x_scale[m] = xs_h;
    if (a.over_flag != nullptr && __hgt(xs_h, a.thr)) atomicOr(a.over_flag, 1u);
  }
  const float sigma = __half2float(a.sigma);
  const bool scan = (a.col_over != nullptr) && (amax > sigma);

  // 5. quantise: q = rint(x / x_scale), IEEE division, clamp to the symmetric range
  uint2* dst = reinterpret_cast<uint2*>(a.q_x + static_cast<size_t>(m) * K);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = gl + i * gsize;
    if (idx < nvec) {
      uint32_t packed[2] = {0u, 0u};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float f = __half2float(v[i].h[j]);
        float q = (xs > 0.f) ? rintf(__fdiv_rn(f, xs)) : 0.f;
        q = fminf(fmaxf(q, -qmax), qmax);
        packed[j >> 2] |= (static_cast<uint32_t>(static_cast<int>(q)) & 0xffu) << ((j & 3) * 8);
        if (scan && fabsf(f) > sigma) a.col_over[idx * 8 + j] = 1;
      }
      dst[idx] = make_uint2(packed[0], packed[1]);
    }
  }
}

// Row loop of one CTA: rows are dealt round-robin over (CTA, group) so that every CTA gets
// ceil(M / gridDim) rows at most.  Every thread of the CTA must call this (it contains __syncthreads).
__device__ __forceinline__ void rowquant_cta(const RowQuantArgs& a, uint8_t* smem) {
  rowquant_build_mask(a, smem);
  const int G = a.group_warps;
  const int ngroups = (blockDim.x >> 5) / G;
  const int group = (threadIdx.x >> 5) / G;
  const int gl = threadIdx.x - group * G * 32;
  if (group >= ngroups) return;   // warps that do not fill a whole group sit the prologue out
  int iter = 0;
  for (int m = blockIdx.x + gridDim.x * group; m < a.M; m += gridDim.x * ngroups, ++iter) {
    if (a.nv <= 8) quantize_row_group<8>(a, m, group, gl, iter, smem);
    else if (a.nv <= 16) quantize_row_group<16>(a, m, group, gl, iter, smem);
    else quantize_row_group<32>(a, m, group, gl, iter, smem);
  }
}

}  // namespace mixq
