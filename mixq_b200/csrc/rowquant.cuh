// Per-token activation prologue of MixLinear: gather the fp16 outlier columns (and zero them),
// row abs-max -> x_scale, symmetric int8 (or int4-range) quantisation, and the outlier scan against
// sigma.  One warp owns one row; reductions are warp shuffles.  Used by the standalone mixlib-style
// entry points (FindRowScale / layernorm_forward_cuda_extract_outliers) and as phase A of the fused
// single-launch kernel.
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
  __half* norm_out;        // [M,K] normed output (then treated as "x" for extract/quantise)
  float eps;
  const int32_t* ind;      // [n_ind] outlier column ids
  int n_ind;
  __half* act_out;         // [M, ld_ao] gathered outlier activations
  int ld_ao;
  int8_t* q_x;             // [M,K] int8 (bit 4: values in [-7,7], one per byte)
  __half* x_scale;         // [>=M]
  int M, K, bit;
  // outlier scan (optional): col_over[c] = 1 if any |x[m,c]| > sigma; *over_flag |= 1 if any x_scale > thr
  __half sigma;
  __half thr;              // fp16(sigma / qmax), what the reference compares x_scale against
  uint8_t* col_over;       // [K] or nullptr
  uint32_t* over_flag;     // or nullptr
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

union H8 {
  uint4 u;
  __half2 h2[4];
  __half h[8];
};

// All 32 lanes of one warp call this with the same m.
__device__ __forceinline__ void quantize_row_warp(const RowQuantArgs& a, int m, int lane) {
  const int K = a.K;
  const int nvec = K >> 3;  // 8 halves per 16-byte vector; K % 8 == 0 is checked on the host
  __half* row = a.x + static_cast<size_t>(m) * K;

  if (a.norm_w != nullptr) {
    // RMSNorm: out = x * rsqrt(mean(x^2) + eps) * w, fp32 accumulation, one rounding to fp16.
    const uint4* src = reinterpret_cast<const uint4*>(row);
    float ss = 0.f;
    for (int i = lane; i < nvec; i += 32) {
      H8 v;
      v.u = __ldg(src + i);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(v.h2[j]);
        ss = fmaf(f.x, f.x, ss);
        ss = fmaf(f.y, f.y, ss);
      }
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(__fdiv_rn(ss, static_cast<float>(K)) + a.eps);
    __half* orow = a.norm_out + static_cast<size_t>(m) * K;
    const uint4* wsrc = reinterpret_cast<const uint4*>(a.norm_w);
    for (int i = lane; i < nvec; i += 32) {
      H8 v, w, o;
      v.u = __ldg(src + i);
      w.u = __ldg(wsrc + i);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        o.h[j] = __float2half_rn(__fmul_rn(__fmul_rn(__half2float(v.h[j]), rstd), __half2float(w.h[j])));
      reinterpret_cast<uint4*>(orow)[i] = o.u;
    }
    row = orow;
    __syncwarp();
  }
  if (a.q_x == nullptr) return;  // plain RMSNorm

  // 1. gather outlier columns, zero them in place
  for (int j = lane; j < a.n_ind; j += 32) {
    const int c = a.ind[j];
    a.act_out[static_cast<size_t>(m) * a.ld_ao + j] = row[c];
    row[c] = __float2half_rn(0.f);
  }
  __syncwarp();

  // 2. row abs-max of what is left
  const uint4* src = reinterpret_cast<const uint4*>(row);
  float amax = 0.f;
  for (int i = lane; i < nvec; i += 32) {
    H8 v;
    v.u = src[i];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(__habs2(v.h2[j]));
      amax = fmaxf(amax, fmaxf(f.x, f.y));
    }
  }
  amax = warp_max(amax);
  const float qmax = (a.bit == 4) ? 7.f : 127.f;
  const __half xs_h = __float2half_rn(__fdiv_rn(amax, qmax));
  const float xs = __half2float(xs_h);
  if (lane == 0) {
    a.x_scale[m] = xs_h;
    if (a.over_flag != nullptr && __hgt(xs_h, a.thr)) atomicOr(a.over_flag, 1u);
  }
  const float sigma = __half2float(a.sigma);
  const bool scan = (a.col_over != nullptr) && (amax > sigma);

  // 3. quantise: q = rint(x / x_scale), IEEE division, clamp to the symmetric range
  uint2* dst = reinterpret_cast<uint2*>(a.q_x + static_cast<size_t>(m) * K);
  for (int i = lane; i < nvec; i += 32) {
    H8 v;
    v.u = src[i];
    uint32_t packed[2] = {0u, 0u};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float f = __half2float(v.h[j]);
      float q = (xs > 0.f) ? rintf(__fdiv_rn(f, xs)) : 0.f;
      q = fminf(fmaxf(q, -qmax), qmax);
      packed[j >> 2] |= (static_cast<uint32_t>(static_cast<int>(q)) & 0xffu) << ((j & 3) * 8);
      if (scan && fabsf(f) > sigma) a.col_over[i * 8 + j] = 1;
    }
    dst[i] = make_uint2(packed[0], packed[1]);
  }
}

}  // namespace mixq
