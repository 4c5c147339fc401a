// Small HBM-bound kernels around the GEMM: standalone row quantisation / RMSNorm, outlier gather,
// split-path dequant, nibble unpack, weight-column gather, outlier-column compaction.
// Each restates one mixlib.* call of the reference (see include/mixq.h for file:line).
#include "mixq_kernels.cuh"

namespace mixq {

// Standalone activation prologue: groups of warps own rows staged in dynamic shared memory (rowquant.cuh).
__global__ void __launch_bounds__(256, 4) rowquant_kernel(const RowQuantArgs a) {
  extern __shared__ __align__(128) uint8_t rowbuf[];
  __shared__ RowQuantSmem sm;
  rowquant_cta(a, &sm, rowbuf);
}

__global__ void extract_outliers_kernel(const int32_t* __restrict__ ind, int n_ind, __half* x, __half* out,
                                        int ld_out, int M, int K) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(M) * n_ind;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i / n_ind);
    const int j = static_cast<int>(i - static_cast<long long>(m) * n_ind);
    const int c = ind[j];
    __half* px = x + static_cast<size_t>(m) * K + c;
    out[static_cast<size_t>(m) * ld_out + j] = *px;
    *px = __float2half_rn(0.f);
  }
}

// y = act(fp16((float(acc)*xs[m])*ws[n] + outl[m,n])), 8 columns per thread.
__global__ void dequant_i32_kernel(const int32_t* __restrict__ acc, const __half* __restrict__ x_scale,
                                   const __half* __restrict__ scale_col, const __half* __restrict__ outl,
                                   int ld_outl, __half* __restrict__ y, int M, int N, int act) {
  pdl_launch_dependents();
  pdl_wait();
  const int nv = N >> 3;
  const long long total = static_cast<long long>(M) * nv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i / nv);
    const int n = static_cast<int>(i - static_cast<long long>(m) * nv) * 8;
    const float xs = __half2float(x_scale[m]);
    const int4 a0 = *reinterpret_cast<const int4*>(acc + static_cast<size_t>(m) * N + n);
    const int4 a1 = *reinterpret_cast<const int4*>(acc + static_cast<size_t>(m) * N + n + 4);
    const int av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    H8 ws, ol, out;
    ws.u = __ldg(reinterpret_cast<const uint4*>(scale_col + n));
    if (outl != nullptr) ol.u = *reinterpret_cast<const uint4*>(outl + static_cast<size_t>(m) * ld_outl + n);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = __fmul_rn(__fmul_rn(static_cast<float>(av[j]), xs), __half2float(ws.h[j]));
      if (outl != nullptr) v = __fadd_rn(v, __half2float(ol.h[j]));
      if (act == 1) v = __fdividef(v, 1.0f + __expf(-v));
      out.h[j] = __float2half_rn(v);
    }
    *reinterpret_cast<uint4*>(y + static_cast<size_t>(m) * N + n) = out.u;
  }
}

__device__ __forceinline__ int nibble_at(const uint8_t* q_w_packed, size_t row, int K, int c) {
  const uint8_t b = q_w_packed[row * (K >> 1) + (c >> 1)];
  const int nib = (c & 1) ? (b >> 4) : (b & 0xF);
  return (nib ^ 8) - 8;  // two's-complement sign extension of 4 bits
}

__global__ void unpack_int4_cols_kernel(const uint8_t* __restrict__ q_w_packed, const int32_t* __restrict__ ind,
                                        int n_ind, __half* out, int ld_out, int N, int K) {
  const long long total = static_cast<long long>(N) * n_ind;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / n_ind);
    const int j = static_cast<int>(i - static_cast<long long>(n) * n_ind);
    out[static_cast<size_t>(n) * ld_out + j] = __int2half_rn(nibble_at(q_w_packed, n, K, ind[j]));
  }
}

// wc[n, col0+j] = fp16(q_w[n, ind[j]]) * scale_col[n]   (one fp16 multiply, as torch does on fp16 tensors)
__global__ void gather_weight_cols_kernel(const void* __restrict__ q_w, const __half* __restrict__ scale_col,
                                          const int32_t* __restrict__ ind, int n_ind, __half* wc, int ld_wc,
                                          int col0, int N, int K, int bit) {
  const long long total = static_cast<long long>(N) * n_ind;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / n_ind);
    const int j = static_cast<int>(i - static_cast<long long>(n) * n_ind);
    const int c = ind[j];
    int q;
    if (bit == 8) q = static_cast<const int8_t*>(q_w)[static_cast<size_t>(n) * K + c];
    else q = nibble_at(static_cast<const uint8_t*>(q_w), n, K, c);
    wc[static_cast<size_t>(n) * ld_wc + col0 + j] = __hmul(__int2half_rn(q), scale_col[n]);
  }
}

// Single block: ordered compaction of the K flag bytes into ascending column ids.
__global__ void __launch_bounds__(1024) compact_cols_kernel(uint8_t* col_over, int K, int32_t* ind_out,
                                                            int max_new, int32_t* n_new) {
  __shared__ int warp_tot[32];
  __shared__ int base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < K; c0 += blockDim.x) {
    const int c = c0 + threadIdx.x;
    const int f = (c < K && col_over[c] != 0) ? 1 : 0;
    if (c < K) col_over[c] = 0;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int pre = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int woff = 0, tot = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) {
      if (w < warp) woff += warp_tot[w];
      tot += warp_tot[w];
    }
    const int pos = base + woff + pre;
    if (f && pos < max_new) ind_out[pos] = c;
    __syncthreads();
    if (threadIdx.x == 0) base += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_new = base;
}

// RoPE + single-query attention for the decode harness (fused/attn.py:219-263).
// qkv row = [H*D | Hkv*D | Hkv*D]; the new key/value are appended at position past_len of the optional cache
// [M, Hkv, cap, D]; out[M, H*D] = softmax(q.K^T * scale) . V.  HF rotate_half convention (pairs i, i + D/2).
// One warp per (token, head): lane l owns the D/32 consecutive dims [l*E, l*E + E) (8-byte vector accesses); its rotation
// partner dims live in lane l ^ 16.  Outside the quantised hot path (the reference calls flash-attn here); kept simple.
template <int D>
struct AttnVec {
  static constexpr int E = D / 32;   // 2 (D = 64) or 4 (D = 128) halves per lane
  static __device__ __forceinline__ void ld(const __half* p_, float (&f)[E]) {
    if (E == 4) {
      const uint2 u = *reinterpret_cast<const uint2*>(p_);
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
      f[0] = a.x; f[1] = a.y; f[E - 2] = b.x; f[E - 1] = b.y;
    } else {
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(p_));
      f[0] = a.x; f[1] = a.y;
    }
  }
  // raw (packed fp16) form of the same vector: 2 registers (D = 128) or 1 (D = 64)
  static __device__ __forceinline__ uint2 ldraw(const __half* p_) {
    if (E == 4) return *reinterpret_cast<const uint2*>(p_);
    return make_uint2(*reinterpret_cast<const uint32_t*>(p_), 0u);
  }
  static __device__ __forceinline__ void unpack(const uint2 u, float (&f)[E]) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    f[0] = a.x; f[1] = a.y;
    if (E == 4) {
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
      f[E - 2] = b.x; f[E - 1] = b.y;
    }
  }
  static __device__ __forceinline__ void st(__half* p_, const float (&f)[E]) {
    if (E == 4) {
      const __half2 a = __floats2half2_rn(f[0], f[1]), b = __floats2half2_rn(f[E - 2], f[E - 1]);
      *reinterpret_cast<uint2*>(p_) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    } else {
      const __half2 a = __floats2half2_rn(f[0], f[1]);
      *reinterpret_cast<uint32_t*>(p_) = *reinterpret_cast<const uint32_t*>(&a);
    }
  }
};

// cos / sin of the rotary angle of dimension pair i at position pos.  Out of line on purpose: powf + sincosf are ~1.5 k SASS
// instructions when inlined per element, and none of it runs at position 0 (the benchflops regime).
static __device__ __noinline__ void rope_angle(float theta, int i, int D, int pos, float* sn, float* cs) {
  const float inv_freq = powf(theta, -2.0f * static_cast<float>(i) / static_cast<float>(D));
  sincosf(static_cast<float>(pos) * inv_freq, sn, cs);
}

// One (token m, head h) by one warp, q / k / v already in registers: RoPE on q and k, cache append, online-softmax
// attention over past_len + 1 keys.
template <int D>
__device__ __forceinline__ void attn_head_compute(const float (&q)[D / 32], const float (&k)[D / 32], const float (&v)[D / 32],
                                                  __half* k_cache, __half* v_cache, int cache_cap, int past_len, int m, int h,
                                                  int H, int Hkv, float theta, float scale, int lane, float (&acc)[D / 32]) {
  constexpr int E = D / 32;
  using V = AttnVec<D>;
  const int hk = h / (H / Hkv);
  float qr[E], kr[E];
  const float sgn = (lane < 16) ? -1.f : 1.f;      // dims < D/2 take -x[d + D/2], the others +x[d - D/2]
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const float qo = __shfl_xor_sync(0xffffffffu, q[e], 16), ko = __shfl_xor_sync(0xffffffffu, k[e], 16);
    float sn = 0.f, cs = 1.f;
    if (past_len > 0) rope_angle(theta, (lane * E + e) % (D / 2), D, past_len, &sn, &cs);
    // round to fp16 like the reference's fp16 tensors do
    qr[e] = __half2float(__float2half_rn(q[e] * cs + sgn * qo * sn));
    kr[e] = __half2float(__float2half_rn(k[e] * cs + sgn * ko * sn));
  }
  if (k_cache != nullptr && h % (H / Hkv) == 0) {
    V::st(k_cache + ((static_cast<size_t>(m) * Hkv + hk) * cache_cap + past_len) * D + lane * E, kr);
    V::st(v_cache + ((static_cast<size_t>(m) * Hkv + hk) * cache_cap + past_len) * D + lane * E, v);
  }
  float mx = -INFINITY, den = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) acc[e] = 0.f;
  for (int t = 0; t <= past_len; ++t) {
    float kt[E], vt[E];
    if (t == past_len) {
#pragma unroll
      for (int e = 0; e < E; ++e) { kt[e] = kr[e]; vt[e] = v[e]; }
    } else {
      V::ld(k_cache + ((static_cast<size_t>(m) * Hkv + hk) * cache_cap + t) * D + lane * E, kt);
      V::ld(v_cache + ((static_cast<size_t>(m) * Hkv + hk) * cache_cap + t) * D + lane * E, vt);
    }
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) s = fmaf(qr[e], kt[e], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    s *= scale;
    const float mn = fmaxf(mx, s);
    const float corr = __expf(mx - mn), pw = __expf(s - mn);
    den = den * corr + pw;
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = acc[e] * corr + pw * vt[e];
    mx = mn;
  }
#pragma unroll
  for (int e = 0; e < E; ++e) acc[e] = acc[e] / den;
}
// this head's slices of the qkv row of token m -> registers
template <int D>
__device__ __forceinline__ void attn_head_load(const __half* __restrict__ qkv, int m, int h, int H, int Hkv, int lane,
                                               float (&q)[D / 32], float (&k)[D / 32], float (&v)[D / 32]) {
  constexpr int E = D / 32;
  const int hk = h / (H / Hkv);
  const int ld = (H + 2 * Hkv) * D;
  const __half* row = qkv + static_cast<size_t>(m) * ld + lane * E;
  AttnVec<D>::ld(row + h * D, q);
  AttnVec<D>::ld(row + (H + hk) * D, k);
  AttnVec<D>::ld(row + (H + Hkv + hk) * D, v);
}
template <int D>
__device__ __forceinline__ void attn_head(const __half* __restrict__ qkv, __half* k_cache, __half* v_cache, int cache_cap,
                                          int past_len, int m, int h, int H, int Hkv, float theta, float scale, int lane,
                                          float (&acc)[D / 32]) {
  float q[D / 32], k[D / 32], v[D / 32];
  attn_head_load<D>(qkv, m, h, H, Hkv, lane, q, k, v);
  attn_head_compute<D>(q, k, v, k_cache, v_cache, cache_cap, past_len, m, h, H, Hkv, theta, scale, lane, acc);
}

template <int D>
__global__ void __launch_bounds__(256) rope_attn_decode_kernel(const __half* __restrict__ qkv, __half* k_cache,
                                                               __half* v_cache, int cache_cap, int past_len,
                                                               __half* __restrict__ out, int M, int H, int Hkv,
                                                               float theta, float scale) {
  constexpr int E = D / 32;
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (w >= static_cast<long long>(M) * H) return;
  const int m = static_cast<int>(w / H), h = static_cast<int>(w % H);
  float acc[E];
  attn_head<D>(qkv, k_cache, v_cache, cache_cap, past_len, m, h, H, Hkv, theta, scale, lane, acc);
  AttnVec<D>::st(out + (static_cast<size_t>(m) * H + h) * D + lane * E, acc);
}

// The same attention with the NEXT MixLinear's activation prologue folded in (o_proj: fused/attn.py:263 calls it in unfused
// mode, i.e. ExtractOutliersAndSetToZeros + FindRowScale on the attention output, linear.py:187-193): one CTA owns one token
// row, its 8 warps walk the row's heads, the fp16 row stays in shared memory, and process_row (rowquant.cuh) gathers / zeroes
// the outlier columns, takes the row abs-max and quantises — o_proj then runs with skip_prologue, exactly like the reference's
// "fused" call mode where a producer (norm.py:24-33) leaves q_xcache / x_scale / activation_outliers in the cache.
// rq.x = optional fp16 [M, H*D] copy of the attention output (outlier columns zeroed, as the reference leaves its tensor).
constexpr int kAttnQuantWarps = 4;   // 128 threads, small register footprint: many CTAs per SM, 512 token rows in ONE wave
// The qkv row of the token comes into shared memory with ONE bulk-async copy and the heads are walked by a ROLLED loop: this
// code runs once per CTA, so an unrolled per-head body is paced by cold instruction fetch (ncu: stall_no_inst 65 % for the
// 8-heads-in-registers version, 12.8 k SASS instructions).
template <int D>
__global__ void __launch_bounds__(kAttnQuantWarps * 32, 4)
rope_attn_decode_quant_kernel(const __half* __restrict__ qkv, __half* k_cache, __half* v_cache, int cache_cap, int past_len,
                              int H, int Hkv, float theta, float scale, const RowQuantArgs rq) {
  constexpr int E = D / 32;
  extern __shared__ __align__(128) uint8_t attn_smem[];
  __shared__ RowQuantSmem sm;
  __shared__ uint64_t row_bar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ld = (H + 2 * Hkv) * D;
  __half* qkv_s = reinterpret_cast<__half*>(attn_smem);                 // [ld] this token's q | k | v
  __half* row_s = qkv_s + ld;                                            // [H * D] attention output row
  if (threadIdx.x == 0) {
    mbar_init(&row_bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();
  uint32_t phase = 0;
  for (int m = blockIdx.x; m < rq.M; m += gridDim.x, phase ^= 1) {
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(&row_bar, static_cast<uint32_t>(ld) * 2u);
      bulk_load(qkv_s, qkv + static_cast<size_t>(m) * ld, static_cast<uint32_t>(ld) * 2u, &row_bar);
    }
    mbar_wait(&row_bar, phase, 8, m);
#pragma unroll 1
    for (int h = warp; h < H; h += kAttnQuantWarps) {
      const int hk = h / (H / Hkv);
      float q[E], k[E], v[E], acc[E];
      AttnVec<D>::ld(qkv_s + h * D + lane * E, q);
      AttnVec<D>::ld(qkv_s + (H + hk) * D + lane * E, k);
      AttnVec<D>::ld(qkv_s + (H + Hkv + hk) * D + lane * E, v);
      attn_head_compute<D>(q, k, v, k_cache, v_cache, cache_cap, past_len, m, h, H, Hkv, theta, scale, lane, acc);
      AttnVec<D>::st(row_s + h * D + lane * E, acc);
      if (rq.x != nullptr) AttnVec<D>::st(rq.x + static_cast<size_t>(m) * rq.K + h * D + lane * E, acc);
    }
    __syncthreads();
    process_row(rq, m, 0, threadIdx.x, 0, &sm, row_s);
    fence_proxy_async_smem();   // the next row's bulk copy overwrites qkv_s, which generic-proxy loads have just read
    __syncthreads();
  }
}

cudaError_t launch_rope_attn_decode(const __half* qkv, __half* k_cache, __half* v_cache, int cache_cap, int past_len,
                                    __half* out, int M, int H, int Hkv, int D, float theta, bool pdl, cudaStream_t st,
                                    const RowQuantArgs* rq) {
  const float scale = 1.0f / sqrtf(static_cast<float>(D));
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(rq != nullptr ? kAttnQuantWarps * 32 : 256);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  if (rq != nullptr) {
    cfg.gridDim = dim3(M);
    cfg.dynamicSmemBytes = static_cast<size_t>(2 * H + 2 * Hkv) * D * 2;   // the token's qkv row + its attention output row
    if (D == 128)
      return cudaLaunchKernelEx(&cfg, rope_attn_decode_quant_kernel<128>, qkv, k_cache, v_cache, cache_cap, past_len, H, Hkv, theta, scale, *rq);
    return cudaLaunchKernelEx(&cfg, rope_attn_decode_quant_kernel<64>, qkv, k_cache, v_cache, cache_cap, past_len, H, Hkv, theta, scale, *rq);
  }
  const long long warps = static_cast<long long>(M) * H;
  cfg.gridDim = dim3(static_cast<int>((warps + 7) / 8));
  if (D == 128)
    return cudaLaunchKernelEx(&cfg, rope_attn_decode_kernel<128>, qkv, k_cache, v_cache, cache_cap, past_len, out, M, H, Hkv, theta, scale);
  return cudaLaunchKernelEx(&cfg, rope_attn_decode_kernel<64>, qkv, k_cache, v_cache, cache_cap, past_len, out, M, H, Hkv, theta, scale);
}

// ---------------------------------------------------------------- all-reduce (+ residual) over NVLink peer memory
// The exchange step of the row-parallel Linears (o_proj, down_proj): every rank has written its fp16 partial [n] into its own
// peer-mapped buffer.  out = fp16( fp16(sum over ranks of partial, fp32, rank order) + residual ) on every rank, bit-identical.
//   one-shot (a.result == nullptr): each rank reads ALL partials (local HBM + peers over NVLink / NVSwitch) and computes the
//     whole vector: (world - 1) * n * 2 bytes of NVLink reads per rank, one handshake.  Best for two ranks.
//   two-shot: rank r reduces only elements [r * n / world, (r + 1) * n / world) and PUSHES that slice into the result buffer
//     of every rank (st over NVLink); a second handshake tells everybody that all slices have landed.  NVLink bytes per
//     rank: 2 * (world - 1) / world * n * 2 instead of (world - 1) * n * 2.  `out` is then the local result buffer itself.
//   flags[r][2 * p + phase]: word in rank r's memory that rank p sets to the exchange count it has reached.
//   *epoch: this rank's count of finished exchanges; every block reads it on entry, the last block to leave bumps it — so
//   the kernel is replayable from a CUDA graph.  Callers alternate between two partial (and result) buffers: a rank overwrites
//   the buffer of exchange e only after its exchange e + 1, which needed every peer's signal e + 1, which a peer sends only
//   after its own exchange e (its reads of that buffer) has completed.
__device__ __forceinline__ void peer_signal(const AllReduceArgs& a, int phase, uint32_t e) {
  __threadfence_system();
  for (int p = 0; p < a.world; ++p)
    if (p != a.rank) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.flags[p] + 2 * a.rank + phase), "r"(e) : "memory");
}
__device__ __forceinline__ void peer_wait(const AllReduceArgs& a, int phase, uint32_t e) {
  const uint64_t t0 = globaltimer_ns();
  for (int p = 0; p < a.world; ++p) {
    if (p == a.rank) continue;
    uint32_t v, spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a.flags[a.rank] + 2 * p + phase) : "memory");
      if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > a.timeout_ns) spin_timeout_trap(11 + phase, p, static_cast<int>(e));
    } while (static_cast<int32_t>(v - e) < 0);
  }
}

__global__ void __launch_bounds__(256) allreduce_residual_kernel(AllReduceArgs a) {
  pdl_launch_dependents();
  pdl_wait();                                   // my partial (previous kernel of the stream) is complete and visible
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) {
    const uint32_t e = *reinterpret_cast<volatile uint32_t*>(a.epoch) + 1;
    s_epoch = e;
    if (blockIdx.x == 0) peer_signal(a, 0, e);  // "my partial e is complete"
    peer_wait(a, 0, e);
  }
  __syncthreads();
  const int buf = a.buf;
  const bool two_shot = a.result[a.rank][buf] != nullptr;
  const long long nv = a.n >> 3;
  // two-shot: this rank owns vectors [v0, v1)
  const long long per = (nv + a.world - 1) / a.world;
  const long long v0 = two_shot ? per * a.rank : 0;
  const long long v1 = two_shot ? (v0 + per < nv ? v0 + per : nv) : nv;
  for (long long i = v0 + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < v1;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int p = 0; p < a.world; ++p) {
      H8 v;
      v.u = __ldcv(reinterpret_cast<const uint4*>(a.partial[p][buf]) + i);   // never from a stale L1 line
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(v.h2[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
    H8 r, o;
    if (a.residual != nullptr) r.u = *(reinterpret_cast<const uint4*>(a.residual) + i);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __half2 y2 = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
      if (a.residual != nullptr) {
        const float2 yf = __half22float2(y2), rf = __half22float2(r.h2[j]);
        y2 = __floats2half2_rn(__fadd_rn(yf.x, rf.x), __fadd_rn(yf.y, rf.y));
      }
      o.h2[j] = y2;
    }
    if (two_shot) {
      for (int p = 0; p < a.world; ++p) *(reinterpret_cast<uint4*>(a.result[p][buf]) + i) = o.u;   // push my slice to everybody
    } else {
      *(reinterpret_cast<uint4*>(a.out) + i) = o.u;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (atomicAdd(a.done, 1u) == gridDim.x - 1) {   // last block out: everybody has read *epoch and pushed its share
      if (two_shot) {
        peer_signal(a, 1, s_epoch);                  // "my slice of exchange e is in every result buffer"
        peer_wait(a, 1, s_epoch);                    // ... and everybody else's is in mine
      }
      *a.done = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(a.epoch) = s_epoch;
    }
  }
}

// ---------------------------------------------------------------- all-reduce (+ residual) through the NVSwitch (NVLS multicast)
// Every rank maps ONE symmetric allocation (partial 0/1, result 0/1, flag words at the same offsets) and a multicast address
// that reaches all copies at once.  Rank r owns elements [r * n / world, (r + 1) * n / world):
//   multimem.ld_reduce(partial[buf] slice)  — the switch adds the world copies (fp32 accumulation) and returns fp16: one
//                                              NVLink read of n / world elements per rank instead of (world - 1) * n / world;
//   + residual (a separate fp16 rounding, as all-reduce followed by `h + y` gives);
//   multimem.st(result[buf] slice)          — the switch writes the slice into every rank's result buffer.
// Handshakes are ONE multimem.red each: flags[phase] of every rank += 1, a rank proceeds when its own word reaches
// world * exchange count.  Phase 0: "my partial is complete"; phase 1: "my slice is in every result buffer".  The exchange
// count lives on the device (graph replayable); callers alternate buf = 0, 1 exactly as for the peer kernel.
__device__ __forceinline__ void mc_signal(const McAllReduceArgs& a, int phase) {
  __threadfence_system();
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(a.mc + a.flags_off + 4 * phase), "r"(1u) : "memory");
}
__device__ __forceinline__ void mc_wait(const McAllReduceArgs& a, int phase, uint32_t e) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a.local + a.flags_off) + phase;
  const uint32_t target = e * static_cast<uint32_t>(a.world);
  const uint64_t t0 = globaltimer_ns();
  uint32_t v, spins = 0;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(w) : "memory");
    if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > a.timeout_ns) spin_timeout_trap(21 + phase, static_cast<int>(v), static_cast<int>(target));
  } while (static_cast<int32_t>(v - target) < 0);
}

__global__ void __launch_bounds__(256) allreduce_multicast_kernel(McAllReduceArgs a) {
  pdl_launch_dependents();
  pdl_wait();                                   // my partial (previous kernel of the stream) is complete and visible
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) {
    const uint32_t e = *reinterpret_cast<volatile uint32_t*>(a.epoch) + 1;
    s_epoch = e;
    if (blockIdx.x == 0) mc_signal(a, 0);
    mc_wait(a, 0, e);
  }
  __syncthreads();
  const int buf = a.buf;
  const long long nv = a.n >> 3;                // 16-byte vectors
  const long long per = nv / a.world;
  const long long v0 = per * a.rank, v1 = v0 + per;
  const uint8_t* src = a.mc + a.partial_off[buf];
  uint8_t* dst = a.mc + a.result_off[buf];
  for (long long i = v0 + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < v1;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    H8 y;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0, %1, %2, %3}, [%4];"
                 : "=r"(y.u.x), "=r"(y.u.y), "=r"(y.u.z), "=r"(y.u.w)
                 : "l"(src + i * 16)
                 : "memory");
    if (a.residual != nullptr) {
      H8 r;
      r.u = *(reinterpret_cast<const uint4*>(a.residual) + i);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 yf = __half22float2(y.h2[j]), rf = __half22float2(r.h2[j]);
        y.h2[j] = __floats2half2_rn(__fadd_rn(yf.x, rf.x), __fadd_rn(yf.y, rf.y));
      }
    }
    asm volatile("multimem.st.relaxed.sys.global.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst + i * 16), "r"(y.u.x), "r"(y.u.y),
                 "r"(y.u.z), "r"(y.u.w)
                 : "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (atomicAdd(a.done, 1u) == gridDim.x - 1) {   // last block out: every block's slice stores have been issued and fenced
      mc_signal(a, 1);                               // "my slice of exchange e is in every result buffer"
      mc_wait(a, 1, s_epoch);                        // ... and everybody else's is in mine
      *a.done = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(a.epoch) = s_epoch;
    }
  }
}

// ---------------------------------------------------------------- fused exchange, second half (see XchgFinishArgs)
// Handshake counters: word `phase` of EVERY rank is incremented by every rank (one multimem.red through the switch, or one red
// per peer); a rank proceeds when its own word reaches the expected count.  Measured on NVSwitch (tools/bench_exchange.py,
// profiles/r02_exchange_anatomy_*): a flag takes 2.8 us one way, a system-scope fence 3 us on an idle SM and 6-9 us when stores
// to peers are in flight — so the kernel issues as few of them in series as the protocol allows:
//   phase 0: ONE fence + red by block 0 ("the pushes of my GEMM — the previous kernel of the stream — are performed");
//   phase 1: every block releases its own slice stores with one red (count = world * blocks per exchange), no block waits
//            for another block of its own grid; block 0 alone waits for the grand total and closes the exchange.
__device__ __forceinline__ void xf_red(const XchgFinishArgs& a, int phase, bool release) {
  if (a.mc_flags != nullptr) {
    if (release) asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(a.mc_flags + phase), "r"(1u) : "memory");
    else asm volatile("multimem.red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(a.mc_flags + phase), "r"(1u) : "memory");
  } else {
    for (int p = 0; p < a.world; ++p) {
      if (release && p == 0) asm volatile("fence.acq_rel.sys;" ::: "memory");
      asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(a.flags[p] + phase), "r"(1u) : "memory");
    }
  }
}
__device__ __forceinline__ void xf_wait(const XchgFinishArgs& a, int phase, uint32_t target) {
  const uint32_t* w = a.flags[a.rank] + phase;
  const uint64_t t0 = globaltimer_ns();
  uint32_t v, spins = 0;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(w) : "memory");
    if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > a.timeout_ns) spin_timeout_trap(31 + phase, static_cast<int>(v), static_cast<int>(target));
  } while (static_cast<int32_t>(v - target) < 0);
}

// ---------------------------------------------------------------- fused exchange, second half, sentinel-polling form
// No handshake at all (see XchgPollArgs): measured on NVSwitch a flag costs 2.8 us one way and a system-scope release after
// stores to peers 6-9 us (profiles/r02_exchange_anatomy_*), and the flag protocol needs two of each per exchange.  Here every
// wait is a spin on local memory for data that is already on its way:
//   own slice: for each 16-byte vector, wait until every rank's partial has landed in its slot (GEMM-epilogue pushes), add them
//   in fp32 in rank order, round, + residual (separate fp16 rounding), store the vector into EVERY rank's result buffer, put
//   the sentinel back into the slots;
//   other slices: wait until their owners' stores have landed in the local result buffer.
// Why nothing can be overwritten too early: a peer pushes into my slots of exchange e only after it has finished exchange e-1,
// which needed my slice of e-1, which I produced after my exchange e-2 had released these very slots (same stream); the same
// chain protects the result buffers.  The other result buffer (read last by this launch, as the residual) is re-armed here.
constexpr uint32_t kSentinel2 = 0xFFFFFFFFu;     // two fp16 sentinels
__device__ __forceinline__ bool has_sentinel(const uint4& v) {
  auto w = [](uint32_t x) { return (x & 0xFFFFu) == 0xFFFFu || (x >> 16) == 0xFFFFu; };
  return w(v.x) || w(v.y) || w(v.z) || w(v.w);
}
__device__ __forceinline__ uint4 poll_vec(const uint4* p, unsigned long long timeout_ns, int what) {
  uint4 v = __ldcv(p);
  if (!has_sentinel(v)) return v;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  do {
    v = __ldcv(p);
    if ((++spins & 0xff) == 0 && globaltimer_ns() - t0 > timeout_ns) spin_timeout_trap(what, static_cast<int>(blockIdx.x), static_cast<int>(threadIdx.x));
  } while (has_sentinel(v));
  return v;
}

__global__ void __launch_bounds__(256) exchange_finish_poll_kernel(XchgPollArgs a) {
  pdl_launch_dependents();
  pdl_wait();                                   // my GEMM (previous kernel of the stream) has pushed all of its tiles
  const bool tr = a.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  if (tr) a.trace[0] = globaltimer_ns();
  const bool one_shot = a.one_shot != 0;
  const int Ns = one_shot ? a.N : a.N / a.world;
  const int vpr = Ns >> 3;                      // 16-byte vectors per slot row
  const long long nv = static_cast<long long>(a.M) * vpr;
  const size_t slot = static_cast<size_t>(a.M) * Ns;
  const size_t col0 = one_shot ? 0 : static_cast<size_t>(a.rank) * Ns;
  const uint4 sent = make_uint4(kSentinel2, kSentinel2, kSentinel2, kSentinel2);
  const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long nthr = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = tid; i < nv; i += nthr) {
    const int row = static_cast<int>(i / vpr), c8 = static_cast<int>(i - static_cast<long long>(row) * vpr);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int s = 0; s < a.world; ++s) {
      uint4* sp = reinterpret_cast<uint4*>(a.recv + s * slot + static_cast<size_t>(row) * Ns) + c8;
      H8 v;
      v.u = poll_vec(sp, a.timeout_ns, 41);
      *sp = sent;                                // consumed: re-arm the slot for the exchange after the next one
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(v.h2[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
    const size_t off = static_cast<size_t>(row) * a.N + col0 + static_cast<size_t>(c8) * 8;
    H8 r, o;
    if (a.residual != nullptr) r.u = *reinterpret_cast<const uint4*>(a.residual + off);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __half2 y2 = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
      if (a.residual != nullptr) {
        const float2 yf = __half22float2(y2), rf = __half22float2(r.h2[j]);
        y2 = __floats2half2_rn(__fadd_rn(yf.x, rf.x), __fadd_rn(yf.y, rf.y));
      }
      o.h2[j] = y2;
    }
    if (a.reset != nullptr) *reinterpret_cast<uint4*>(a.reset + off) = sent;   // (after the residual read of the same vector)
    if (one_shot) {
      *reinterpret_cast<uint4*>(a.result[a.rank] + off) = o.u;
    } else if (a.mc_result != nullptr) {
      asm volatile("multimem.st.relaxed.sys.global.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(a.mc_result + off), "r"(o.u.x), "r"(o.u.y),
                   "r"(o.u.z), "r"(o.u.w)
                   : "memory");
    } else {
      for (int p = 0; p < a.world; ++p) *reinterpret_cast<uint4*>(a.result[p] + off) = o.u;
    }
  }
  if (tr) a.trace[3] = globaltimer_ns();
  if (!one_shot) {
    // the other ranks' slices: re-arm the dead buffer, then wait for the owners' stores to land here
    const int vprN = a.N >> 3;
    const long long nvN = static_cast<long long>(a.M) * vprN;
    const int lo = static_cast<int>(col0 >> 3), hi = lo + vpr;
    for (long long i = tid; i < nvN; i += nthr) {
      const int c = static_cast<int>(i % vprN);
      if (c >= lo && c < hi) {                  // my own slice: re-armed above; a multicast store comes back through the switch
        if (a.mc_result != nullptr) (void)poll_vec(reinterpret_cast<const uint4*>(a.result[a.rank]) + i, a.timeout_ns, 43);
        continue;
      }
      if (a.reset != nullptr) *(reinterpret_cast<uint4*>(a.reset) + i) = sent;
      (void)poll_vec(reinterpret_cast<const uint4*>(a.result[a.rank]) + i, a.timeout_ns, 42);
    }
  }
  if (tr) a.trace[7] = globaltimer_ns();
}

// ---------------------------------------------------------------- the same, + the NEXT MixLinear's activation prologue
// After an exchange every rank holds whole rows of the residual stream h, and the Linear that follows (W_pack / the SwiGLU pair)
// starts with RMSNorm -> outlier gather -> row abs-max -> quantise over exactly those rows (norm.py:24-33) — on EVERY rank, the
// same work, 3.5 us of prologue + a 1.8 us grid barrier per launch.  Here one CTA owns one token row: it reduces the row's own
// slice, broadcasts it, collects the other slices as they land (all through the row buffer in shared memory) and runs
// process_row (rowquant.cuh, the very code of the Linear's phase A) on it.  The Linear then runs with skip_prologue, the
// reference's "fused" call mode with this kernel as the producer.
constexpr int kXchgQuantThreads = 128;
__global__ void __launch_bounds__(kXchgQuantThreads, 4) exchange_finish_rowquant_kernel(XchgPollArgs a, const RowQuantArgs rq) {
  extern __shared__ __align__(128) uint8_t xq_smem[];
  __shared__ RowQuantSmem sm;
  __half* row_s = reinterpret_cast<__half*>(xq_smem);
  const uint32_t row_sa = smem_u32(row_s);
  pdl_launch_dependents();
  pdl_wait();
  const bool one_shot = a.one_shot != 0;
  const int Ns = one_shot ? a.N : a.N / a.world;
  const int vpr = Ns >> 3, vprN = a.N >> 3;
  const size_t slot = static_cast<size_t>(a.M) * Ns;
  const int lo = one_shot ? 0 : (a.rank * Ns) >> 3;       // first 16-byte vector of the own slice within a row
  const uint4 sent = make_uint4(kSentinel2, kSentinel2, kSentinel2, kSentinel2);
  for (int m = blockIdx.x, iter = 0; m < a.M; m += gridDim.x, ++iter) {
    const size_t rowoff = static_cast<size_t>(m) * a.N;
    for (int c8 = threadIdx.x; c8 < vpr; c8 += kXchgQuantThreads) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int s = 0; s < a.world; ++s) {
        uint4* sp = reinterpret_cast<uint4*>(a.recv + s * slot + static_cast<size_t>(m) * Ns) + c8;
        H8 v;
        v.u = poll_vec(sp, a.timeout_ns, 44);
        *sp = sent;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(v.h2[j]);
          acc[2 * j] += f.x;
          acc[2 * j + 1] += f.y;
        }
      }
      const size_t off = rowoff + static_cast<size_t>(lo + c8) * 8;
      H8 r, o;
      if (a.residual != nullptr) r.u = *reinterpret_cast<const uint4*>(a.residual + off);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __half2 y2 = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
        if (a.residual != nullptr) {
          const float2 yf = __half22float2(y2), rf = __half22float2(r.h2[j]);
          y2 = __floats2half2_rn(__fadd_rn(yf.x, rf.x), __fadd_rn(yf.y, rf.y));
        }
        o.h2[j] = y2;
      }
      if (a.reset != nullptr) *reinterpret_cast<uint4*>(a.reset + off) = sent;
      if (one_shot) {
        *reinterpret_cast<uint4*>(a.result[a.rank] + off) = o.u;
      } else if (a.mc_result != nullptr) {
        asm volatile("multimem.st.relaxed.sys.global.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(a.mc_result + off), "r"(o.u.x), "r"(o.u.y),
                     "r"(o.u.z), "r"(o.u.w)
                     : "memory");
      } else {
        for (int p = 0; p < a.world; ++p) *reinterpret_cast<uint4*>(a.result[p] + off) = o.u;
      }
      sts128(row_sa + (lo + c8) * 16, o.u);
    }
    if (!one_shot) {
      for (int c = threadIdx.x; c < vprN; c += kXchgQuantThreads) {
        const bool own = c >= lo && c < lo + vpr;
        if (own && a.mc_result == nullptr) continue;
        const size_t i = static_cast<size_t>(m) * vprN + c;
        if (!own && a.reset != nullptr) *(reinterpret_cast<uint4*>(a.reset) + i) = sent;
        const uint4 v = poll_vec(reinterpret_cast<const uint4*>(a.result[a.rank]) + i, a.timeout_ns, 45);
        if (!own) sts128(row_sa + c * 16, v);
      }
    }
    __syncthreads();
    process_row(rq, m, 0, threadIdx.x, iter, &sm, row_s);
    __syncthreads();
  }
}

cudaError_t launch_exchange_finish_rowquant(const XchgPollArgs& a, const RowQuantArgs& rq, int grid, bool pdl, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kXchgQuantThreads);
  cfg.dynamicSmemBytes = static_cast<size_t>(a.N) * 2;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, exchange_finish_rowquant_kernel, a, rq);
}

__global__ void pingpong_kernel(uint32_t* mine, uint32_t* peer, uint32_t* mc, int iters, int rank, unsigned long long* out_ns) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const uint64_t t0 = globaltimer_ns();
  for (int i = 1; i <= iters; ++i) {
    if (rank == 0) {
      if (mc) asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc), "r"(1u) : "memory");
      else asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer), "r"(static_cast<uint32_t>(i)) : "memory");
    }
    uint32_t v;
    const uint32_t target = mc ? static_cast<uint32_t>(2 * i - (rank == 0 ? 0 : 1)) : static_cast<uint32_t>(i);
    uint32_t spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (++spins > 400000000u) return;
    } while (static_cast<int32_t>(v - target) < 0);
    if (rank != 0) {
      if (mc) asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc), "r"(1u) : "memory");
      else asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer), "r"(static_cast<uint32_t>(i)) : "memory");
    }
  }
  *out_ns = globaltimer_ns() - t0;
}

__global__ void __launch_bounds__(256) exchange_finish_kernel(XchgFinishArgs a) {
  pdl_launch_dependents();
  pdl_wait();                                   // my GEMM (previous kernel of the stream) has pushed all of its tiles
  __shared__ uint32_t s_epoch;
  const bool tr = a.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  if (tr) a.trace[0] = globaltimer_ns();
  if (threadIdx.x == 0) {
    const uint32_t e = *reinterpret_cast<volatile uint32_t*>(a.epoch) + 1;
    s_epoch = e;
    if (!(a.ablate & 2)) {
      if (blockIdx.x == 0) xf_red(a, 0, true);  // "my pushes of exchange e are performed" (release: cumulative over pdl_wait)
      if (tr) a.trace[1] = globaltimer_ns();
      xf_wait(a, 0, e * static_cast<uint32_t>(a.world));   // ... and so are everybody's into my slots
    }
  }
  __syncthreads();
  if (tr) a.trace[2] = globaltimer_ns();
  const bool one_shot = a.one_shot != 0;        // every rank holds every rank's FULL partial: reduce everything locally, no broadcast
  const int Ns = one_shot ? a.N : a.N / a.world;
  const int vpr = Ns >> 3;                      // 16-byte vectors per slot row
  const long long nv = static_cast<long long>(a.M) * vpr;
  const size_t slot = static_cast<size_t>(a.M) * Ns;
  const size_t col0 = one_shot ? 0 : static_cast<size_t>(a.rank) * Ns;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < ((a.ablate & 4) ? 0 : nv);
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / vpr), c8 = static_cast<int>(i - static_cast<long long>(row) * vpr);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int s = 0; s < a.world; ++s) {
      H8 v;
      v.u = __ldcv(reinterpret_cast<const uint4*>(a.recv + s * slot + static_cast<size_t>(row) * Ns) + c8);   // written by peers: never from L1
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(v.h2[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
    const size_t off = static_cast<size_t>(row) * a.N + col0 + static_cast<size_t>(c8) * 8;
    H8 r, o;
    if (a.residual != nullptr) r.u = *reinterpret_cast<const uint4*>(a.residual + off);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __half2 y2 = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
      if (a.residual != nullptr) {
        const float2 yf = __half22float2(y2), rf = __half22float2(r.h2[j]);
        y2 = __floats2half2_rn(__fadd_rn(yf.x, rf.x), __fadd_rn(yf.y, rf.y));
      }
      o.h2[j] = y2;
    }
    if (one_shot) {
      *reinterpret_cast<uint4*>(a.result[a.rank] + off) = o.u;
    } else if (a.mc_result != nullptr) {
      asm volatile("multimem.st.relaxed.sys.global.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(a.mc_result + off), "r"(o.u.x), "r"(o.u.y),
                   "r"(o.u.z), "r"(o.u.w)
                   : "memory");
    } else {
      for (int p = 0; p < a.world; ++p) *reinterpret_cast<uint4*>(a.result[p] + off) = o.u;
    }
  }
  __syncthreads();                              // every thread's slice stores happen-before thread 0's release below
  if (tr) a.trace[3] = globaltimer_ns();
  if (threadIdx.x == 0) {
    const uint32_t e = s_epoch;
    if (one_shot || (a.ablate & 3)) {
      // nothing was sent to anybody: the exchange closes when every block of THIS grid has read the epoch
      __threadfence();
      if (atomicAdd(a.done, 1u) == gridDim.x - 1) {
        *a.done = 0;
        __threadfence();
        *reinterpret_cast<volatile uint32_t*>(a.epoch) = e;
      }
    } else {
      xf_red(a, 1, true);                       // "this block's part of my slice is in every result buffer"
      if (tr) a.trace[4] = globaltimer_ns();
      if (blockIdx.x == 0) {                    // every block of every rank has released: all slices are in MY result buffer,
        xf_wait(a, 1, e * static_cast<uint32_t>(a.world) * gridDim.x);   // and every block of this grid is past its epoch read
        if (tr) a.trace[7] = globaltimer_ns();
        *reinterpret_cast<volatile uint32_t*>(a.epoch) = e;
      }
    }
  }
}

// ---------------------------------------------------------------- QUIK MixedQLinear (mixquant/modules/qlinear.py:82-152)
// quik.asymmetric.quantize(x, int_indices, fp_indices, bits) (qlinear.py:117-120), one CTA per token row:
//   zero = min over the int columns, scale = (max - min) / (2^bits - 1) (both stored as fp16: meta[0][m], meta[1][m]),
//   q = rn((x - zero) / scale) - 2^(bits-1) in [-2^(b-1), 2^(b-1) - 1], one value per byte (the tcgen05 path has no int4 MMA);
//   fp_x = x[:, fp_indices].  HBM-bound byte work: x is read twice (second pass from L1/L2), 1 byte written per element.
__global__ void __launch_bounds__(256) quik_quantize_kernel(const __half* __restrict__ x, const int64_t* __restrict__ int_idx,
                                                            int n_int, const int64_t* __restrict__ fp_idx, int n_fp, int bits,
                                                            int8_t* __restrict__ q, __half* __restrict__ meta,
                                                            __half* __restrict__ fp_x, int M, int K) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_mn[8], s_mx[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    const __half* row = x + static_cast<size_t>(m) * K;
    float mn = INFINITY, mx = -INFINITY;
    for (int j = threadIdx.x; j < n_int; j += blockDim.x) {
      const float v = __half2float(row[int_idx[j]]);
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
    __syncthreads();
    mn = s_mn[0];
    mx = s_mx[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); }
    const float levels = static_cast<float>((1 << bits) - 1);
    const __half scale_h = __float2half_rn(__fdiv_rn(mx - mn, levels)), zero_h = __float2half_rn(mn);
    const float scale = __half2float(scale_h), zero = __half2float(zero_h);
    const float r = scale > 0.f ? __fdiv_rn(1.0f, scale) : 0.f;
    const float half_range = static_cast<float>(1 << (bits - 1));
    if (threadIdx.x == 0) {
      meta[m] = scale_h;
      meta[M + m] = zero_h;
    }
    int8_t* qrow = q + static_cast<size_t>(m) * n_int;
    for (int j = threadIdx.x; j < n_int; j += blockDim.x) {
      const float v = __half2float(row[int_idx[j]]);
      float t = rintf(__fmul_rn(v - zero, r)) - half_range;
      t = fminf(fmaxf(t, -half_range), half_range - 1.f);
      qrow[j] = static_cast<int8_t>(t);
    }
    for (int j = threadIdx.x; j < n_fp; j += blockDim.x) fp_x[static_cast<size_t>(m) * n_fp + j] = row[fp_idx[j]];
    __syncthreads();
  }
}
// The addend of quik.asymmetric.dequantize (qlinear.py:149-150): out[m,n] = fp16((zero[m] + 2^(b-1) scale[m]) * reduced_w[n]
// + fp_result[m,n]); the int GEMM's epilogue then adds it to acc * scale[m] * ws[n] (mixq_int4/int8_fused_dequantize).
__global__ void quik_addend_kernel(const __half* __restrict__ meta, const __half* __restrict__ reduced_w,
                                   const __half* __restrict__ fp_result, int ld_fp, __half* __restrict__ out, int M, int N, int bits) {
  pdl_launch_dependents();
  pdl_wait();
  const int nv = N >> 3;
  const long long total = static_cast<long long>(M) * nv;
  const float half_range = static_cast<float>(1 << (bits - 1));
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i / nv), n = static_cast<int>(i - static_cast<long long>(m) * nv) * 8;
    const float shift = __fadd_rn(__half2float(meta[M + m]), __fmul_rn(half_range, __half2float(meta[m])));
    H8 rw, fp, o;
    rw.u = __ldg(reinterpret_cast<const uint4*>(reduced_w + n));
    if (fp_result != nullptr) fp.u = *reinterpret_cast<const uint4*>(fp_result + static_cast<size_t>(m) * ld_fp + n);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = __fmul_rn(shift, __half2float(rw.h[j]));
      if (fp_result != nullptr) v = __fadd_rn(v, __half2float(fp.h[j]));
      o.h[j] = __float2half_rn(v);
    }
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(m) * N + n) = o.u;
  }
}

__global__ void mul_inplace_kernel(__half2* a, const __half2* __restrict__ b, long long n2) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n2;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    a[i] = __hmul2(a[i], b[i]);
}

}  // namespace mixq
