// Dequant epilogue shared by the 1-CTA and 2-CTA MixLinear kernels:
//   y = act( fp16( (f32(acc_int) * x_scale[m]) * scale_col[n] + fp16(acc_outl) [+ outl[m,n]] ) ), then the reference's
//   separate fp16 adds: + bias (linear.py:284-285), + residual (decoder layer).
#pragma once
#include "mixq_gemm.cuh"

namespace mixq {

__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }
// fp16 + fp16 the way torch does it on fp16 tensors: add in fp32, round once to fp16.
__device__ __forceinline__ __half2 hadd2_via_f32(__half2 a, __half2 b) {
  const float2 af = __half22float2(a), bf = __half22float2(b);
  return __floats2half2_rn(__fadd_rn(af.x, bf.x), __fadd_rn(af.y, bf.y));
}

// Where the fp16 output columns [n, ...) of a tile go (all columns of a tile share the answer): `base + row * ld + n`.
// Normally y [M,N]; with the tensor-parallel push the tile's column slice j = n / peer_cols is rank j's receive slot
// [M, peer_cols] (the returned base is shifted by -j * peer_cols so that the global column index still addresses it).
__device__ __forceinline__ __half* y_base(const LinearParams& p, int n, int& ld) {
  if (p.peer_cols > 0) {
    const int j = n / p.peer_cols;
    ld = p.peer_cols;
    return p.y_peer[j] - static_cast<ptrdiff_t>(j) * p.peer_cols;
  }
  ld = p.N;
  return p.y;
}

// Destinations of the tile that holds output column n: ONE (y, or the owning rank's slot) unless the broadcast push is on.
// A rolled loop around a single call site: the epilogue is instruction-fetch sensitive, its body must not be duplicated.
__device__ __forceinline__ int y_ndest(const LinearParams& p) { return p.peer_bcast > 0 ? p.peer_bcast : 1; }
__device__ __forceinline__ __half* y_dest(const LinearParams& p, int n, int d, int& ld) {
  if (p.peer_bcast > 0) {
    ld = p.N;
    return p.y_peer[d];
  }
  return y_base(p, n, ld);
}
template <class F>
__device__ __forceinline__ void for_each_ydest(const LinearParams& p, int n, F f) {
  const int nd = y_ndest(p);
#pragma unroll 1
  for (int d = 0; d < nd; ++d) {
    int ld;
    __half* b = y_dest(p, n, d, ld);
    f(b, ld);
  }
}

// One 32-column slab of one accumulator row: TMEM -> registers -> dequant (+outliers, +bias, SiLU) -> global.
// Every lane of the warp must call this (tcgen05.ld is warp-collective); row_ok masks the stores.
template <bool HAS_O>
__device__ __forceinline__ void epilogue_chunk(const LinearParams& p, uint32_t t_int, uint32_t t_out, int row,
                                               bool row_ok, int n, float xs) {
  uint32_t acc[32];
  uint32_t oacc[32];
  tmem_ld_32x32(t_int, acc);
  if (HAS_O) tmem_ld_32x32(t_out, oacc);
  tmem_ld_wait();
  if (!row_ok || n >= p.N) return;
  if (p.epilogue == EPI_RAW_I32) {
    int32_t* dst = p.y_i32 + static_cast<size_t>(row) * p.N + n;
#pragma unroll
    for (int g = 0; g < 8; ++g)
      if (n + g * 4 < p.N)
        reinterpret_cast<uint4*>(dst)[g] = make_uint4(acc[4 * g], acc[4 * g + 1], acc[4 * g + 2], acc[4 * g + 3]);
    return;
  }
  const bool has_outl = p.outl != nullptr;
  const bool has_bias = p.bias != nullptr;
  const bool has_res = p.residual != nullptr;
#pragma unroll
  for (int g = 0; g < 4; ++g) {   // 8 columns per 16-byte store
    const int ng = n + g * 8;
    if (ng < p.N) {
      const uint4 wsu = __ldg(reinterpret_cast<const uint4*>(p.scale_col + ng));
      uint4 olu = make_uint4(0, 0, 0, 0), bsu = make_uint4(0, 0, 0, 0), rsu = make_uint4(0, 0, 0, 0);
      if (has_outl) olu = *reinterpret_cast<const uint4*>(p.outl + static_cast<size_t>(row) * p.ld_outl + ng);
      if (has_bias) bsu = __ldg(reinterpret_cast<const uint4*>(p.bias + ng));
      if (has_res) rsu = *reinterpret_cast<const uint4*>(p.residual + static_cast<size_t>(row) * p.ld_res + ng);
      const uint32_t wsw[4] = {wsu.x, wsu.y, wsu.z, wsu.w};
      const uint32_t olw[4] = {olu.x, olu.y, olu.z, olu.w};
      const uint32_t bsw[4] = {bsu.x, bsu.y, bsu.z, bsu.w};
      const uint32_t rsw[4] = {rsu.x, rsu.y, rsu.z, rsu.w};
      uint32_t ow[4];
#pragma unroll
      for (int j2 = 0; j2 < 4; ++j2) {
        const float2 wf = __half22float2(*reinterpret_cast<const __half2*>(&wsw[j2]));
        const float2 of = __half22float2(*reinterpret_cast<const __half2*>(&olw[j2]));
        float v[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = g * 8 + j2 * 2 + h;
          float t = __fmul_rn(__fmul_rn(static_cast<float>(static_cast<int32_t>(acc[c])), xs), h ? wf.y : wf.x);
          // the reference's torch.mm(activation_outliers, weight_cache.T) returns fp16 (linear.py:248):
          // round the fp32 tensor-core sum to fp16 before it joins the dequantised int part
          if (HAS_O) t = __fadd_rn(t, __half2float(__float2half_rn(__uint_as_float(oacc[c]))));
          if (has_outl) t = __fadd_rn(t, h ? of.y : of.x);
          if (p.act == 1) t = silu_f(t);
          v[h] = t;
        }
        __half2 o2 = __floats2half2_rn(v[0], v[1]);
        // y1 += bias (linear.py:284-285) and the decoder's residual add are separate fp16 ops in the
        // reference: each rounds to fp16 again
        if (has_bias) o2 = hadd2_via_f32(o2, *reinterpret_cast<const __half2*>(&bsw[j2]));
        if (has_res) o2 = hadd2_via_f32(o2, *reinterpret_cast<const __half2*>(&rsw[j2]));
        ow[j2] = *reinterpret_cast<const uint32_t*>(&o2);
      }
      const uint4 ov = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      for_each_ydest(p, n, [&](__half* yb, int ldy) { *reinterpret_cast<uint4*>(yb + static_cast<size_t>(row) * ldy + ng) = ov; });
    }
  }
}


// A span of `ncols` (multiple of 16) accumulator columns of one row, as a ROLLED loop over 16-column groups: the body is
// ~200 instructions and stays in the instruction cache (the fully unrolled 32-column form above is paced by
// instruction fetch when only 1-2 warps per scheduler run it).  Every lane of the warp must call this.
template <bool HAS_O>
__device__ __forceinline__ void epilogue_span(const LinearParams& p, uint32_t t_int, uint32_t t_out, int row, bool row_ok,
                                              int n0, int ncols, float xs, const __half* addend, int ld_addend) {
  const bool has_outl = addend != nullptr;
  const bool has_bias = p.bias != nullptr;
  const bool has_res = p.residual != nullptr;
  const bool raw = p.epilogue == EPI_RAW_I32;
  const bool silu = p.act == 1;
#pragma unroll 1
  for (int c = 0; c < ncols; c += 16) {
    uint32_t acc[16];
    uint32_t oacc[16];
    tmem_ld_32x16(t_int + c, acc);
    if (HAS_O) tmem_ld_32x16(t_out + c, oacc);
    tmem_ld_wait();
    const int n = n0 + c;
    if (!row_ok || n >= p.N) continue;
    if (raw) {
      int32_t* dst = p.y_i32 + static_cast<size_t>(row) * p.N + n;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (n + g * 4 < p.N)
          reinterpret_cast<uint4*>(dst)[g] = make_uint4(acc[4 * g], acc[4 * g + 1], acc[4 * g + 2], acc[4 * g + 3]);
      continue;
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {   // 8 columns per 16-byte store
      const int ng = n + g * 8;
      if (ng < p.N) {
        const uint4 wsu = __ldg(reinterpret_cast<const uint4*>(p.scale_col + ng));
        uint4 olu = make_uint4(0, 0, 0, 0), bsu = make_uint4(0, 0, 0, 0), rsu = make_uint4(0, 0, 0, 0);
        if (has_outl) olu = *reinterpret_cast<const uint4*>(addend + static_cast<size_t>(row) * ld_addend + ng);
        if (has_bias) bsu = __ldg(reinterpret_cast<const uint4*>(p.bias + ng));
        if (has_res) rsu = *reinterpret_cast<const uint4*>(p.residual + static_cast<size_t>(row) * p.ld_res + ng);
        const uint32_t wsw[4] = {wsu.x, wsu.y, wsu.z, wsu.w};
        const uint32_t olw[4] = {olu.x, olu.y, olu.z, olu.w};
        const uint32_t bsw[4] = {bsu.x, bsu.y, bsu.z, bsu.w};
        const uint32_t rsw[4] = {rsu.x, rsu.y, rsu.z, rsu.w};
        uint32_t ow[4];
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {
          const float2 wf = __half22float2(*reinterpret_cast<const __half2*>(&wsw[j2]));
          const float2 of = __half22float2(*reinterpret_cast<const __half2*>(&olw[j2]));
          float v[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int cc = g * 8 + j2 * 2 + h;
            float t = __fmul_rn(__fmul_rn(static_cast<float>(static_cast<int32_t>(acc[cc])), xs), h ? wf.y : wf.x);
            // torch.mm(activation_outliers, weight_cache.T) returns fp16 (linear.py:248): round the fp32 tensor-core sum
            if (HAS_O) t = __fadd_rn(t, __half2float(__float2half_rn(__uint_as_float(oacc[cc]))));
            if (has_outl) t = __fadd_rn(t, h ? of.y : of.x);
            if (silu) t = silu_f(t);
            v[h] = t;
          }
          __half2 o2 = __floats2half2_rn(v[0], v[1]);
          if (has_bias) o2 = hadd2_via_f32(o2, *reinterpret_cast<const __half2*>(&bsw[j2]));
          if (has_res) o2 = hadd2_via_f32(o2, *reinterpret_cast<const __half2*>(&rsw[j2]));
          ow[j2] = *reinterpret_cast<const uint32_t*>(&o2);
        }
        const uint4 ov = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        for_each_ydest(p, n0, [&](__half* yb, int ldy) { *reinterpret_cast<uint4*>(yb + static_cast<size_t>(row) * ldy + ng) = ov; });
      }
    }
  }
}


// Split-K (1-CTA kernel).  A partial int32 accumulator tile (128 rows x 128 columns) travels through a 64 KB workspace block
// laid out [16-column group g][row][16]: every tcgen05.ld group of a warp (32 rows x 64 bytes) is 2 KB of contiguous memory —
// coalesced stores on the way out, ONE bulk copy per block into the finisher's (idle) pipeline stages on the way in, and
// conflict-free 16-byte shared-memory reads in the fold.  Every lane of the warp must call these (tcgen05.ld / st are
// warp-collective); row_in_tile = TMEM lane = 32 * quarter + lane.
constexpr int kSplitKBlockInts = 128 * 128;
__device__ __forceinline__ void splitk_stage_partial(uint32_t t_int, uint32_t block_sa, int row_in_tile) {
  // TMEM -> the block layout in (idle) shared memory; one bulk store moves it to the workspace afterwards
#pragma unroll 1
  for (int g = 0; g < 8; g += 2) {
    uint32_t acc[32];
    tmem_ld_32x32(t_int + g * 16, acc);
    tmem_ld_wait();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t dst = block_sa + static_cast<uint32_t>((g + h) * 128 + row_in_tile) * 64u;
#pragma unroll
      for (int v = 0; v < 4; ++v)
        sts128(dst + v * 16, make_uint4(acc[16 * h + 4 * v], acc[16 * h + 4 * v + 1], acc[16 * h + 4 * v + 2], acc[16 * h + 4 * v + 3]));
    }
  }
}
// The tile's last split: add `nparts` blocks (already in shared memory at parts_sa, 64 KB apart) to the accumulator, in place in
// TMEM.  Integer adds: exact and order-independent, so the dequant epilogue that follows sees the unsplit accumulator.
__device__ __forceinline__ void splitk_fold_smem(uint32_t t_int, uint32_t parts_sa, int nparts, int row_in_tile) {
#pragma unroll 1
  for (int g = 0; g < 8; ++g) {
    uint32_t acc[16];
    tmem_ld_32x16(t_int + g * 16, acc);
    tmem_ld_wait();
    for (int sp = 0; sp < nparts; ++sp) {
      const uint32_t src = parts_sa + static_cast<uint32_t>(sp) * (kSplitKBlockInts * 4) + static_cast<uint32_t>(g * 128 + row_in_tile) * 64u;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const uint4 pv = lds128(src + v * 16);
        acc[4 * v] += pv.x;
        acc[4 * v + 1] += pv.y;
        acc[4 * v + 2] += pv.z;
        acc[4 * v + 3] += pv.w;
      }
    }
    tmem_st_32x16(t_int + g * 16, acc);
  }
  tmem_st_wait();
}

}  // namespace mixq

// ================================================================ coalesced epilogue (2-CTA kernel)
// A thread owns one accumulator ROW (that is how tcgen05.ld hands out TMEM), so direct 16-byte stores from the 32 lanes of
// a warp touch 32 different rows: 32 memory transactions of 16 bytes per instruction (measured: 12 us to drain a
// 128 x 352 tile).  Each epilogue warp therefore owns a 4 KB staging tile (32 rows x 64 fp16, 16-byte chunks XOR-swizzled
// by row so that both the row-wise and the 8-lanes-per-row accesses are conflict free) and moves 64 columns at a time
// between it and global memory with 8 lanes per row: 4 full 128-byte lines per instruction.
//
// The epilogue is bound by INSTRUCTION ISSUE (ncu + clock64 profile, profiles/r01_ncu_summary.md: tcgen05.ld+wait is 3 %
// of it; two epilogue warps per scheduler at ~260 SASS instructions per 16 columns): the variants below are templated on
// what the call actually needs, address shared memory through 32-bit shared-window addresses (ld/st.shared, no generic
// 64-bit pointer arithmetic), and bring residual / addend tiles in through the same coalesced staging path.
// -DMIXQ_EPI_TRACE (tuning builds only): clock64 stamps of CTA 0 / warp 4 inside the outlier-pass epilogue, 8 per call,
// at p.trace[1800 + 8 * call]
#ifdef MIXQ_EPI_TRACE
#define MIXQ_EPI_STAMP(j)                                                                                          \
  do {                                                                                                             \
    if (p.trace != nullptr && blockIdx.x == 0 && (threadIdx.x >> 5) == 4 && lane == 0 && p.trace[1799] < 24)       \
      p.trace[1800 + 8 * p.trace[1799] + (j)] = static_cast<unsigned long long>(clock64());                                                  \
  } while (0)
#define MIXQ_EPI_NEXT()                                                                                            \
  do {                                                                                                             \
    if (p.trace != nullptr && blockIdx.x == 0 && (threadIdx.x >> 5) == 4 && lane == 0) p.trace[1799] += 1;         \
  } while (0)
#else
#define MIXQ_EPI_STAMP(j) ((void)0)
#define MIXQ_EPI_NEXT() ((void)0)
#endif

namespace mixq {

constexpr int kEpiStageBytes = 32 * 128;

// stage[r][c] (r < 32 rows, c < 8 chunks of 16 bytes) lives at stage_sa + r * 128 + ((c ^ (r & 7)) << 4)
__device__ __forceinline__ uint32_t epi_slot_sa(uint32_t stage_sa, int row, int chunk) {
  return stage_sa + row * 128 + ((chunk ^ (row & 7)) << 4);
}
// global [m_base + r][n_blk + 8c] -> stage[r][c]  for r < 32, c < nchunks (rows >= M / columns >= N skipped).
// Lane l moves chunk c = l & 7 of rows (l >> 3) + 4 it: slot address = (base0 + 512 it) ^ ((it & 1) << 6).
__device__ __forceinline__ void epi_stage_in(uint32_t stage_sa, const __half* g, int ld, int m_base, int n_blk, int nchunks, int M,
                                             int N, int lane) {
  const int c = lane & 7, r0 = lane >> 3;
  if (c >= nchunks || n_blk + c * 8 >= N) return;
  const __half* src = g + static_cast<size_t>(m_base + r0) * ld + n_blk + c * 8;
  const size_t step = static_cast<size_t>(4) * ld;
  const uint32_t base0 = stage_sa + r0 * 128 + ((c ^ r0) << 4);
  if (m_base + 32 <= M) {
    uint4 v[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) v[it] = *reinterpret_cast<const uint4*>(src + it * step);
#pragma unroll
    for (int it = 0; it < 8; ++it) sts128((base0 + it * 512) ^ ((it & 1) << 6), v[it]);
  } else {
#pragma unroll
    for (int it = 0; it < 8; ++it)
      if (m_base + it * 4 + r0 < M) sts128((base0 + it * 512) ^ ((it & 1) << 6), *reinterpret_cast<const uint4*>(src + it * step));
  }
}
__device__ __forceinline__ void epi_stage_out(uint32_t stage_sa, __half* g, int ld, int m_base, int n_blk, int nchunks, int M,
                                              int N, int lane) {
  const int c = lane & 7, r0 = lane >> 3;
  if (c >= nchunks || n_blk + c * 8 >= N) return;
  __half* dst = g + static_cast<size_t>(m_base + r0) * ld + n_blk + c * 8;
  const size_t step = static_cast<size_t>(4) * ld;
  const uint32_t base0 = stage_sa + r0 * 128 + ((c ^ r0) << 4);
  if (m_base + 32 <= M) {
#pragma unroll
    for (int it = 0; it < 8; ++it) *reinterpret_cast<uint4*>(dst + it * step) = lds128((base0 + it * 512) ^ ((it & 1) << 6));
  } else {
#pragma unroll
    for (int it = 0; it < 8; ++it)
      if (m_base + it * 4 + r0 < M) *reinterpret_cast<uint4*>(dst + it * step) = lds128((base0 + it * 512) ^ ((it & 1) << 6));
  }
}

// v = (f32(acc) * xs) * ws [+ fp16(f32 outlier sum)] for the two accumulator columns that share one 32-bit scale word.
// torch.mm(activation_outliers, weight_cache.T) returns fp16 (linear.py:248): the fp32 tensor-core sum is rounded to fp16
// before it joins the dequantised int part.
template <bool HAS_O>
__device__ __forceinline__ float2 dequant2(uint32_t a0, uint32_t a1, uint32_t o0, uint32_t o1, float xs, uint32_t ws2) {
  const float2 wf = __half22float2(*reinterpret_cast<const __half2*>(&ws2));
  float2 v;
  v.x = __fmul_rn(__fmul_rn(static_cast<float>(static_cast<int32_t>(a0)), xs), wf.x);
  v.y = __fmul_rn(__fmul_rn(static_cast<float>(static_cast<int32_t>(a1)), xs), wf.y);
  if (HAS_O) {
    const float2 of = __half22float2(__floats2half2_rn(__uint_as_float(o0), __uint_as_float(o1)));
    v.x = __fadd_rn(v.x, of.x);
    v.y = __fadd_rn(v.y, of.y);
  }
  return v;
}

// 16 accumulator columns of one row -> two 16-byte chunks of the staging tile.
template <bool HAS_O, int MODE, int OFF, int NA>
__device__ __forceinline__ void epilogue_group16(const LinearParams& p, const uint32_t (&acc)[NA], const uint32_t (&oacc)[NA],
                                                 float xs, uint32_t scale_sa16, uint32_t row_sa, uint32_t sw, int chunk0, int row,
                                                 bool row_ok, int n16) {
  const bool has_add = MODE == 2 && p.outl != nullptr;
  const bool has_bias = MODE == 2 && p.bias != nullptr;
  const bool has_res = MODE == 2 && p.residual != nullptr;
  const bool silu = MODE == 2 && p.act == 1;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const uint4 wsu = lds128(scale_sa16 + g * 16);
    const uint32_t slot = row_sa + (((chunk0 + g) ^ sw) << 4);
    const uint32_t wsw[4] = {wsu.x, wsu.y, wsu.z, wsu.w};
    uint32_t ow[4];
    if (MODE == 0) {
#pragma unroll
      for (int j2 = 0; j2 < 4; ++j2) {
        const int cc = OFF + g * 8 + j2 * 2;
        const float2 v = dequant2<HAS_O>(acc[cc], acc[cc + 1], oacc[cc], oacc[cc + 1], xs, wsw[j2]);
        const __half2 o2 = __floats2half2_rn(v.x, v.y);
        ow[j2] = *reinterpret_cast<const uint32_t*>(&o2);
      }
    } else if (MODE == 1) {
      const uint4 rsu = lds128(slot);
      const uint32_t rsw[4] = {rsu.x, rsu.y, rsu.z, rsu.w};
#pragma unroll
      for (int j2 = 0; j2 < 4; ++j2) {
        const int cc = OFF + g * 8 + j2 * 2;
        const float2 v = dequant2<HAS_O>(acc[cc], acc[cc + 1], oacc[cc], oacc[cc + 1], xs, wsw[j2]);
        // the decoder's residual add is a separate fp16 op in the reference: round, then add in fp32, round again
        const __half2 o2 = hadd2_via_f32(__floats2half2_rn(v.x, v.y), *reinterpret_cast<const __half2*>(&rsw[j2]));
        ow[j2] = *reinterpret_cast<const uint32_t*>(&o2);
      }
    } else {
      const int ng = n16 + g * 8;
      const bool ok = row_ok && ng < p.N;
      uint4 olu = make_uint4(0, 0, 0, 0), bsu = make_uint4(0, 0, 0, 0), rsu = make_uint4(0, 0, 0, 0);
      if (has_add) olu = lds128(slot);
      if (has_bias && ng < p.N) bsu = __ldg(reinterpret_cast<const uint4*>(p.bias + ng));
      if (has_res && ok) rsu = *reinterpret_cast<const uint4*>(p.residual + static_cast<size_t>(row) * p.ld_res + ng);
      const uint32_t olw[4] = {olu.x, olu.y, olu.z, olu.w};
      const uint32_t bsw[4] = {bsu.x, bsu.y, bsu.z, bsu.w};
      const uint32_t rsw[4] = {rsu.x, rsu.y, rsu.z, rsu.w};
#pragma unroll
      for (int j2 = 0; j2 < 4; ++j2) {
        const int cc = OFF + g * 8 + j2 * 2;
        float2 v = dequant2<HAS_O>(acc[cc], acc[cc + 1], oacc[cc], oacc[cc + 1], xs, wsw[j2]);
        if (has_add) {
          const float2 of = __half22float2(*reinterpret_cast<const __half2*>(&olw[j2]));
          v.x = __fadd_rn(v.x, of.x);
          v.y = __fadd_rn(v.y, of.y);
        }
        if (silu) {
          v.x = silu_f(v.x);
          v.y = silu_f(v.y);
        }
        __half2 o2 = __floats2half2_rn(v.x, v.y);
        // y1 += bias (linear.py:284-285) and the residual add are separate fp16 ops in the reference
        if (has_bias) o2 = hadd2_via_f32(o2, *reinterpret_cast<const __half2*>(&bsw[j2]));
        if (has_res) o2 = hadd2_via_f32(o2, *reinterpret_cast<const __half2*>(&rsw[j2]));
        ow[j2] = *reinterpret_cast<const uint32_t*>(&o2);
      }
    }
    sts128(slot, make_uint4(ow[0], ow[1], ow[2], ow[3]));
  }
}

// One contiguous run of `ncols` (multiple of 16) output columns of a tile for the 32 rows of this warp:
//   y = act(fp16((f32(acc_int) * xs) * ws + fp16(acc_outl) [+ addend])) [+ bias] [+ residual]
// t_int / t_outl: TMEM addresses (lane quarter included) of the first int32 / fp32-outlier accumulator column of the run
// (t_outl unused when !HAS_O); scale_sa: shared-window address of scale_col (fp16) of the run's first column.
// MODE 0: nothing else.  MODE 1: + residual (staged in through the tile; `pre_staged`: the caller already brought the
// first 64-column block in while it waited for the accumulator).  MODE 2: any of addend / bias / residual / SiLU.
// Without outliers the accumulator is read 32 columns per tcgen05.ld / wait (half the exposed round trips; a software-
// pipelined second register set was tried: ptxas keeps it in local memory).
template <bool HAS_O, int MODE>
__device__ __forceinline__ void epilogue_run_coalesced(const LinearParams& p, uint32_t stage_sa, uint32_t t_int, uint32_t t_outl,
                                                       int m_base, int n0, int ncols, float xs, uint32_t scale_sa, int lane,
                                                       bool pre_staged = false) {
  const bool has_add = MODE == 2 && p.outl != nullptr;
  const int row = m_base + lane;
  const bool row_ok = row < p.M;
  const uint32_t row_sa = stage_sa + lane * 128;
  const uint32_t sw = lane & 7;
  const int ngroups = ncols >> 4;
#pragma unroll 1
  for (int g = 0; g < ngroups; g += 2) {
    if ((g & 3) == 0) {     // first group of a 64-column block: bring the residual / addend tile in
      const int bc = (ncols - g * 16 < 64) ? (ncols - g * 16) : 64;
      if (MODE == 1 && !(pre_staged && g == 0)) epi_stage_in(stage_sa, p.residual, p.ld_res, m_base, n0 + g * 16, bc >> 3, p.M, p.N, lane);
      else if (has_add) epi_stage_in(stage_sa, p.outl, p.ld_outl, m_base, n0 + g * 16, bc >> 3, p.M, p.N, lane);
      if (MODE == 1 || has_add) __syncwarp();
    }
    const bool two = g + 1 < ngroups;
    if (!HAS_O && two && !(p.ablate & 8)) {
      uint32_t acc[32];
      tmem_ld_32x32(t_int + g * 16, acc);
      tmem_ld_wait();
      epilogue_group16<false, MODE, 0, 32>(p, acc, acc, xs, scale_sa + g * 32, row_sa, sw, (g & 3) * 2, row, row_ok, n0 + g * 16);
      epilogue_group16<false, MODE, 16, 32>(p, acc, acc, xs, scale_sa + (g + 1) * 32, row_sa, sw, ((g + 1) & 3) * 2, row, row_ok, n0 + (g + 1) * 16);
    } else {
#pragma unroll 1
      for (int h = 0; h < (two ? 2 : 1); ++h) {
        uint32_t acc[16];
        uint32_t oacc[16];
        MIXQ_EPI_STAMP(0 + 3 * h);
        tmem_ld_32x16(t_int + (g + h) * 16, acc);
        if (HAS_O) tmem_ld_32x16(t_outl + (g + h) * 16, oacc);
        tmem_ld_wait();
        MIXQ_EPI_STAMP(1 + 3 * h);
        epilogue_group16<HAS_O, MODE, 0, 16>(p, acc, oacc, xs, scale_sa + (g + h) * 32, row_sa, sw, ((g + h) & 3) * 2, row, row_ok, n0 + (g + h) * 16);
        MIXQ_EPI_STAMP(2 + 3 * h);
      }
    }
    const int last = two ? g + 1 : g;
    if ((last & 3) == 3 || last == ngroups - 1) {   // block complete (or run finished): 64 columns out, coalesced
      __syncwarp();
      MIXQ_EPI_STAMP(6);
      for_each_ydest(p, n0, [&](__half* yb, int ldy) {
        epi_stage_out(stage_sa, yb, ldy, m_base, n0 + (g & ~3) * 16, ((last & 3) + 1) * 2, p.M, p.N, lane);
      });
      __syncwarp();
      MIXQ_EPI_STAMP(7);
    }
  }
  MIXQ_EPI_NEXT();
}

}  // namespace mixq

// ================================================================ SwiGLU pair epilogue (2-CTA kernel, pair mode)
// In pair mode CTA 0 of the pair stages gate_proj's weight rows and CTA 1 up_proj's rows for the SAME output columns, so in
// every CTA each MMA column chunk holds gate columns in its first half and the matching up columns in its second half.  The
// reference (fused/mlp.py:61-64) does
//     up = up_proj(x);  gate = gate_proj.forward_without_preconditionFusedSilu(x);  gate *= up        (all fp16 tensors)
// so y = fp16( fp16(silu(v_gate)) * fp16(v_up) ).  Each epilogue warp takes every other 32-column block of the run and reads
// BOTH accumulators for it, so gate and up meet in registers: no cross-warp exchange, no named barriers (the first version had
// a gate warp publishing through shared memory to an up warp — 40 % of its time was the two barriers per block, ncu r02).
namespace mixq {

template <bool HAS_O>
__device__ __forceinline__ void epilogue_run_swiglu(const LinearParams& p, uint32_t my_sa, uint32_t t_g, uint32_t t_u, uint32_t t_og,
                                                    uint32_t t_ou, int m_base, int n0, int ncols, float xs, uint32_t sc_g_sa,
                                                    uint32_t sc_u_sa, int lane, int half) {
  const uint32_t sw = lane & 7;
  const uint32_t my_row = my_sa + lane * 128;
#pragma unroll 1
  for (int c0 = half * 32; c0 < ncols; c0 += 64) {
    const int bc = (ncols - c0 < 32) ? (ncols - c0) : 32;     // 16 or 32 columns
#pragma unroll 1
    for (int c = 0; c < bc; c += 16) {
      uint32_t g[16], u[16], og[16], ou[16];
      tmem_ld_32x16(t_g + c0 + c, g);
      tmem_ld_32x16(t_u + c0 + c, u);
      if (HAS_O) {
        tmem_ld_32x16(t_og + c0 + c, og);
        tmem_ld_32x16(t_ou + c0 + c, ou);
      }
      tmem_ld_wait();
#pragma unroll
      for (int gi = 0; gi < 2; ++gi) {
        const uint4 wg4 = lds128(sc_g_sa + (c0 + c + gi * 8) * 2);
        const uint4 wu4 = lds128(sc_u_sa + (c0 + c + gi * 8) * 2);
        const uint32_t wg[4] = {wg4.x, wg4.y, wg4.z, wg4.w};
        const uint32_t wu[4] = {wu4.x, wu4.y, wu4.z, wu4.w};
        uint32_t ow[4];
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {
          const int cc = gi * 8 + j2 * 2;
          float2 vg = dequant2<HAS_O>(g[cc], g[cc + 1], og[cc], og[cc + 1], xs, wg[j2]);
          const float2 vu = dequant2<HAS_O>(u[cc], u[cc + 1], ou[cc], ou[cc + 1], xs, wu[j2]);
          vg.x = silu_f(vg.x);
          vg.y = silu_f(vg.y);
          const __half2 pr = __hmul2(__floats2half2_rn(vg.x, vg.y), __floats2half2_rn(vu.x, vu.y));
          ow[j2] = *reinterpret_cast<const uint32_t*>(&pr);
        }
        sts128(my_row + ((static_cast<uint32_t>((c >> 3) + gi) ^ sw) << 4), make_uint4(ow[0], ow[1], ow[2], ow[3]));
      }
    }
    __syncwarp();
    epi_stage_out(my_sa, p.y, p.N, m_base, n0 + c0, bc >> 3, p.M, p.N, lane);
    __syncwarp();
  }
}

}  // namespace mixq
