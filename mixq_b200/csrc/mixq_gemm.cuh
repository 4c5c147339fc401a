// Kernel-side parameter block shared between mixq_gemm.cu and mixq_api.cu.
#pragma once
#include "ptx.cuh"
#include "rowquant.cuh"

namespace mixq {

enum EpilogueKind : int { EPI_DEQUANT_F16 = 0, EPI_RAW_I32 = 1 };

struct LinearParams {
  CUtensorMap tm_a;    // q_x   int8 [M,K]      box 128 B x 128 rows, SWIZZLE_128B
  CUtensorMap tm_b;    // q_w   int8 [N,K]      box 128 B x BN rows  (W4: uint8 [N,K/2], box 64 B x BN rows)
  CUtensorMap tm_oa;   // activation outliers fp16 [M,n_ind]  box 64 x 128
  CUtensorMap tm_ob;   // weight_cache        fp16 [N,n_ind]  box 64 x BN
  CUtensorMap tm_b2;   // SwiGLU pair: up_proj q_w (rank 1 of the CTA pair stages these rows instead)
  CUtensorMap tm_ob2;  // SwiGLU pair: up_proj weight_cache
  const __half* scale_col2;   // SwiGLU pair: up_proj scale_col
  int pair_swiglu;
  RowQuantArgs rq;     // phase A (fused prologue); rq.q_x == nullptr -> no prologue
  const __half* x_scale;
  const __half* scale_col;
  const __half* bias;
  const __half* outl;  // optional precomputed fp16 [M,N] addend (mixlib.int8FusedDequantize's 5th argument)
  int ld_outl;
  const __half* residual;  // optional fp16 [M, ld_res]: y = fp16(y + residual) (decoder residual stream)
  int ld_res;
  __half* y;
  // tensor-parallel push (row-parallel Linear, reduce-scatter fused into the epilogue): when peer_cols > 0 the output column
  // slice j = n / peer_cols of every tile is stored straight into rank j's receive slot for this rank,
  // y_peer[j][row * peer_cols + (n - j * peer_cols)] — peer memory over NVLink for j != rank — instead of y[row * N + n]
  __half* y_peer[8];
  int peer_cols;
  int peer_bcast;      // > 0: every tile goes, whole, to each of y_peer[0 .. peer_bcast) (fp16 [M,N] receive slots): the one-shot
                       // exchange for small worlds — no second phase, (world - 1) * M * N fp16 over NVLink per rank
  // split-K (1-CTA kernel, M <= 128: a handful of tiles would leave most SMs idle): `splits` CTAs share a tile's K range, the
  // first splits - 1 park their int32 partial in sk_ws [splits - 1][M, N] and bump sk_cnt[tile], the last one adds them up
  int splits;
  int32_t* sk_ws;
  uint32_t* sk_cnt;
  int32_t* y_i32;
  int M, N, K;
  int n_out;           // outlier columns multiplied on the tensor cores (0 = none)
  int act;
  int epilogue;
  int fused_prologue;
  uint32_t* grid_sync;
  int bn;              // 2-CTA kernel: run-time tile width W (multiple of 32, <= 512)
  int nstages;         // 2-CTA kernel: pipeline stages that fit PIPE_BYTES at this W
  int stage_bytes;     // 2-CTA kernel: bytes between consecutive stages (>= k_atoms * (16 KB + W/2 * 128), multiple of 1024)
  int ablate;          // tuning aid (MIXQ_DEBUG_ABLATE; results are garbage): 1 = no MMAs, 2 = no activation loads, 4 = no weight loads, 8 = 16-column epilogue reads, 16 = single outlier pass buffer (8 and 16 keep results exact)
  int k_atoms;         // 2-CTA kernel: 128-byte k-atoms per pipeline stage and TMA op (1, or 2 with 3-D tm_a / tm_b / tm_b2)
  int w4;              // 2-CTA kernel: weights are packed nibbles (tm_b / tm_b2: uint8 [N, K/2], box 64 B x W/2 rows, no swizzle);
                       // the epilogue warps unpack each 64-byte row into the SWIZZLE_128B int8 tile during the mainloop
  int npacked;         // W4: slots of the packed-row ring that follows the nstages main stages in shared memory
  const uint8_t* q_w;  // raw weight pointer + row pitch in bytes (L2 prefetch of the weight stream)
  long long q_w_pitch;
  unsigned long long* trace;  // optional [gridDim.x * 8] globaltimer stamps (mixq_set_trace_buffer), debug/tuning only
};

template <int BN, bool W4>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK_BYTES = 128;                       // one SWIZZLE_128B row: 128 int8 or 64 fp16
  static constexpr int A_BYTES = BM * BK_BYTES;              // 16 KB
  static constexpr int B_BYTES = BN * BK_BYTES;              // 16 / 32 KB
  static constexpr int BP_BYTES = W4 ? BN * 64 : 0;          // packed-nibble landing buffer
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES + BP_BYTES;
  static constexpr int STAGES = W4 ? (BN == 128 ? 5 : 3) : (BN == 128 ? 6 : 4);
  static constexpr int NUM_THREADS = W4 ? 384 : 256;         // W4 adds 4 unpack warps
  static constexpr int EPI_WARPS = 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 512 /*RowQuantSmem*/ +
                                    EPI_WARPS * 32 * 128 /*epilogue staging*/ + 512 /*scale_col of the tile*/;
};

template <int BN, bool W4>
__global__ void mixq_linear_kernel(const __grid_constant__ LinearParams p);

// 2-CTA (cta_group::2) kernel: a CTA pair owns a 256 x W tile (W <= 512); each CTA stages 128 activation rows + W/2 weight
// rows per k-block; the number of pipeline stages follows from W (8 at W = 128 ... 4 at W = 512).
struct Gemm2Cfg {
  static constexpr int A_BYTES = 128 * 128;                  // 16 KB
  static constexpr int PIPE_BYTES = 192 * 1024;              // all stages together
  static constexpr int MAX_STAGES = 8;
  static constexpr int EPI_WARPS = 8;
  static constexpr int NUM_THREADS = 128 + 32 * EPI_WARPS;   // TMA x2, MMA, TMEM-alloc + epilogue warps
  static constexpr int SMEM_BYTES = PIPE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 512 /*RowQuantSmem*/ +
                                    EPI_WARPS * 32 * 128 /*epilogue staging*/ + 1024 /*scale_col of the tile*/;
};
template <bool W4>
__global__ void mixq_linear2_kernel(const __grid_constant__ LinearParams p);

// TMEM plan of the 2-CTA kernel (512 columns), ONE definition for the kernel and for the host-side planner
// (mixq_plan_linear, tests/test_host_cpu.py).  SLOTS int32 accumulators of W columns: two when a pair has several tiles and
// they fit, so that the MMAs of tile i + 1 run while the epilogue drains tile i.  What is left holds the fp32 accumulator of
// the skinny outlier GEMM: all W columns at once when they fit (one pass), otherwise NB = 1 or 2 buffers of R columns that
// the MMA warp refills while the epilogue drains the tile pass by pass (balanced: 352 -> 6 x 64, not 5 x 64 + 32).  Without
// outliers a "pass" is the whole tile and the pass buffers ARE the accumulator slots.
struct TmemPlan {
  int slots;    // int32 accumulator slots (1 or 2)
  int passes;   // P: epilogue passes per tile
  int pass_cols;// R: accumulator columns per pass (both CTA halves together)
  int buffers;  // NB: pass buffers / barrier pairs in use (1 or 2)
  __host__ __device__ int columns(int W, bool has_o) const { return slots * W + (has_o ? buffers * pass_cols : 0); }
};
__host__ __device__ inline TmemPlan plan_tmem(int W, bool has_o, int tiles_of_pair, bool single_buffer) {
  TmemPlan t;
  t.slots = (tiles_of_pair > 1 && 2 * W + (has_o ? 64 : 0) <= 512) ? 2 : 1;
  t.passes = 1;
  t.pass_cols = W;
  t.buffers = has_o ? 1 : t.slots;
  if (has_o) {
    const int spare = 512 - t.slots * W;             // the host guarantees W <= 448 with outliers: spare >= 64
    if (spare < W) {
      if (spare >= 128 && !single_buffer) { t.buffers = 2; t.pass_cols = (spare >> 1) & ~31; }
      else { t.buffers = 1; t.pass_cols = spare & ~31; }
      t.passes = (W + t.pass_cols - 1) / t.pass_cols;
      t.pass_cols = ((W + t.passes - 1) / t.passes + 31) & ~31;
    }
  }
  return t;
}

}  // namespace mixq
