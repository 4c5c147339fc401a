#pragma once
#include "rowquant.cuh"

namespace mixq {

__global__ void rowquant_kernel(const RowQuantArgs a);
__global__ void extract_outliers_kernel(const int32_t* ind, int n_ind, __half* x, __half* out, int ld_out, int M,
                                        int K);
__global__ void dequant_i32_kernel(const int32_t* acc, const __half* x_scale, const __half* scale_col,
                                   const __half* outl, int ld_outl, __half* y, int M, int N, int act);
__global__ void unpack_int4_cols_kernel(const uint8_t* q_w_packed, const int32_t* ind, int n_ind, __half* out,
                                        int ld_out, int N, int K);
__global__ void gather_weight_cols_kernel(const void* q_w, const __half* scale_col, const int32_t* ind, int n_ind,
                                          __half* wc, int ld_wc, int col0, int N, int K, int bit);
__global__ void compact_cols_kernel(uint8_t* col_over, int K, int32_t* ind_out, int max_new, int32_t* n_new);
cudaError_t launch_rope_attn_decode(const __half* qkv, __half* k_cache, __half* v_cache, int cache_cap, int past_len,
                                    __half* out, int M, int H, int Hkv, int D, float theta, bool pdl, cudaStream_t st,
                                    const RowQuantArgs* rq = nullptr);   // rq: fold the next MixLinear's activation prologue in
template <int D>
__global__ void rope_attn_decode_quant_kernel(const __half* qkv, __half* k_cache, __half* v_cache, int cache_cap, int past_len,
                                              int H, int Hkv, float theta, float scale, const RowQuantArgs rq);
__global__ void mul_inplace_kernel(__half2* a, const __half2* b, long long n2);
// QUIK (mixquant/modules/qlinear.py): asymmetric per-token activation quantisation + the zero-point correction addend
__global__ void quik_quantize_kernel(const __half* x, const int64_t* int_idx, int n_int, const int64_t* fp_idx, int n_fp, int bits,
                                     int8_t* q, __half* meta, __half* fp_x, int M, int K);
__global__ void quik_addend_kernel(const __half* meta, const __half* reduced_w, const __half* fp_result, int ld_fp, __half* out,
                                   int M, int N, int bits);

constexpr int kMaxPeers = 8;
struct AllReduceArgs {
  const __half* partial[kMaxPeers][2];   // [rank][buffer]: every rank's two partial buffers as mapped into THIS process
  __half* result[kMaxPeers][2];          // two-shot only: [rank][buffer] result buffers as mapped into this process (else nullptr)
  uint32_t* flags[kMaxPeers];            // [rank]: that rank's 2 * kMaxPeers flag words, as mapped into this process
  uint32_t* epoch;                       // local: exchanges finished so far
  uint32_t* done;                        // local: blocks finished in the current launch
  const __half* residual;                // local [n] or nullptr
  __half* out;                           // local [n]
  long long n;                           // elements, multiple of 8
  int world, rank, buf;
  unsigned long long timeout_ns;         // a peer that stays silent this long is reported (printf + trap); mixq_set_peer_timeout_ms
};
__global__ void allreduce_residual_kernel(AllReduceArgs a);

// The same exchange through an NVLink-SHARP multicast mapping (NVSwitch reduces and broadcasts: multimem.ld_reduce / multimem.st).
struct McAllReduceArgs {
  uint8_t* mc;                 // multicast address of the symmetric allocation (every rank's copy at once)
  uint8_t* local;              // this rank's copy of the same allocation
  unsigned long long partial_off[2], result_off[2], flags_off;   // byte offsets inside the allocation (identical on all ranks)
  uint32_t* epoch;             // local: exchanges finished so far
  uint32_t* done;              // local: blocks finished in the current launch
  const __half* residual;      // local [n] or nullptr
  long long n;                 // elements, multiple of 8 * world
  int world, rank, buf;
  unsigned long long timeout_ns;
};
__global__ void allreduce_multicast_kernel(McAllReduceArgs a);


// Second half of the fused row-parallel exchange.  The GEMM epilogues have PUSHED every rank's partial of column slice j into
// rank j's receive slots (LinearParams::y_peer); this kernel shakes hands, reduces this rank's slice from its local slots
// (fp32, rank order, one rounding to fp16), adds the residual slice as a separate fp16 op and broadcasts the slice into every
// rank's result buffer — through the switch (multimem.st) when a multicast mapping exists, else with one store per peer.
struct XchgFinishArgs {
  const __half* recv;          // local: [world][M, Ns] receive slots of this exchange (slot s = rank s's partial of MY slice)
  __half* result[kMaxPeers];   // every rank's result buffer [M, N] of this exchange as mapped here (own rank: local)
  __half* mc_result;           // multicast address of the result buffer, or nullptr
  uint32_t* flags[kMaxPeers];  // every rank's two handshake counters as mapped here (own rank: local)
  uint32_t* mc_flags;          // multicast address of the counters, or nullptr
  uint32_t* epoch;             // local: exchanges finished so far
  uint32_t* done;              // local: blocks finished in the current launch
  const __half* residual;      // local [M, N] or nullptr
  int M, N, world, rank;
  int one_shot;                // 1: recv holds every rank's FULL partial [world][M, N] (broadcast push): reduce all of it, no second phase
  unsigned long long timeout_ns;
  int ablate;                  // tuning aid (MIXQ_DEBUG_XF; results are garbage): 1 = no second handshake, 2 = no handshakes, 4 = no data
  unsigned long long* trace;   // tuning aid: 8 globaltimer stamps of block 0 (mixq_set_trace_buffer) or nullptr
};
__global__ void exchange_finish_kernel(XchgFinishArgs a);
// The same second half WITHOUT flags or fences: the data is its own signal.  Receive slots and result buffers hold a sentinel
// (fp16 0xFFFF: a NaN payload no arithmetic produces — hardware NaNs are 0x7FFF) until a peer's store lands; a reader spins on
// the vector it needs until no half of it is the sentinel, and whoever consumes a vector puts the sentinel back.
struct XchgPollArgs {
  __half* recv;                // local: [world][M, Ns] receive slots of this exchange (one_shot: [world][M, N])
  __half* result[kMaxPeers];   // every rank's result buffer [M, N] of this exchange as mapped here (own rank: local)
  __half* mc_result;           // multicast address of the result buffer, or nullptr
  __half* reset;               // local: the OTHER result buffer (last read, as the residual, by this launch): sentinel-filled here
  const __half* residual;      // local [M, N] or nullptr
  int M, N, world, rank;
  int one_shot;
  unsigned long long timeout_ns;
  unsigned long long* trace;
};
__global__ void exchange_finish_poll_kernel(XchgPollArgs a);
// + RMSNorm / outlier gather / quantisation of every row for the Linear that follows (one CTA per token row)
cudaError_t launch_exchange_finish_rowquant(const XchgPollArgs& a, const RowQuantArgs& rq, int grid, bool pdl, cudaStream_t st);
// tuning aid: `iters` flag round trips between two ranks inside one launch; *out_ns = elapsed nanoseconds (rank 0)
__global__ void pingpong_kernel(uint32_t* mine, uint32_t* peer, uint32_t* mc, int iters, int rank, unsigned long long* out_ns);

}  // namespace mixq
