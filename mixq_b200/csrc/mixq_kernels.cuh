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

}  // namespace mixq
