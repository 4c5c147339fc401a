/*
 * libmixq_sm100 — C ABI of the B200-native MixLinear hot path.
 *
 * This is the drop-in boundary for the reference's native extension module `mixlib`
 * (github.com/Qcompiler/QComplier, quantkernel; NOT vendored in /root/reference — only its call
 * sites are).  Each entry point below names the `mixlib.*` call it replaces, cited by
 * file:line under /root/reference.  `mixq_b200/mixlib.py` is the ctypes binding a maintainer
 * would drop in as the `mixlib` module.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless it says "host";
 *   - the caller owns every buffer, nothing is allocated inside; `stream` is a cudaStream_t
 *     (NULL = legacy default stream); calls are asynchronous on that stream;
 *   - return 0 on success, a cudaError_t (>0) or a MIXQ_E* code (<0) otherwise;
 *     mixq_last_error() returns a human-readable message for the calling thread;
 *   - matrices are row-major and dense unless an explicit leading dimension `ld*` (in elements)
 *     is given; fp16 = IEEE binary16; activations x[M,K], weights q_w[N,K] (K contiguous);
 *   - M >= 1; K % 16 == 0; N % 8 == 0 (TMA / 16-byte vector alignment); base pointers 16-byte aligned;
 *     the activation prologue (FindRowScale, RMSNorm, fused phase A) needs K <= 32768.
 *   - thread-compatible, not thread-safe on the same buffers (the reference shares one
 *     MixLibCache across all layers and relies on single-stream order: Cache.py:5-25).
 */
#ifndef MIXQ_H_
#define MIXQ_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIXQ_EINVAL (-1)   /* bad argument (shape/alignment/bit width) */
#define MIXQ_EARCH (-2)    /* device is not sm_100 */
#define MIXQ_EDRIVER (-3)  /* cuTensorMapEncodeTiled unavailable / failed */

#define MIXQ_ACT_NONE 0
#define MIXQ_ACT_SILU 1

const char* mixq_last_error(void);
int mixq_version(void);
/* Number of kernels this library has launched on the calling process so far (bench.py's gpu_launches). */
unsigned long long mixq_launch_count(void);
/* Tuning/testing override of the GEMM tile width for calls that do not carry their own tile_n:
 * 0 = heuristic (default); a multiple of 32 up to 512 for the 2-CTA kernel (M > 128, bit 8; capped at 448 with outlier
 * columns), of which the 1-CTA kernel honours 128 and 256 only.  mixq_plan_linear reports what a shape resolves to. */
int mixq_set_tile_n(int tile_n);
/* How a MixLinear GEMM of this shape would be launched (pure host arithmetic; sms <= 0 = ask the current device):
 * which kernel, tile width, k-atoms per TMA op and pipeline stage, pipeline depth, work split, TMEM plan. */
typedef struct mixq_linear_plan {
  int two_cta;        /* 1: mixq_linear2_kernel (CTA pairs, 256 x tile_w tiles); 0: mixq_linear_kernel (128 x tile_w) */
  int tile_w;         /* output columns of a tile */
  int k_atoms;        /* 128-byte k-atoms per TMA op / stage */
  int stage_bytes, nstages;
  int tiles, units, tiles_per_unit;   /* units = CTA pairs (two_cta) or CTAs; tiles_per_unit = the busiest unit */
  int acc_slots;      /* int32 accumulators resident in TMEM (2: tile i + 1's MMAs overlap tile i's epilogue) */
  int passes, pass_cols, pass_buffers;   /* epilogue passes per tile and the fp32 outlier accumulator buffers */
  int tmem_cols;      /* TMEM columns in use (<= 512) */
} mixq_linear_plan;
int mixq_plan_linear(int M, int N, int K, int bit, int n_ind, int swiglu_pair, int tile_n, int sms, mixq_linear_plan* out);
/* How many CTAs per tile a launch with a split-K workspace of `splitk_ws_bytes` would use along K (1 = no split; see
 * mixq_linear_args.splitk_ws).  Host-only, no launch; sms <= 0 = the current device.  < 0 on bad arguments. */
int mixq_plan_split_k(int M, int N, int K, int bit, int n_ind, int sms, long long splitk_ws_bytes);
/* Programmatic dependent launch (default on; MIXQ_PDL=0 in the environment or on = 0 turns it off): every kernel of the
 * library is launched so that its CTAs may take an SM as soon as the previous kernel's CTA there has exited, set up, and
 * prefetch constants (the quantised weights) while the previous kernel drains; each kernel waits for its predecessors
 * (griddepcontrol.wait) before touching any activation tensor.  Results are identical either way. */
int mixq_set_pdl(int on);
/* How the launches that contain a grid barrier (mixq_linear_fused with its activation prologue) guarantee that all their CTAs
 * are resident at once (environment MIXQ_GRID_BARRIER = pdl | coop | split):
 *   0 "pdl"   default: programmatic dependent launch, one CTA per SM; co-residency holds because the process owns the GPU
 *             (the previous kernel's CTAs leave without waiting for anybody).  The library checks with
 *             cudaOccupancyMaxActiveClusters that the grid fits the device and refuses the launch otherwise.
 *   1 "coop"  cudaLaunchAttributeCooperative on those launches (no PDL overlap for them): the driver enforces co-residency.
 *   2 "split" two launches — activation prologue (row quantiser), then the GEMM with skip_prologue — and no grid barrier:
 *             for processes that share the GPU with other streams, MPS clients or profilers.  Results are bit-identical. */
int mixq_set_grid_barrier_mode(int mode);

/* Tuning aid: device buffer of 148*8 uint64 that every subsequent mixq_linear_fused / GEMM launch fills with
 * %globaltimer stamps per CTA (start, prologue done, grid barrier passed, first MMA, last MMA, epilogue done).
 * NULL (default) disables it. */
int mixq_set_trace_buffer(void* buf);

/* ---- mixlib.FindRowScale(x, x_scale, M, K, bit) -> q_x        (linear.py:190-193, :221)
 * x_scale[m] = fp16(max_k |x[m,k]| / (2^(bit-1)-1));  q_x[m,k] = clamp(rint(x/x_scale)).
 * q_x is int8 [M,K] for bit 8 AND bit 4 (bit 4: values in [-7,7], one per byte — the packed form is
 * opaque to the reference's Python, and Blackwell has no int4 MMA to feed it to). */
int mixq_find_row_scale(const void* x, void* x_scale, void* q_x, int M, int K, int bit, void* stream);

/* Same, plus the scan the reference does on the host side of the call (linear.py:201 and FindOutliers,
 * linear.py:157-161): *over_flag |= 1 when any x_scale[m] > fp16(sigma/qmax); col_over[c] = 1 for every column
 * with some |x[m,c]| > sigma.  Either output may be NULL. */
int mixq_find_row_scale_scan(const void* x, void* x_scale, void* q_x, int M, int K, int bit, float sigma,
                             uint8_t* col_over, uint32_t* over_flag, void* stream);

/* ---- mixlib.ExtractOutliersAndSetToZeros(ind, x) -> out[M,n]   (linear.py:189, :205)
 * out[m,j] = x[m,ind[j]];  x[m,ind[j]] = 0 IN PLACE.  ld_out >= n_ind. */
int mixq_extract_outliers_and_set_to_zeros(const int32_t* ind, int n_ind, void* x, void* out, int ld_out, int M,
                                           int K, void* stream);

/* ---- mixlib.int8FusedDequantize / int8FusedDequantizeSilu      (linear.py:251-256, :268-273, :337-351)
 * y[m,n] = act( fp16( (float(sum_k q_x[m,k]*q_w[n,k]) * x_scale[m]) * scale_col[n] + outl[m,n] ) )
 * outl may be NULL (== the reference passing cache.zeros).  ld_outl in elements.
 * tcgen05 kind::i8 GEMM on TMA-staged tiles, fused epilogue. */
int mixq_int8_fused_dequantize(const void* q_x, const void* q_w, const void* x_scale, const void* scale_col,
                               const void* outl, int ld_outl, void* y, int M, int N, int K, int act,
                               void* stream);

/* ---- mixlib.int4FusedDequantize / ...Silu                      (linear.py:259-265, :278-283, :360-366)
 * q_w is the packed-nibble weight uint8 [N,K/2] (low nibble = even k, two's complement, linear.py:14-18);
 * q_x is int8 [M,K] as produced by mixq_find_row_scale(bit=4).  K is the LOGICAL K (the reference passes K/2). */
int mixq_int4_fused_dequantize(const void* q_x, const void* q_w_packed, const void* x_scale,
                               const void* scale_col, const void* outl, int ld_outl, void* y, int M, int N,
                               int K, int act, void* stream);

/* ---- mixlib.gemm(q_x, q_w, M, N, K) -> int32 [M,N]             (linear.py:235, :321; arch==9 split path) */
int mixq_gemm_i8(const void* q_x, const void* q_w, int32_t* y, int M, int N, int K, void* stream);

/* ---- mixlib.dequantizeInt8 / dequantizeInt8Silu                (linear.py:238, :241, :324, :327) */
int mixq_dequantize_int8(const int32_t* acc, const void* x_scale, const void* scale_col, const void* outl,
                         int ld_outl, void* y, int M, int N, int act, void* stream);

/* ---- mixlib.unpack_int4_to_fp16(q_w, ind) -> fp16 [N, n]       (linear.py:20-22)
 * out[n,j] = sign-extended nibble of column ind[j] (un-scaled). */
int mixq_unpack_int4_to_fp16(const void* q_w_packed, const int32_t* ind, int n_ind, void* out, int ld_out, int N,
                             int K, void* stream);

/* ---- mixlib.layernorm_forward_cuda(x, w, out, eps)             (fused/norm.py:21) — RMSNorm */
int mixq_rmsnorm(const void* x, const void* w, void* out, float eps, int M, int K, void* stream);

/* ---- mixlib.layernorm_forward_cuda_extract_outliers[_int4](x, w, out, eps, ind, x_scale) -> (outl, q_x)
 *      (fused/norm.py:25-33).  out has the ind columns zeroed (same as the unfused path, linear.py:189). */
int mixq_rmsnorm_extract_outliers(const void* x, const void* w, void* out, float eps, const int32_t* ind,
                                  int n_ind, void* x_scale, void* act_out, int ld_ao, void* q_x, int M, int K,
                                  int bit, void* stream);

/* ---- weight_cache columns: q_weight[:,ind].half() * scale_col.T  (linear.py:207, :209-210, :305-308)
 * wc[n, col0 + j] = fp16(q_w[n, ind[j]]) * scale_col[n]   (fp16 multiply, like the reference) */
int mixq_gather_weight_columns(const void* q_w, const void* scale_col, const int32_t* ind, int n_ind, void* wc,
                               int ld_wc, int col0, int N, int K, int bit, void* stream);

/* ---- FindOutliers (linear.py:157-161) on the device: compact col_over[K] (bytes set by the scan) into
 * ascending column ids appended at ind[n_ind...] ; *n_new (device int32) receives the count; col_over is
 * cleared.  At most max_new ids are written. */
int mixq_compact_outlier_columns(uint8_t* col_over, int K, int32_t* ind_out, int max_new, int32_t* n_new,
                                 void* stream);

/* ---- the north-star entry: steady-state MixLinear_GEMM.forward as ONE launch (linear.py:165-289, :291-376).
 * Phase A (every CTA): [RMSNorm ->] gather+zero outlier columns -> row absmax -> x_scale -> int8 q_x,
 * outlier scan against sigma; one grid barrier; phase B: persistent warp-specialised tcgen05 GEMM:
 * int8 x int8 -> int32 in TMEM over K, fp16 x fp16 -> fp32 in TMEM over the outlier columns, epilogue
 * y = act(fp16((acc*x_scale[m])*scale_col[n] + acc_outl + bias[n])). */
typedef struct mixq_linear_args {
  /* activations */
  void* x;                 /* fp16 [M,K]; outlier columns zeroed in place unless norm_weight is set */
  const void* norm_weight; /* optional fp16 [K]: x is un-normed, norm_out receives RMSNorm(x)*w */
  void* norm_out;          /* fp16 [M,K] RMSNorm(x)*w with the ind columns zeroed, or NULL when nobody reads it */
  float eps;
  int M, N, K;
  /* weights */
  const void* q_weight;    /* int8 [N,K] (bit 8) | uint8 [N,K/2] packed nibbles (bit 4) */
  const void* scale_col;   /* fp16 [N] */
  const void* bias;        /* fp16 [N] or NULL; y = fp16(y + bias) as a second rounding (linear.py:284-285) */
  int bit;                 /* 8 or 4 */
  /* SwiGLU pair (fused/mlp.py:61-64): when q_weight_up is set, q_weight / scale_col / weight_cache are gate_proj's and
   * these are up_proj's (same N, K, ind, ld_wc; bit 8, M > 128, no bias); y[M,N] = fp16( fp16(silu(gate)) * fp16(up) ),
   * i.e. up_proj(x), gate_proj.forward_without_preconditionFusedSilu(x) and `gate *= up` in ONE launch. */
  const void* q_weight_up;
  const void* scale_col_up;
  const void* weight_cache_up;
  /* fp16 outlier path */
  const int32_t* ind;      /* [n_ind] */
  int n_ind;
  const void* weight_cache;/* fp16 [N, ld_wc], first n_ind columns valid */
  int ld_wc;               /* multiple of 8 */
  /* caller-owned scratch (what MixLibCache carries: Cache.py:8,17 + linear.py:190) */
  void* q_x;               /* int8 [M,K] */
  void* x_scale;           /* fp16 [>=M] */
  void* act_outliers;      /* fp16 [M, ld_ao] */
  int ld_ao;               /* multiple of 8 */
  /* outlier scan */
  float sigma;             /* threshold (Cache.py:6,12: 6) */
  uint8_t* col_over;       /* [K] bytes or NULL */
  uint32_t* over_flag;     /* or NULL; |= 1 when max(x_scale) > fp16(sigma/qmax) (linear.py:201) */
  /* output */
  const void* residual;    /* fp16 [M, ld_res] or NULL: y = fp16(y + residual) — the decoder layer's residual add */
  int ld_res;
  void* y;                 /* fp16 [M,N] */
  int act;                 /* MIXQ_ACT_* */
  /* control */
  int skip_prologue;       /* 1: q_x / x_scale / act_outliers already valid (gate_proj, linear.py:291-376) */
  uint32_t* grid_sync;     /* one zero-initialised u32 in device memory, reused across launches */
  int tile_n;              /* 0 = auto; else a multiple of 32 up to 512 (the 1-CTA kernel, M <= 128 or W4, honours 128 / 256 only) */
  /* tensor-parallel push: the reduce-scatter half of the row-parallel exchange fused into the GEMM epilogue.  peer_cols =
   * N / world > 0: output column slice j = n / peer_cols goes to y_peer[j] (fp16 [M, peer_cols]: rank j's receive slot for THIS
   * rank, a peer-mapped pointer for j != rank) instead of y; no bias / residual / addend / SwiGLU pair; peer_cols % 128 == 0
   * (% 256 for M <= 128).  mixq_exchange_finish then reduces and broadcasts.  peer_cols = 0: off. */
  void* y_peer[8];
  int peer_cols;
  /* one-shot variant for small worlds: peer_bcast = number of destinations > 0 (peer_cols = 0): every tile is stored, whole, into
   * each of y_peer[0 .. peer_bcast) (fp16 [M,N] receive slots, one per rank incl. this one); mixq_exchange_finish(one_shot = 1)
   * then reduces all of them locally with a single handshake. */
  int peer_bcast;
  /* split-K workspace (optional; M <= 128 only): with a handful of 128-row tiles most SMs would idle, so up to 4 CTAs share a
   * tile's K range and meet through this buffer — int32 partial sums, added exactly, so the result is bit-identical to the
   * unsplit launch.  Caller-owned device memory, ZERO-FILLED ONCE by the caller (the first 4096 bytes are per-tile counters
   * that every launch leaves at zero again), then one 64 KB block per (128 x 128 tile, non-final split):
   * 4096 + ceil(M/128) * ceil(N/128) * 3 * 65536 bytes allow the full split, less caps the split factor; one launch at a time
   * per workspace.  NULL = never split. */
  void* splitk_ws;
  long long splitk_ws_bytes;
} mixq_linear_args;

int mixq_linear_fused(const mixq_linear_args* args /* host */, void* stream);

/* ---- decode-harness glue, outside the quantised path (the reference calls flash-attn: fused/attn.py:239-258).
 * RoPE(q,k at position past_len, HF rotate_half) + single-query attention over an optional KV cache
 * [M, Hkv, cache_cap, D] (new k/v appended at past_len) -> out[M, H*D].  qkv row = [H*D | Hkv*D | Hkv*D]. */
int mixq_rope_attention_decode(const void* qkv, void* k_cache, void* v_cache, int cache_cap, int past_len, void* out,
                               int M, int H, int Hkv, int D, float theta, void* stream);

/* The same attention with the activation prologue of the Linear that consumes it (o_proj, called in unfused mode at
 * fused/attn.py:263 => linear.py:187-193 ExtractOutliersAndSetToZeros + FindRowScale) folded in: one CTA owns one token row,
 * so the row abs-max is known as soon as the row's heads are done.  Writes q_x int8 [M, H*D], x_scale fp16 [M] and
 * act_outliers fp16 [M, ld_ao] (columns ind[0..n_ind)); `out` (fp16 [M, H*D], outlier columns zeroed like the reference
 * leaves its tensor) is optional — NULL keeps no fp16 copy.  The consumer then runs mixq_linear_fused(skip_prologue = 1):
 * the reference's "fused" call mode (fused/norm.py:24-33 is the other producer of this kind). */
int mixq_rope_attention_decode_quant(const void* qkv, void* k_cache, void* v_cache, int cache_cap, int past_len, void* out,
                                     int M, int H, int Hkv, int D, float theta, const int32_t* ind, int n_ind,
                                     void* act_outliers, int ld_ao, void* q_x, void* x_scale, int bit, void* stream);

/* ---- QUIK MixedQLinear (mixquant/modules/qlinear.py:41-211; `quik` is an un-vendored third-party extension: parity
 * unpinned, see oracle/quik_oracle.py).
 * quik.asymmetric.quantize(x, int_indices, fp_indices, bits) (qlinear.py:117-120): per token row zero = min, scale =
 * (max - min) / (2^bits - 1) over the int columns -> meta fp16 [2, M] (row 0 scales, row 1 zeros); q int8 [M, n_int] =
 * rn((x - zero) / scale) - 2^(bits-1), one value per byte; fp_x fp16 [M, n_fp] = x[:, fp_indices].  Indices are int64
 * (the reference registers them as torch.long buffers, qlinear.py:66-69). */
int mixq_quik_quantize(const void* x, const int64_t* int_indices, int n_int, const int64_t* fp_indices, int n_fp, int bits,
                       void* q, void* meta, void* fp_x, int M, int K, void* stream);
/* The addend of quik.asymmetric.dequantize (qlinear.py:149-150): out fp16 [M,N] = (zero[m] + 2^(bits-1) scale[m]) * reduced_w[n]
 * + fp_result[m,n] (fp_result may be NULL); the int GEMM (mixq_int4_fused_dequantize / mixq_int8_fused_dequantize with
 * x_scale = meta row 0 and outl = this addend) completes y = acc * scale[m] * weights_scales[n] + addend. */
int mixq_quik_addend(const void* meta, const void* reduced_w, const void* fp_result, int ld_fp, void* out, int M, int N,
                     int bits, void* stream);

/* Tuning aid: `iters` flag round trips between two ranks inside ONE launch (rank 0 writes rank 1's word, rank 1 answers);
 * *out_ns (device uint64) = elapsed nanoseconds.  mine / peer: this rank's word and the other rank's word as mapped here;
 * mc: multicast address of the word (then both sides use multimem.red) or NULL. */
/* Re-read the MIXQ_DEBUG_* tuning variables (they are otherwise read once, when the library loads — never per launch). */
void mixq_reload_debug_env(void);
int mixq_debug_pingpong(void* mine, void* peer, void* mc, int iters, int rank, void* out_ns, void* stream);

/* elementwise gate *= up (mlp.py:64) kept for the decode harness */
int mixq_mul_inplace(void* a, const void* b, long long n, void* stream);

/* ---- The exchange step of the row-parallel Linears (o_proj, down_proj) under tensor parallelism, over NVLink peer memory.
 * The reference has no multi-GPU code (models/base.py:196-225 is accelerate placement); this replaces ncclAllReduce + the
 * decoder's residual add by ONE kernel per rank: out = fp16( fp16(sum over ranks of partial, fp32, rank order) + residual ).
 * Set-up (once per process): every rank allocates two partial buffers (two-shot: and two result buffers) and 16 flag words with mixq_peer_alloc,
 * publishes their handles (mixq_ipc_get_handle, 64 bytes) to the other ranks of the node, and maps theirs
 * (mixq_ipc_open_handle).  Per exchange: the rank's Linear writes its partial into ITS buffer `buf`, then every rank calls
 * mixq_allreduce_residual with the same n and buf.  Callers ALTERNATE buf = 0, 1, 0, 1, ... over successive exchanges (a rank
 * may overwrite a buffer only after the following exchange, which proves every peer has finished reading it); a CUDA graph
 * that is replayed must therefore hold an even number of exchanges.  `epoch` / `done`: two zero-initialised local uint32 (the
 * kernel keeps the exchange count on the device, so the call is graph replayable). */
typedef struct mixq_allreduce_args {
  const void* partial0[8]; /* [rank]: that rank's partial buffer 0, as mapped in this process (own rank: the local pointer) */
  const void* partial1[8]; /* [rank]: buffer index 1 */
  void* flags[8];          /* [rank]: that rank's 16 flag words (uint32), as mapped in this process */
  void* result0[8];        /* two-shot (world >= 4 pays off): [rank]: that rank's result buffer 0 / 1, as mapped in this process; */
  void* result1[8];        /*   every rank reduces 1/world of the vector and pushes it to all; out = own result buffer `buf`. NULL = one-shot */
  void* epoch;             /* local uint32 */
  void* done;              /* local uint32 */
  const void* residual;    /* local fp16 [n] or NULL */
  void* out;               /* local fp16 [n] */
  long long n;             /* elements, multiple of 8 */
  int world, rank;         /* 2 <= world <= 8 */
  int buf;                 /* 0 or 1: which partial buffer holds this exchange */
} mixq_allreduce_args;
int mixq_peer_alloc(unsigned long long bytes, void** ptr);
int mixq_peer_free(void* ptr);
int mixq_ipc_get_handle(const void* ptr, void* handle64);
int mixq_ipc_open_handle(const void* handle64, void** ptr);
int mixq_ipc_close_handle(void* ptr);
int mixq_allreduce_residual(const mixq_allreduce_args* a, void* stream);
/* The same exchange through an NVLink-SHARP (NVLS) multicast mapping: ONE symmetric allocation per rank (e.g. from
 * torch.distributed._symmetric_memory: plumbing) holding partial 0 / 1, result 0 / 1 and 8 flag bytes at the same offsets on
 * every rank, `local` = this rank's copy, `mc` = the multicast address of all copies.  Rank r reduces elements
 * [r n / world, (r + 1) n / world) with multimem.ld_reduce (the switch adds the copies, fp32 accumulation), adds the residual
 * slice as a separate fp16 rounding and multimem.st's the slice into every rank's result buffer; both handshakes are one
 * multimem.red each.  The row-parallel Linear writes its partial into local + partial_off[buf]; the result of the exchange is
 * local + result_off[buf] (valid until the exchange after the next one).  buf alternates 0, 1 as above; n % (8 * world) == 0. */
typedef struct mixq_mc_allreduce_args {
  void* mc;
  void* local;
  unsigned long long partial_off[2];
  unsigned long long result_off[2];
  unsigned long long flags_off;   /* 8 zero-initialised bytes (two uint32 counters), 16-byte aligned */
  void* epoch;                    /* local uint32, zero-initialised */
  void* done;                     /* local uint32, zero-initialised */
  const void* residual;           /* local fp16 [n] or NULL */
  long long n;
  int world, rank, buf;
} mixq_mc_allreduce_args;
int mixq_allreduce_multicast(const mixq_mc_allreduce_args* a, void* stream);

/* The FUSED row-parallel exchange: (1) the row-parallel MixLinear's epilogue pushes column slice j of its fp16 partial straight
 * into rank j's receive slot (mixq_linear_args.y_peer / peer_cols: NVLink stores from the epilogue warps, overlapped with the
 * GEMM's own tail); (2) mixq_exchange_finish shakes hands, reduces this rank's slice from its LOCAL slots (fp32, rank order, one
 * rounding to fp16), adds the residual slice as a separate fp16 rounding and broadcasts the slice into every rank's result
 * buffer (multimem.st through the switch when mc_result is set, else one store per peer), and shakes hands again.  Per rank
 * 2 (world-1)/world n fp16 cross NVLink instead of (world-1) n, and only the small second half is exposed.
 *   recv      : local fp16 [world][M, N/world] receive slots of THIS exchange (slot s = rank s's partial of my slice)
 *               (one_shot: [world][M, N], every rank's full partial)
 *   result[r] : rank r's result buffer fp16 [M, N] of this exchange as mapped here (own rank: the local pointer)
 *   flags[r]  : rank r's two zero-initialised uint32 handshake counters as mapped here; mc_flags / mc_result: multicast
 *               addresses of the same buffers, or NULL
 * Callers alternate between two (recv, result) buffer sets exactly as for mixq_allreduce_residual. */
typedef struct mixq_exchange_finish_args {
  const void* recv;
  void* result[8];
  void* mc_result;
  void* flags[8];
  void* mc_flags;
  void* epoch;             /* local uint32, zero-initialised */
  void* done;              /* local uint32, zero-initialised */
  const void* residual;    /* local fp16 [M, N] or NULL */
  int M, N, world, rank;
  int one_shot;            /* 1: recv = [world][M, N] FULL partials (peer_bcast push): reduce everything locally, one handshake,
                            *    result[rank] only (no broadcast, result / mc_result of the peers unused) */
} mixq_exchange_finish_args;
int mixq_exchange_finish(const mixq_exchange_finish_args* a, void* stream);

/* The same second half with NO flags and NO fences — the data is its own signal (measured on NVSwitch: a flag takes 2.8 us one
 * way and a system-scope release after stores to peers 6-9 us; the flag protocol above pays two of each per exchange).  Receive
 * slots and result buffers hold the sentinel fp16 0xFFFF (a NaN payload no arithmetic produces) until a peer's store lands:
 * a reader spins on the 16-byte vector it needs until none of its halves is the sentinel, and whoever consumes a vector puts
 * the sentinel back.  The caller fills recv / result buffers with 0xFFFF once, alternates two buffer sets, and passes as
 * `reset` the local result buffer of the PREVIOUS exchange (dead after this launch has read it as the residual).
 * Ordering needs no handshake: a peer pushes into this rank's slots of exchange e only after finishing exchange e - 1, which
 * needed this rank's slice of e - 1, produced after exchange e - 2 released those slots (stream order). */
typedef struct mixq_exchange_poll_args {
  void* recv;              /* local fp16 [world][M, N/world] (one_shot: [world][M, N]) receive slots of this exchange */
  void* result[8];         /* rank r's result buffer of this exchange as mapped here (one_shot: only [rank] is used) */
  void* mc_result;         /* multicast address of the result buffer or NULL */
  void* reset;             /* local: the other result buffer, re-armed (sentinel-filled) by this launch; or NULL */
  const void* residual;    /* local fp16 [M, N] or NULL (may be the `reset` buffer) */
  int M, N, world, rank;
  int one_shot;
} mixq_exchange_poll_args;
int mixq_exchange_finish_poll(const mixq_exchange_poll_args* a, void* stream);

/* The same exchange + the activation prologue of the MixLinear that consumes its result (fused/norm.py:24-33 in front of W_pack /
 * up_proj+gate_proj): after the exchange every rank holds whole rows of the residual stream, and every rank would run the same
 * RMSNorm -> ExtractOutliers -> FindRowScale -> quantise over them at the head of its next launch.  One CTA owns one token row
 * here, so the row is normalised and quantised as soon as its last slice has landed: writes q_x int8 [M, N], x_scale fp16 [M],
 * act_outliers fp16 [M, ld_ao] (normed values of columns ind[0..n_ind)) besides the result buffer.  The consumer runs
 * mixq_linear_fused(skip_prologue = 1) — the reference's "fused" call mode.  Bit-identical to finish + separate prologue. */
int mixq_exchange_finish_poll_quant(const mixq_exchange_poll_args* a, const void* norm_weight, float eps, const int32_t* ind,
                                    int n_ind, void* act_outliers, int ld_ao, void* q_x, void* x_scale, int bit, void* stream);

/* How long (milliseconds, default 120 000; environment MIXQ_PEER_TIMEOUT_MS) an exchange waits for a silent peer before the
 * kernel reports the stall (device printf + trap => a sticky CUDA error on the host).  Ranks drift apart by seconds around
 * host-synchronising phases (outlier discovery, graph capture, rank-0-only work): callers should still put a process-group
 * barrier between such a phase and the next exchange. */
int mixq_set_peer_timeout_ms(long long ms);

#ifdef __cplusplus
}
#endif
#endif /* MIXQ_H_ */
