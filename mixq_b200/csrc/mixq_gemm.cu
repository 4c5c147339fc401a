// The MixLinear hot-path kernel for sm_100a.
//
//   y[M,N] = act( fp16( (int32(q_x . q_w^T) * x_scale[m]) * scale_col[n]          <- tcgen05 kind::i8, TMEM s32
//                       + act_outliers[M,n_out] . weight_cache[N,n_out]^T            <- tcgen05 kind::f16, TMEM f32
//                       + outl[m,n] + bias[n] ) )
//
// One persistent CTA per SM, warp-specialised:
//   warp 0 / 3  TMA producers: activations / weights (cp.async.bulk.tensor, SWIZZLE_128B boxes, mbarrier complete_tx)
//   warp 1      MMA issuer     (one lane issues tcgen05.mma; tcgen05.commit frees smem stages / publishes TMEM)
//   warp 2      TMEM allocator (512 columns; double-buffered accumulators when they fit)
//   warps 4-7   epilogue       (tcgen05.ld 32x32b -> registers -> dequant/bias/SiLU -> per-warp staging tile -> 128-byte stores;
//                               the coalesced epilogue of the 2-CTA kernel, scale_col of the tile staged in shared memory)
//   warps 8-11  (W4 only) nibble unpack: packed uint8 tile -> sign-extended int8 tile in the swizzled layout
// Split-K (p.splits > 1; M <= 128 shapes with few tiles, e.g. the per-rank shapes of a TP = 8 70B model: one SM lands only
// ~63 B/clk, so a 128 x 128 tile's mainloop is bound by ONE SM's share of the L2 bandwidth): a work unit is (tile, split) and
// covers 1/splits of the int8 k-blocks; the first splits - 1 units of a tile store their int32 partial to a 64 KB workspace
// block and bump the tile's counter, the last unit (which also runs the outlier k-blocks) waits for the counter, bulk-copies
// the blocks into its idle pipeline stages, folds them into its TMEM accumulator and runs the dequant epilogue.  Integer sums:
// bit-identical to the unsplit launch.  A split launch is a single wave (units <= SMs, planned on the host).
// With fused_prologue every warp first runs the activation prologue (rowquant.cuh) on its share of the
// rows, the grid meets at one barrier, and the producer — which already has the first weight tiles in
// flight — starts feeding q_x tiles.
//
// Reference behaviour this replaces: mixlib.int8FusedDequantize[Silu] / int4FusedDequantize[Silu] / gemm and
// the torch.mm outlier GEMM in /root/reference/mixquant/modules/linear.py:234-283, :320-366.
#include "epilogue.cuh"

namespace mixq {

template <int BN, bool W4>
__global__ void __launch_bounds__(GemmCfg<BN, W4>::NUM_THREADS, 1)
mixq_linear_kernel(const __grid_constant__ LinearParams p) {
  using Cfg = GemmCfg<BN, W4>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BM = Cfg::BM;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* bar_empty = bar_full + STAGES;
  uint64_t* bar_ready = bar_empty + STAGES;  // W4: unpacked B tile is ready for the MMA
  uint64_t* bar_tfull = bar_ready + STAGES;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint64_t* bar_sk = bar_tempty + 2;          // split-K: the other splits' partial tiles have landed in shared memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_sk + 1);

  auto stage_a = [&](int s) { return smem + s * Cfg::STAGE_BYTES; };
  auto stage_b = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
  auto stage_bp = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES + Cfg::B_BYTES; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int nk = (p.K + 127) / 128;                 // int8 k-blocks of 128
  const int nko = (p.n_out + 63) / 64;              // fp16 outlier k-blocks of 64
  const int nkt = nk + nko;
  const int MB = (p.M + BM - 1) / BM;
  const int NB = (p.N + BN - 1) / BN;
  const int ntiles = MB * NB;
  const int S = p.splits;                           // >= 1
  const int nunits = ntiles * S;
  const int nk_s = (nk + S - 1) / S;                // int8 k-blocks per split (the host keeps every split non-empty)
  const int acc_cols = BN * (nko > 0 ? 2 : 1);      // s32 accumulator [+ f32 outlier accumulator]
  const int nacc = (2 * acc_cols <= 512) ? 2 : 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_a);
    tma_prefetch_desc(&p.tm_b);
    if (nko > 0) {
      tma_prefetch_desc(&p.tm_oa);
      tma_prefetch_desc(&p.tm_ob);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
      mbar_init(&bar_ready[s], 128);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_tfull[s], 1);
      mbar_init(&bar_tempty[s], 128);
    }
    mbar_init(bar_sk, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  RowQuantSmem* rq_sm = reinterpret_cast<RowQuantSmem*>(smem + STAGES * Cfg::STAGE_BYTES + 256);
  if (p.fused_prologue) rowquant_init(p.rq, rq_sm);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents(); // the next kernel's CTAs may queue up behind this grid
  const uint32_t tmem_base = *tmem_slot;
  unsigned long long* trace = p.trace ? p.trace + static_cast<size_t>(blockIdx.x) * 8 : nullptr;
  if (trace && threadIdx.x == 0) trace[0] = globaltimer_ns();

  // ------------------------------------------------------------------ producer helpers
  // Work items of this CTA in issue order: (tile i, k-block kb), kb in [0, nkt).
  // Pre-barrier we may only touch the weight operand; the activation operand exists after phase A.
  constexpr uint32_t kStageTx = Cfg::A_BYTES + (W4 ? 0 : Cfg::B_BYTES);
  // unit u = (tile, split): int8 k-blocks [kb0, kb0 + nki) and, for the tile's last split, the nko outlier k-blocks
  struct Unit { int tile, split, kb0, nki, nkt; };
  auto unit_of = [&](int u) {
    Unit w;
    w.tile = u / S;
    w.split = u - w.tile * S;
    w.kb0 = w.split * nk_s;
    const int kb1 = (w.kb0 + nk_s < nk) ? w.kb0 + nk_s : nk;
    w.nki = kb1 - w.kb0;
    w.nkt = w.nki + (w.split == S - 1 ? nko : 0);
    return w;
  };
  // item j of a unit -> the k-block index produce() understands: [0, nk) int8, nk + o outlier
  auto item_kb = [&](const Unit& w, int j) { return j < w.nki ? w.kb0 + j : nk + (j - w.nki); };
  auto produce = [&](int tile, int kb, int s, bool do_act, bool do_wgt, bool arm) {
    const int m0 = (tile % MB) * BM;
    const int n0 = (tile / MB) * BN;
    if (kb < nk) {
      if (arm) mbar_arrive_expect_tx(&bar_full[s], Cfg::A_BYTES + (W4 ? Cfg::BP_BYTES : Cfg::B_BYTES));
      if (do_wgt) {
        if (W4) tma_load_2d(&p.tm_b, &bar_full[s], stage_bp(s), kb * 64, n0, kEvictFirst);
        else tma_load_2d(&p.tm_b, &bar_full[s], stage_b(s), kb * 128, n0, kEvictFirst);
      }
      if (do_act) tma_load_2d(&p.tm_a, &bar_full[s], stage_a(s), kb * 128, m0, kEvictLast);
    } else {
      const int ko = (kb - nk) * 64;
      if (arm) mbar_arrive_expect_tx(&bar_full[s], Cfg::A_BYTES + Cfg::B_BYTES);
      if (do_wgt) tma_load_2d(&p.tm_ob, &bar_full[s], stage_b(s), ko, n0, kEvictFirst);
      if (do_act) tma_load_2d(&p.tm_oa, &bar_full[s], stage_a(s), ko, m0, kEvictLast);
    }
    (void)kStageTx;
  };

  const int my_units = (static_cast<int>(blockIdx.x) < nunits)
                           ? (nunits - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1
                           : 0;
  int my_items = 0;
  for (int i = 0; i < my_units; ++i) my_items += unit_of(blockIdx.x + i * gridDim.x).nkt;
  (void)nkt;
  // phase A parks its activation rows in the LAST pipeline stages; the weight prefetch may use the others
  const int row_stages = p.fused_prologue
                             ? static_cast<int>((static_cast<size_t>(p.rq.ngroups) * p.rq.K * 2 + Cfg::STAGE_BYTES - 1) / Cfg::STAGE_BYTES)
                             : 0;
  const int free_stages = STAGES - row_stages;
  const int n_pre = my_items < free_stages ? my_items : free_stages;

  // The quantised weights are constants: fill every free pipeline stage with them BEFORE waiting for the kernels ahead
  // of us in the stream (programmatic dependent launch) — and, with the fused prologue, before phase A.
  if (warp == 3 && lane == 0) {
    int it = 0;
    for (int i = 0; i < my_units && it < n_pre; ++i) {
      const Unit w = unit_of(blockIdx.x + i * gridDim.x);
      for (int j = 0; j < w.nkt && it < n_pre; ++j, ++it) produce(w.tile, item_kb(w, j), it, /*act*/ false, /*wgt*/ true, /*arm*/ true);
    }
  }
  __syncwarp();
  pdl_wait();              // everything below reads or writes tensors that earlier kernels touch

  // ------------------------------------------------------------------ phase A (fused prologue)
  if (p.fused_prologue) {
    uint8_t* rowbuf = smem + static_cast<size_t>(free_stages) * Cfg::STAGE_BYTES;
    rowquant_begin(p.rq, rq_sm, rowbuf);
    rowquant_run(p.rq, rq_sm, rowbuf);
    if (trace && threadIdx.x == 0) trace[1] = globaltimer_ns();
    fence_proxy_async_all();   // q_x / act_outliers were written through the generic proxy; TMA reads them next
    grid_barrier(p.grid_sync);
    if (trace && threadIdx.x == 0) trace[2] = globaltimer_ns();
  }

  // ------------------------------------------------------------------ roles
  // Two producer threads (one TMA op costs its issuing thread ~300 cycles: tools/tma_bw.cu): warp 3 streams the weights
  // and arms the stage barrier, warp 0 streams the activations.
  if (warp == 0 || warp == 3) {
    const bool wgt = warp == 3;
    fence_proxy_async_all();
    int it = 0, s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < my_units; ++i) {
      const Unit w = unit_of(blockIdx.x + i * gridDim.x);
      for (int j = 0; j < w.nkt; ++j, ++it) {
        if (it >= n_pre) mbar_wait(&bar_empty[s], ph ^ 1, 1, s);
        if (elect_one()) {
          if (wgt) {
            if (it >= n_pre) produce(w.tile, item_kb(w, j), s, false, true, true);
          } else {
            produce(w.tile, item_kb(w, j), s, true, false, false);
          }
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // warp-uniform loop, one elected lane issues (see elect_one in ptx.cuh)
    constexpr uint32_t idesc_i8 = make_idesc_i8(BM, BN);
    constexpr uint32_t idesc_f16 = make_idesc_f16(BM, BN);
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < my_units; ++i) {
      const Unit w = unit_of(blockIdx.x + i * gridDim.x);
      const int as = i % nacc;
      const uint32_t aph = (i / nacc) & 1;
      mbar_wait(&bar_tempty[as], aph ^ 1, 2, as);
      tc_fence_after();
      const uint32_t d_int = tmem_base + as * acc_cols;
      const uint32_t d_out = d_int + BN;
      for (int kb = 0; kb < w.nkt; ++kb) {
        if (W4) mbar_wait(&bar_ready[s], ph, 3, s);
        mbar_wait(&bar_full[s], ph, 4, s);
        tc_fence_after();
        const uint64_t da = make_sw128_kmajor_desc(smem_u32(stage_a(s)));
        const uint64_t db = make_sw128_kmajor_desc(smem_u32(stage_b(s)));
        if (elect_one()) {
          if (trace && i == 0 && kb == 0) trace[3] = globaltimer_ns();
          if (kb < w.nki) {
#pragma unroll
            for (int k = 0; k < 4; ++k)   // 4 x (K = 32 int8 = 32 B); +2 in the >>4-encoded start address
              umma_i8(d_int, da + 2 * k, db + 2 * k, idesc_i8, (kb | k) != 0);
          } else {
            const int kbo = kb - w.nki;
            int ksteps = (p.n_out - kbo * 64 + 15) / 16;
            if (ksteps > 4) ksteps = 4;
            for (int k = 0; k < ksteps; ++k)  // K = 16 fp16 = 32 B
              umma_f16(d_out, da + 2 * k, db + 2 * k, idesc_f16, (kbo | k) != 0);
          }
          umma_commit(&bar_empty[s]);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(&bar_tfull[as]);
      __syncwarp();
    }
    if (trace && lane == 0) trace[4] = globaltimer_ns();
  } else if (warp >= 4 && warp < 8) {
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES + 256 + 512;
    const uint32_t stage_sa = smem_u32(epi_smem + (warp - 4) * kEpiStageBytes);      // this warp's 4 KB staging tile
    __half* s_scale = reinterpret_cast<__half*>(epi_smem + Cfg::EPI_WARPS * kEpiStageBytes);
    const uint32_t scale_sa = smem_u32(s_scale);
    const int mode = (p.outl != nullptr || p.bias != nullptr || p.act == 1) ? 2 : (p.residual != nullptr ? 1 : 0);
    uint32_t sk_uses = 0;      // finished split tiles so far (phase of bar_sk)
    for (int i = 0; i < my_units; ++i) {
      const Unit w = unit_of(blockIdx.x + i * gridDim.x);
      const int tile = w.tile;
      const int m0 = (tile % MB) * BM;
      const int n0 = (tile / MB) * BN;
      const int as = i % nacc;
      const uint32_t aph = (i / nacc) & 1;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      float xs = 0.f;
      if (p.epilogue == EPI_DEQUANT_F16 && row_ok) xs = __half2float(p.x_scale[row]);
      // scale_col of the tile -> shared memory while the accumulator is still being computed (every row reads all of it;
      // fetched per 16-column group from global it cost one exposed L2 round trip per group: 6 us per 128 x 128 tile)
      if (p.epilogue == EPI_DEQUANT_F16 && !(S > 1 && w.split < S - 1)) {
        named_bar_sync(13, 128);        // the previous tile's readers are done with s_scale
        for (int j = (threadIdx.x - 128) * 8; j < BN; j += 128 * 8)
          *reinterpret_cast<uint4*>(s_scale + j) = (n0 + j < p.N) ? __ldg(reinterpret_cast<const uint4*>(p.scale_col + n0 + j)) : make_uint4(0, 0, 0, 0);
        named_bar_sync(13, 128);
      }

      mbar_wait_warp(&bar_tfull[as], aph, 5, as);
      tc_fence_after();
      const uint32_t t_int = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * acc_cols;
      const uint32_t t_out = t_int + BN;

      const bool partial = S > 1 && w.split < S - 1;
      // split-K workspace: tile t owns S - 1 blocks of 64 KB, one per non-final split
      int32_t* sk_blocks = p.sk_ws + static_cast<size_t>(tile) * (S - 1) * kSplitKBlockInts;
      if (partial) {
        // partial sum -> shared memory (the pipeline stages are idle: single wave, this is the CTA's only unit) -> ONE bulk store
        // into this split's workspace block -> counter bump.  (Direct 16-byte stores + __threadfence by 128 threads took 4 us.)
        splitk_stage_partial(t_int, smem_u32(smem), q * 32 + lane);
        fence_proxy_async_smem();
        named_bar_sync(13, 128);
        if (warp == 4 && lane == 0) {
          bulk_store(sk_blocks + static_cast<size_t>(w.split) * kSplitKBlockInts, smem_u32(smem), kSplitKBlockInts * 4);
          tma_store_commit();
          tma_store_wait_all<0>();     // the writes are complete, not just read out of shared memory
          fence_proxy_async_all();
          __threadfence();
          atomicAdd(p.sk_cnt + tile, 1u);
        }
      } else {
        if (S > 1) {
          // the tile's last split: wait for the others (they hold SMs of this very grid and never wait for anybody), bring their
          // blocks into the pipeline stages — idle by now: a split launch is a single wave, this is the CTA's only unit — with
          // one bulk copy each, and fold them into the accumulator
          if (warp == 4 && lane == 0) {
            const uint64_t t0 = globaltimer_ns();
            uint32_t spins = 0;
            while (ld_acquire_u32(p.sk_cnt + tile) < static_cast<uint32_t>(S - 1)) {
              if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > MIXQ_SPIN_TIMEOUT_NS) spin_timeout_trap(10, tile, S);
            }
            p.sk_cnt[tile] = 0;        // re-armed for the next launch (nobody else touches it any more in this one)
            fence_proxy_async_all();   // (generic-proxy acquire above -> the async-proxy reads of the bulk copies below)
            mbar_arrive_expect_tx(bar_sk, static_cast<uint32_t>(S - 1) * (kSplitKBlockInts * 4));
            for (int sp = 0; sp < S - 1; ++sp)
              bulk_load(smem + sp * (kSplitKBlockInts * 4), sk_blocks + static_cast<size_t>(sp) * kSplitKBlockInts, kSplitKBlockInts * 4, bar_sk);
            if (trace) trace[6] = globaltimer_ns();    // partials have arrived in the workspace
          }
          mbar_wait_warp(bar_sk, sk_uses & 1u, 11, tile);
          ++sk_uses;
          splitk_fold_smem(t_int, smem_u32(smem), S - 1, q * 32 + lane);
          if (trace && warp == 4 && lane == 0) trace[7] = globaltimer_ns();    // folded into the accumulator
        }
        if (p.epilogue != EPI_DEQUANT_F16) {       // raw int32 accumulators (mixlib.gemm)
          epilogue_span<false>(p, t_int, 0u, row, row_ok, n0, BN, xs, nullptr, 0);
        } else {
          const int mb = m0 + q * 32;
          if (nko > 0) {
            if (mode == 0) epilogue_run_coalesced<true, 0>(p, stage_sa, t_int, t_out, mb, n0, BN, xs, scale_sa, lane);
            else if (mode == 1) epilogue_run_coalesced<true, 1>(p, stage_sa, t_int, t_out, mb, n0, BN, xs, scale_sa, lane);
            else epilogue_run_coalesced<true, 2>(p, stage_sa, t_int, t_out, mb, n0, BN, xs, scale_sa, lane);
          } else {
            if (mode == 0) epilogue_run_coalesced<false, 0>(p, stage_sa, t_int, 0u, mb, n0, BN, xs, scale_sa, lane);
            else if (mode == 1) epilogue_run_coalesced<false, 1>(p, stage_sa, t_int, 0u, mb, n0, BN, xs, scale_sa, lane);
            else epilogue_run_coalesced<false, 2>(p, stage_sa, t_int, 0u, mb, n0, BN, xs, scale_sa, lane);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&bar_tempty[as]);
    }
    if (trace && warp == 4 && lane == 0) trace[5] = globaltimer_ns();
  } else if (W4 && warp >= 8) {
    // Nibble unpack: thread t owns weight rows t, t+128, ... of the tile.  Packed row = 64 B (128 nibbles,
    // low nibble = even k: linear.py:14-18); unpacked row = 128 B written as 8 x 16-byte chunks at the
    // SWIZZLE_128B position chunk ^ (row & 7) — the layout the UMMA descriptor expects.
    const int t = threadIdx.x - 256;
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < my_units; ++i) {
      const Unit w = unit_of(blockIdx.x + i * gridDim.x);
      for (int kb = 0; kb < w.nkt; ++kb) {
        mbar_wait(&bar_full[s], ph, 6, s);
        if (kb < w.nki) {
          for (int r = t; r < BN; r += 128) {
            const uint4* src = reinterpret_cast<const uint4*>(stage_bp(s) + r * 64);
            uint8_t* drow = stage_b(s) + r * 128;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const uint4 pk = src[v];
              const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
              uint32_t o[8];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t lo = nib_lo_s8x4(w[j]);
                const uint32_t hi = nib_hi_s8x4(w[j]);
                o[2 * j] = __byte_perm(lo, hi, 0x5140);
                o[2 * j + 1] = __byte_perm(lo, hi, 0x7362);
              }
              const int c0 = (2 * v) ^ (r & 7);
              const int c1 = (2 * v + 1) ^ (r & 7);
              *reinterpret_cast<uint4*>(drow + c0 * 16) = make_uint4(o[0], o[1], o[2], o[3]);
              *reinterpret_cast<uint4*>(drow + c1 * 16) = make_uint4(o[4], o[5], o[6], o[7]);
            }
          }
          fence_proxy_async_smem();   // generic-proxy smem writes -> visible to tcgen05.mma (async proxy)
        }
        mbar_arrive(&bar_ready[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template __global__ void mixq_linear_kernel<128, false>(const __grid_constant__ LinearParams);
template __global__ void mixq_linear_kernel<256, false>(const __grid_constant__ LinearParams);
template __global__ void mixq_linear_kernel<128, true>(const __grid_constant__ LinearParams);
template __global__ void mixq_linear_kernel<256, true>(const __grid_constant__ LinearParams);

}  // namespace mixq
