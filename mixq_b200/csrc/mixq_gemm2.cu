// The MixLinear hot-path kernel, 2-CTA form (tcgen05 cta_group::2) — used for M > 128.
//
// Why: at M = 512 the 1-CTA kernel (mixq_gemm.cu, 128 x BN tiles) is bound by L2 -> SM traffic, not by the tensor pipe
// or HBM: every 128-row tile re-reads its weight tile and every BN-column tile re-reads the activations
// (ncu: l1tex__m_xbar2l1tex_read_bytes = 6x the algorithmic bytes, tensor pipe 21-27 % active).  A CTA PAIR owns a
// 256 x BN output tile: each CTA stages its own 128 activation rows and only HALF of the BN weight rows, one
// tcgen05.mma.cta_group::2 (issued by the pair's leader) reads both halves — half the L2 bytes per MMA cycle.
// BN is a run-time multiple of 32 (<= 256) chosen on the host so that the tile count fills the 74 pairs evenly.
//
//   warp 0 / 3  TMA producers: activations / weights (both CTAs; loads land in the local smem, complete_tx on the
//               LEADER's mbarrier)
//   warp 1      MMA issuer   (leader CTA only; tcgen05.commit multicast frees the stage in both CTAs)
//   warp 2      TMEM allocator (cta_group::2, 512 columns in each CTA)
//   warps 4-11  epilogue     (2 warps per TMEM lane quarter, one half of the tile's columns each)
// Phase A (activation prologue, rowquant.cuh) and the grid barrier are the same as in the 1-CTA kernel.
//
// Reference behaviour this replaces: mixlib.int8FusedDequantize[Silu] and the torch.mm outlier GEMM in
// /root/reference/mixquant/modules/linear.py:244-283, :329-351.
#include "epilogue.cuh"

namespace mixq {

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Gemm2Cfg::NUM_THREADS, 1)
mixq_linear2_kernel(const __grid_constant__ LinearParams p) {
  using Cfg = Gemm2Cfg;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);   // used in the leader only
  uint64_t* bar_empty = bar_full + STAGES;                                              // per CTA (multicast commit)
  uint64_t* bar_tfull = bar_empty + STAGES;                                             // per CTA (multicast commit)
  uint64_t* bar_tempty = bar_tfull + 4;                                                 // used in the leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 4);

  auto stage_a = [&](int s) { return smem + s * Cfg::STAGE_BYTES; };
  auto stage_b = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader of the pair
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  const int bn = p.bn;                               // tile width, multiple of 32
  const int bh = bn >> 1;                            // weight rows staged by each CTA
  const int nk = (p.K + 127) / 128;                  // int8 k-blocks of 128
  const int nko = (p.n_out + 63) / 64;               // fp16 outlier k-blocks of 64
  const int nkt = nk + nko;
  const int MP = (p.M + 255) / 256;
  const int NT = (p.N + bn - 1) / bn;
  const int ntiles = MP * NT;
  const bool has_o = nko > 0;
  // int32 accumulator slots in a ring (a tile's epilogue overlaps the next tiles' MMAs); the fp32 outlier slot is single
  int nint = (512 - (has_o ? bn : 0)) / bn;
  if (nint > 4) nint = 4;
  const uint32_t col_outl = static_cast<uint32_t>(nint * bn);
  const uint32_t stage_tx = 2u * (Cfg::A_BYTES + static_cast<uint32_t>(bh) * 128u);   // both CTAs' bytes land on one barrier

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_a);
    tma_prefetch_desc(&p.tm_b);
    if (has_o) {
      tma_prefetch_desc(&p.tm_oa);
      tma_prefetch_desc(&p.tm_ob);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&bar_tfull[s], 1);
      mbar_init(&bar_tempty[s], 2 * Cfg::EPI_WARPS);   // one arrival per epilogue warp of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_slot, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // the peer's barriers exist before anything remote touches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  unsigned long long* trace = p.trace ? p.trace + static_cast<size_t>(blockIdx.x) * 8 : nullptr;
  if (trace && threadIdx.x == 0) trace[0] = globaltimer_ns();

  // ------------------------------------------------------------------ producer helper
  auto produce = [&](int tile, int kb, int s, bool do_act, bool do_wgt, bool arm) {
    const int m0 = (tile % MP) * 256 + static_cast<int>(rank) * 128;
    const int n0 = (tile / MP) * bn + static_cast<int>(rank) * bh;
    const uint32_t full_leader = mapa_u32(smem_u32(&bar_full[s]), 0);
    if (arm && leader) mbar_arrive_expect_tx(&bar_full[s], stage_tx);
    if (kb < nk) {
      if (do_wgt) tma_load_2d_2cta(&p.tm_b, full_leader, stage_b(s), kb * 128, n0, kEvictFirst);
      if (do_act) tma_load_2d_2cta(&p.tm_a, full_leader, stage_a(s), kb * 128, m0, kEvictLast);
    } else {
      const int ko = (kb - nk) * 64;
      if (do_wgt) tma_load_2d_2cta(&p.tm_ob, full_leader, stage_b(s), ko, n0, kEvictFirst);
      if (do_act) tma_load_2d_2cta(&p.tm_oa, full_leader, stage_a(s), ko, m0, kEvictLast);
    }
  };

  const int my_tiles = (pair < ntiles) ? (ntiles - 1 - pair) / npairs + 1 : 0;
  const int my_items = my_tiles * nkt;
  const int row_stages = p.fused_prologue
                             ? static_cast<int>((static_cast<size_t>(p.rq.ngroups) * p.rq.K * 2 + Cfg::STAGE_BYTES - 1) / Cfg::STAGE_BYTES)
                             : 0;
  const int free_stages = STAGES - row_stages;
  const int n_pre = (p.fused_prologue && my_items > 0) ? (my_items < free_stages ? my_items : free_stages) : 0;

  // ------------------------------------------------------------------ phase A (fused prologue)
  if (p.fused_prologue) {
    RowQuantSmem* rq_sm = reinterpret_cast<RowQuantSmem*>(smem + STAGES * Cfg::STAGE_BYTES + 256);
    uint8_t* rowbuf = smem + static_cast<size_t>(free_stages) * Cfg::STAGE_BYTES;
    rowquant_begin(p.rq, rq_sm, rowbuf);      // activation rows first: they are on the critical path
    if (warp == 3 && lane == 0) {
      for (int it = 0; it < n_pre; ++it) {
        const int tile = pair + (it / nkt) * npairs;
        produce(tile, it % nkt, it, /*act*/ false, /*wgt*/ true, /*arm*/ true);
      }
    }
    __syncwarp();
    rowquant_run(p.rq, rq_sm, rowbuf);
    if (trace && threadIdx.x == 0) trace[1] = globaltimer_ns();
    fence_proxy_async_all();   // q_x / act_outliers were written through the generic proxy; TMA reads them next
    grid_barrier(p.grid_sync);
    if (trace && threadIdx.x == 0) trace[2] = globaltimer_ns();
  }

  // ------------------------------------------------------------------ roles
  // Two producer threads: one TMA op costs its issuing thread ~300 cycles (tools/tma_bw.cu: 54 B/clk/SM with one issuer,
  // 71 with two), and a pair tile needs up to 64 B/clk/SM.  (Tried and dropped: pulling the weight boxes into L2 ahead of
  // the loads — cp.async.bulk.prefetch.tensor costs a TMA issue slot per box and made the loop 15 % slower; a spare warp
  // issuing prefetch.global.L2 changed nothing: the loop is bound by bytes landing per SM, not by DRAM latency.)  Warp 3 streams the weights (and arms the stage barrier),
  // warp 0 the activations.
  if (warp == 0 || warp == 3) {
    if (lane == 0) {
      const bool wgt = warp == 3;
      fence_proxy_async_all();
      int it = 0, s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int tile = pair + i * npairs;
        for (int kb = 0; kb < nkt; ++kb, ++it) {
          if (it >= n_pre) mbar_wait(&bar_empty[s], ph ^ 1, 1, s);
          if (wgt) {
            if (it >= n_pre) produce(tile, kb, s, false, true, true);
          } else {
            produce(tile, kb, s, true, false, false);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      const uint32_t idesc_i8 = make_idesc_i8_rt(256, bn);
      const uint32_t idesc_f16 = make_idesc_f16_rt(256, bn);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int slot = i % nint;
        const uint32_t use = static_cast<uint32_t>(i / nint);
        mbar_wait(&bar_tempty[slot], (use & 1) ^ 1, 2, slot);      // the epilogues of both CTAs drained this slot
        tc_fence_after();
        const uint32_t d_int = tmem_base + static_cast<uint32_t>(slot * bn);
        const uint32_t d_out = tmem_base + col_outl;
        for (int kb = 0; kb < nkt; ++kb) {
          if (kb == nk && nint >= 2 && i > 0) {
            // the single fp32 outlier slot is still being read by the previous tile's epilogue
            mbar_wait(&bar_tempty[(i - 1) % nint], static_cast<uint32_t>((i - 1) / nint) & 1, 8, i);
            tc_fence_after();
          }
          mbar_wait(&bar_full[s], ph, 4, s);
          tc_fence_after();
          if (trace && i == 0 && kb == 0) trace[3] = globaltimer_ns();
          const uint64_t da = make_sw128_kmajor_desc(smem_u32(stage_a(s)));
          const uint64_t db = make_sw128_kmajor_desc(smem_u32(stage_b(s)));
          if (kb < nk) {
#pragma unroll
            for (int k = 0; k < 4; ++k)   // 4 x (K = 32 int8 = 32 B); +2 in the >>4-encoded start address
              umma_i8_2cta(d_int, da + 2 * k, db + 2 * k, idesc_i8, (kb | k) != 0);
          } else {
            const int kbo = kb - nk;
            int ksteps = (p.n_out - kbo * 64 + 15) / 16;
            if (ksteps > 4) ksteps = 4;
            for (int k = 0; k < ksteps; ++k)  // K = 16 fp16 = 32 B
              umma_f16_2cta(d_out, da + 2 * k, db + 2 * k, idesc_f16, (kbo | k) != 0);
          }
          umma_commit_2cta(&bar_empty[s], 0x3);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit_2cta(&bar_tfull[slot], 0x3);
      }
      if (trace) trace[4] = globaltimer_ns();
    }
  } else if (warp >= 4) {
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = (warp - 4) >> 2;       // which half of the tile's columns
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = pair + i * npairs;
      const int m0 = (tile % MP) * 256 + static_cast<int>(rank) * 128;
      const int n0 = (tile / MP) * bn;
      const int slot = i % nint;
      const uint32_t use = static_cast<uint32_t>(i / nint);
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      float xs = 0.f;
      if (p.epilogue == EPI_DEQUANT_F16 && row_ok) xs = __half2float(p.x_scale[row]);

      mbar_wait(&bar_tfull[slot], use & 1, 5, slot);
      tc_fence_after();
      const uint32_t t_int = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(slot * bn);
      const uint32_t t_out = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + col_outl;
      const int span = bn >> 1;           // this warp's contiguous half of the tile's columns
      const int c0 = half * span;
      if (has_o) epilogue_span<true>(p, t_int + c0, t_out + c0, row, row_ok, n0 + c0, span, xs);
      else epilogue_span<false>(p, t_int + c0, 0u, row, row_ok, n0 + c0, span, xs);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_tempty[slot]), 0));
    }
    if (trace && warp == 4 && lane == 0) trace[5] = globaltimer_ns();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // nobody leaves while the peer may still read our smem / signal our barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace mixq
