// The MixLinear hot-path kernel, 2-CTA "strip" form (tcgen05 cta_group::2) — used for M > 128.
//
// What bounds this GEMM on B200 (tools/tma_bw.cu, tools/mma_bw.cu, profiles/): the int8 tensor pipe does 8192 MAC/clk/SM,
// but one SM cannot LAND more than ~55-70 B/clk through TMA (about 200 KB of smem in flight against >1 us of latency),
// and at M = 512 there are only a couple of tiles per SM, so pipeline fill, drain and wave quantisation are first-order.
// Hence:
//   * a CTA PAIR owns a 256 x W output tile (one tcgen05.mma.cta_group::2 reads the weight halves staged by both CTAs):
//     bytes landing per MMA cycle = 8192/W + 32 per SM, so W is made as wide as the problem allows — up to 512, i.e. the
//     WHOLE of TMEM for one int32 accumulator, issued as two MMA column chunks per k-step;
//   * W is chosen on the host so that every pair gets ONE tile when possible (no second, half-empty wave);
//   * the int32 accumulator takes W TMEM columns, the fp32 accumulator of the skinny fp16 outlier GEMM only what is left
//     (R = 512 - W): the outlier k-blocks come LAST in the tile, stay resident in their pipeline stages, and the MMA
//     warp re-issues the (tiny) outlier MMAs for R columns at a time while the epilogue warps drain the tile pass by
//     pass.  Nothing leaves the chip, and the reference's rounding (torch.mm -> fp16, linear.py:248) is kept.
//
//   warp 0 / 3  TMA producers: activations / weights (both CTAs; loads land locally, complete_tx on the LEADER's mbarrier)
//   warp 1      MMA issuer   (leader CTA only; tcgen05.commit multicast frees the stage in both CTAs)
//   warp 2      TMEM allocator (cta_group::2, 512 columns in each CTA)
//   warps 4-11  epilogue     (2 warps per TMEM lane quarter, one half of the tile's columns each)
// Phase A (activation prologue, rowquant.cuh) and the grid barrier are the same as in the 1-CTA kernel.
//
// Column bookkeeping: CTA r of the pair stages weight rows [n0 + r*W/2, +W/2).  MMA chunk c (N_c columns, using rows
// [o_c, o_c + N_c/2) of each CTA's half) therefore produces accumulator columns [0, N_c/2) = output columns
// n0 + o_c + j and [N_c/2, N_c) = n0 + W/2 + o_c + j: epilogue warp-half h walks output columns [n0 + h*W/2, +W/2).
//
// Reference behaviour this replaces: mixlib.int8FusedDequantize[Silu] and the torch.mm outlier GEMM in
// /root/reference/mixquant/modules/linear.py:244-283, :329-351.
#include "epilogue.cuh"

namespace mixq {

// W4 (packed-nibble weights, unpacked in shared memory by the epilogue warps) is a TEMPLATE parameter: as a run-time flag its
// mere presence cost the W8 instantiation ~1 us per launch (A/B builds, profiles/r02_ab_*: register allocation and scheduling
// of the MMA / producer loops).
template <bool W4>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Gemm2Cfg::NUM_THREADS, 1)
mixq_linear2_kernel(const __grid_constant__ LinearParams p) {
  using Cfg = Gemm2Cfg;
  constexpr int MAXS = Cfg::MAX_STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + Cfg::PIPE_BYTES);    // used in the leader only
  uint64_t* bar_empty = bar_full + MAXS;                                       // per CTA (multicast commit)
  uint64_t* bar_tfull = bar_empty + MAXS;                                      // [0]: one phase per epilogue pass; per CTA
  uint64_t* bar_tempty = bar_tfull + 2;                                        // [0]: pass consumed by both CTAs' epilogues; leader
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);
  // W4 only.  The packed weight rows travel through their OWN ring of p.npacked slots behind the main stages (deep enough to
  // cover HBM latency; the 3 main stages only have to cover the unpack -> MMA -> commit chain): bar_bfull / bar_bempty are this
  // CTA's (slot landed / slot consumed by its 8 unpack warps; they live in the slack of the RowQuantSmem block); bar_ready
  // (leader) counts the unpack warps of both CTAs that have expanded a main stage's weight half to int8.
  uint64_t* bar_bfull = reinterpret_cast<uint64_t*>(smem + Cfg::PIPE_BYTES + 256 + 384);
  uint64_t* bar_bempty = bar_bfull + MAXS;
  uint64_t* bar_ready = reinterpret_cast<uint64_t*>(smem + Cfg::PIPE_BYTES + 168);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader of the pair
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  const int W = p.bn;                                // tile width, multiple of 32, <= 512
  const int bh = W >> 1;                             // weight rows staged by each CTA
  const int n1 = (W <= 256) ? W : (((W >> 1) + 31) & ~31);   // MMA column chunks
  const int n2 = W - n1;
  const int stage_bytes = p.stage_bytes;
  const int nstages = p.nstages;
  // One TMA op costs the SM's TMA unit ~340 clocks whatever its size (tools/tma_mc_bw.cu: 16 KB boxes land at 43 B/clk/SM,
  // 32 KB boxes at 78), and a k-block needs two ops (activations, weights): narrow tiles, whose MMAs take < 700 clocks per
  // k-block, are paced by the op count.  They therefore move KA = 2 k-atoms (256 bytes of K) per op and stage.
  const int KA = p.k_atoms;
  const int nk = (p.K + 128 * KA - 1) / (128 * KA);  // int8 pipeline items (KA k-atoms of 128 bytes each)
  const int nko = (p.n_out + 63) / 64;               // fp16 outlier k-blocks of 64 — the LAST items of a tile
  const int nkt = nk + nko;
  const int MP = (p.M + 255) / 256;
  const bool pairm = p.pair_swiglu != 0;             // SwiGLU pair: CTA 0 stages gate rows, CTA 1 the same up rows
  const int wout = pairm ? bh : W;                   // output columns per tile
  const int NT = (p.N + wout - 1) / wout;
  const int ntiles = MP * NT;
  const bool has_o = nko > 0;
  constexpr bool w4 = W4;
  // bytes per k-atom landing on the leader's full barrier: both CTAs' activations and (W8) weights; W4 weights land on bar_bfull
  const uint32_t atom_tx = w4 ? 2u * static_cast<uint32_t>(Cfg::A_BYTES) : 2u * static_cast<uint32_t>(Cfg::A_BYTES + bh * 128);
  const uint32_t atom_tx_o = 2u * static_cast<uint32_t>(Cfg::A_BYTES + bh * 128);   // fp16 outlier k-blocks: never packed
  const uint32_t b_atom = static_cast<uint32_t>(bh) * 128u;                         // bytes between the k-atoms of a weight stage
  // TMEM plan (plan_tmem, mixq_gemm.cuh).  Passes are numbered globally (G = tile * P + c): pass G uses barrier pair /
  // buffer G % NB in phase (G / NB) & 1.
  const int my_tiles_geo = (pair < ntiles) ? (ntiles - 1 - pair) / npairs + 1 : 0;
  // W4: the epilogue warps unpack during the mainloop, so tiles do not overlap: one accumulator slot
  const TmemPlan tp = plan_tmem(W, has_o, w4 ? 1 : my_tiles_geo, (p.ablate & 16) != 0);   // (tuning knob 16: one big pass buffer)
  const int SLOTS = tp.slots, P = tp.passes, R = tp.pass_cols, NB = tp.buffers;
  const uint32_t outl_col0 = static_cast<uint32_t>(SLOTS * W);   // first TMEM column of the outlier buffers

  auto stage_a = [&](int s) { return smem + static_cast<size_t>(s) * stage_bytes; };
  auto stage_b = [&](int s) { return smem + static_cast<size_t>(s) * stage_bytes + static_cast<size_t>(KA) * Cfg::A_BYTES; };
  const int NP = p.npacked;                                   // W4: packed-row ring
  const uint32_t pk_pitch = (static_cast<uint32_t>(bh) * 64u + 127u) & ~127u;
  auto packed_slot = [&](int ps) { return smem + static_cast<size_t>(nstages) * stage_bytes + static_cast<size_t>(ps) * pk_pitch; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_a);
    tma_prefetch_desc(&p.tm_b);
    if (pairm) tma_prefetch_desc(&p.tm_b2);
    if (has_o) {
      tma_prefetch_desc(&p.tm_oa);
      tma_prefetch_desc(&p.tm_ob);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MAXS; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
      if (w4) {
        mbar_init(&bar_bfull[s], 1);
        mbar_init(&bar_bempty[s], Cfg::EPI_WARPS);      // this CTA's unpack warps (both groups)
        mbar_init(&bar_ready[s], 2 * Cfg::EPI_WARPS);   // ... of both CTAs
      }
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_tfull[s], 1);
      mbar_init(&bar_tempty[s], 2 * Cfg::EPI_WARPS);   // one arrival per epilogue warp of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_slot, 512);
    tmem_relinquish_2cta();
  }
  RowQuantSmem* rq_sm = reinterpret_cast<RowQuantSmem*>(smem + Cfg::PIPE_BYTES + 256);
  if (p.fused_prologue) rowquant_init(p.rq, rq_sm);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // the peer's barriers exist before anything remote touches them
  tc_fence_after();
  pdl_launch_dependents(); // the next kernel's CTAs may queue up behind this grid (they take an SM as soon as ours exits)
  const uint32_t tmem_base = *tmem_slot;
  unsigned long long* trace = p.trace ? p.trace + static_cast<size_t>(blockIdx.x) * 8 : nullptr;
  if (trace && threadIdx.x == 0) trace[0] = globaltimer_ns();

  // ------------------------------------------------------------------ producer helper
  // k-block order within a tile: the nk int8 blocks, then the nko outlier blocks.
  // (Tried and dropped: letting every N-strip start its K loop at a different block so that the ~35 pairs of an M-block
  // do not ask L2 for the same q_x lines at the same moment — 10 % SLOWER: simultaneous identical requests are merged
  // in L2, lock-step is the cheap case.)
  auto produce = [&](int tile, int kb, int s, bool do_act, bool do_wgt, bool arm) {
    const int m0 = (tile % MP) * 256 + static_cast<int>(rank) * 128;
    const int n0 = pairm ? (tile / MP) * bh : (tile / MP) * W + static_cast<int>(rank) * bh;
    const CUtensorMap* tmb = (pairm && rank == 1) ? &p.tm_b2 : &p.tm_b;
    const CUtensorMap* tmob = (pairm && rank == 1) ? &p.tm_ob2 : &p.tm_ob;
    const uint32_t full_leader = mapa_u32(smem_u32(&bar_full[s]), 0);
    if (kb < nk) {
      if (p.ablate & 6) {   // tuning aid: leave one operand stream out (the MMAs then chew on stale shared memory)
        if (p.ablate & 2) do_act = false;
        if (p.ablate & 4) do_wgt = false;
        if (arm && leader)
          mbar_arrive_expect_tx(&bar_full[s], 2u * KA * (((p.ablate & 2) ? 0u : Cfg::A_BYTES) + ((p.ablate & 4) ? 0u : b_atom)));
        arm = false;
      }
      if (arm && leader) mbar_arrive_expect_tx(&bar_full[s], atom_tx * static_cast<uint32_t>(KA));
      if (w4) {   // activations only: the packed weights go through produce_packed / the unpack warps
        if (do_act) tma_load_2d_2cta(&p.tm_a, full_leader, stage_a(s), kb * 128, m0, kEvictLast);
      } else if (KA == 1) {
        const int k0 = kb * 128;
        if (do_wgt) tma_load_2d_2cta(tmb, full_leader, stage_b(s), k0, n0, kEvictFirst);
        if (do_act) tma_load_2d_2cta(&p.tm_a, full_leader, stage_a(s), k0, m0, kEvictLast);
      } else {   // atoms past the end of K are out of bounds of the 3-D view: zero-filled, counted, and harmless in the MMA
        if (do_wgt) tma_load_3d_2cta(tmb, full_leader, stage_b(s), 0, n0, kb * KA, kEvictFirst);
        if (do_act) tma_load_3d_2cta(&p.tm_a, full_leader, stage_a(s), 0, m0, kb * KA, kEvictLast);
      }
    } else {
      if (arm && leader) mbar_arrive_expect_tx(&bar_full[s], atom_tx_o);
      const int ko = (kb - nk) * 64;
      if (do_wgt) tma_load_2d_2cta(tmob, full_leader, stage_b(s), ko, n0, kEvictFirst);
      if (do_act) tma_load_2d_2cta(&p.tm_oa, full_leader, stage_a(s), ko, m0, kEvictLast);
    }
  };

  // W4: packed weight rows of int8 k-block kb of `tile` -> packed slot ps (64 bytes of every weight row, no swizzle)
  auto produce_packed = [&](int tile, int kb, int ps) {
    const int n0 = pairm ? (tile / MP) * bh : (tile / MP) * W + static_cast<int>(rank) * bh;
    const CUtensorMap* tmb = (pairm && rank == 1) ? &p.tm_b2 : &p.tm_b;
    mbar_arrive_expect_tx(&bar_bfull[ps], static_cast<uint32_t>(bh) * 64u);
    tma_load_2d_2cta(tmb, smem_u32(&bar_bfull[ps]), packed_slot(ps), (kb * 128) >> 1, n0, kEvictFirst);
  };

  const int my_tiles = (pair < ntiles) ? (ntiles - 1 - pair) / npairs + 1 : 0;
  const int my_items = my_tiles * nkt;
  const int row_stages = p.fused_prologue
                             ? static_cast<int>((static_cast<size_t>(p.rq.ngroups) * p.rq.K * 2 + stage_bytes - 1) / stage_bytes)
                             : 0;
  const int free_stages = nstages - row_stages;
  // W4: nothing is prefetched into the main stages (their weight halves are written by the unpack warps); the whole packed
  // ring is filled instead
  const int n_pre = w4 ? 0 : (my_items < free_stages ? my_items : free_stages);
  const int my_packed = my_tiles * nk;
  const int n_pre_p = my_packed < NP ? my_packed : NP;

  // The quantised weights are constants: fill every free pipeline stage with them BEFORE waiting for the kernels ahead
  // of us in the stream (programmatic dependent launch) — and, with the fused prologue, before phase A.
  if (warp == 3 && lane == 0) {
    for (int it = 0; it < n_pre; ++it) {
      const int tile = pair + (it / nkt) * npairs;
      produce(tile, it % nkt, it, /*act*/ false, /*wgt*/ true, /*arm*/ true);
    }
    if (w4)
      for (int pi = 0; pi < n_pre_p; ++pi) produce_packed(pair + (pi / nk) * npairs, pi % nk, pi);
    // (Tried and dropped: cp.async.bulk.prefetch.tensor of the next ring's worth of weight boxes here, to cover the ~1.3 us MMA
    // stall at the first wrap of the ring — HBM latency exceeds the 4 stages' cover — made every launch ~2 us SLOWER.)
  }
  __syncwarp();
  pdl_wait();              // everything below reads or writes tensors that earlier kernels touch

  // ------------------------------------------------------------------ phase A (fused prologue)
  if (p.fused_prologue) {
    uint8_t* rowbuf = smem + static_cast<size_t>(free_stages) * stage_bytes;
    rowquant_begin(p.rq, rq_sm, rowbuf);
    rowquant_run(p.rq, rq_sm, rowbuf);
    if (trace && threadIdx.x == 0) trace[1] = globaltimer_ns();
    fence_proxy_async_all();   // q_x / act_outliers were written through the generic proxy; TMA reads them next
    grid_barrier(p.grid_sync, p.trace ? p.trace + 1700 + static_cast<size_t>(blockIdx.x) * 4 : nullptr);
    if (trace && threadIdx.x == 0) trace[2] = globaltimer_ns();
  }

  // ------------------------------------------------------------------ roles
  // Two producer threads: one TMA op costs its issuing thread ~300 cycles (tools/tma_bw.cu: 54 B/clk/SM with one issuer,
  // 71 with two).  Warp 3 streams the weights (and arms the stage barrier), warp 0 the activations.
  // (Tried and dropped: pulling the weight boxes into L2 ahead of the loads — cp.async.bulk.prefetch.tensor costs a TMA
  // issue slot per box and made the loop 15 % slower; a spare warp issuing prefetch.global.L2 changed nothing.)
  if (w4 && (warp == 0 || warp == 3)) {
    // W4 producers.  Warp 3: ONLY the packed weight rows, into the packed ring — gated by bar_bempty alone, so it runs up to NP
    // k-blocks ahead of the MMAs (it must not touch the main stages: a producer that skips uses of a stage cannot wait on that
    // stage's parity barrier).  Warp 0: everything that lands in a main stage — the activations of every item and both operands
    // of the fp16 outlier blocks — and the arming of the leader's full barrier.
    const bool wgt = warp == 3;
    fence_proxy_async_all();
    int it = 0, s = 0, pi = 0, ps = 0;
    uint32_t ph = 0, pph = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = pair + i * npairs;
      for (int kb = 0; kb < nkt; ++kb, ++it) {
        if (wgt) {
          if (kb < nk) {
            if (pi >= NP) {
              mbar_wait(&bar_bempty[ps], pph ^ 1, 16, ps);
              if (elect_one()) produce_packed(tile, kb, ps);
              __syncwarp();
            }
            ++pi;
            if (++ps == NP) { ps = 0; pph ^= 1; }
          }
        } else {
          if (it >= nstages) mbar_wait(&bar_empty[s], ph ^ 1, 1, s);
          if (elect_one()) {
            if (kb < nk) {
              if (leader) mbar_arrive_expect_tx(&bar_full[s], atom_tx);
              produce(tile, kb, s, true, false, false);
            } else {
              produce(tile, kb, s, true, true, true);
            }
          }
          __syncwarp();
          if (++s == nstages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 0 || warp == 3) {
    const bool wgt = warp == 3;
    fence_proxy_async_all();
    int it = 0, s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = pair + i * npairs;
      for (int kb = 0; kb < nkt; ++kb, ++it) {
        if (it >= n_pre) mbar_wait(&bar_empty[s], ph ^ 1, 1, s);
        if (elect_one()) {
          if (wgt) {
            if (it >= n_pre) produce(tile, kb, s, false, true, true);
          } else {
            produce(tile, kb, s, true, false, false);
          }
          if (p.trace && blockIdx.x == 0 && it < 48) p.trace[2048 + (wgt ? 256 : 512) + it] = globaltimer_ns();
        }
        __syncwarp();
        if (++s == nstages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      const uint32_t idesc_i8_1 = make_idesc_i8_rt(256, n1), idesc_i8_2 = make_idesc_i8_rt(256, n2 > 0 ? n2 : 32);
      const uint32_t b2_off32 = static_cast<uint32_t>((n1 >> 1) * 128) >> 4;   // chunk 2 starts n1/2 rows into the half
      const uint32_t lo_a0 = desc_lo_sw128(smem_u32(stage_a(0))), lo_b0 = desc_lo_sw128(smem_u32(stage_b(0)));
      const uint32_t stage_step = static_cast<uint32_t>(stage_bytes) >> 4;   // descriptor low-word steps: stage, k-atom
      const uint32_t a_step = Cfg::A_BYTES >> 4, b_step = b_atom >> 4;
      const bool two = n2 > 0;
      int s = 0;
      uint32_t ph = 0;
      // pass buffer b has been handed back (bar_tempty[b] completed) waited_b times so far
      uint32_t waited0 = 0, waited1 = 0;
      auto wait_consumed = [&](int G) {      // global pass G has been drained by every epilogue warp of the pair
        if (G < 0) return;
        const int b = G % NB;
        const uint32_t k = static_cast<uint32_t>(G / NB);
        uint32_t& waited = b ? waited1 : waited0;
        while (waited <= k) {
          mbar_wait(&bar_tempty[b], waited & 1, 2, b);
          ++waited;
        }
        tc_fence_after();
      };
      for (int i = 0; i < my_tiles; ++i) {
        const int G0 = i * P;                                       // first pass of this tile
        wait_consumed((i - SLOTS + 1) * P - 1);                     // the tile that last used this accumulator slot has left TMEM
        const uint32_t d1 = tmem_base + static_cast<uint32_t>((SLOTS == 2 ? (i & 1) : 0) * W), d2 = d1 + static_cast<uint32_t>(n1);
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&bar_full[s], ph, 4, s);
          if (w4) mbar_wait(&bar_ready[s], ph, 14, s);   // both CTAs' weight halves are unpacked
          tc_fence_after();
          const uint32_t lo_a = lo_a0 + static_cast<uint32_t>(s) * stage_step;
          const uint32_t lo_b = lo_b0 + static_cast<uint32_t>(s) * stage_step;
          if (elect_one()) {
            if (trace && i == 0 && kb == 0) trace[3] = globaltimer_ns();
            if (trace && blockIdx.x == 0 && i * nk + kb < 96) p.trace[2048 + i * nk + kb] = globaltimer_ns();
            if (!(p.ablate & 1)) {
              for (int a = 0; a < KA; ++a) {
                const uint32_t acc = (kb | a) != 0;
                if (two) umma_i8_2cta_atom2(d1, d2, lo_a + a * a_step, lo_b + a * b_step, lo_b + a * b_step + b2_off32, idesc_i8_1, idesc_i8_2, acc);
                else umma_i8_2cta_atom(d1, lo_a + a * a_step, lo_b + a * b_step, idesc_i8_1, acc);
              }
            }
            umma_commit_2cta(&bar_empty[s], 0x3);
          }
          __syncwarp();
          if (++s == nstages) { s = 0; ph ^= 1; }
        }
        if (trace && lane == 0 && i == 0) trace[6] = globaltimer_ns();   // int8 k-blocks of the first tile issued
        if (!has_o) {
          if (elect_one()) umma_commit_2cta(&bar_tfull[G0 % NB], 0x3);
          __syncwarp();
        } else {
          // the nko outlier k-blocks sit in the next nko stages and stay there for all P passes
          const int s_o = s;
          const uint32_t ph_o = ph;
          for (int kbo = 0; kbo < nko; ++kbo) {
            int so = s_o + kbo;
            uint32_t pho = ph_o;
            if (so >= nstages) { so -= nstages; pho ^= 1; }
            mbar_wait(&bar_full[so], pho, 4, so);
            if (w4) mbar_wait(&bar_ready[so], pho, 14, so);
          }
          tc_fence_after();
          for (int c = 0; c < P; ++c) {
            const int b = (G0 + c) % NB;
            wait_consumed(G0 + c - NB);   // the previous pass in this buffer was consumed
            const int nc = (W - c * R) < R ? (W - c * R) : R;
            const uint32_t idesc_f16 = make_idesc_f16_rt(256, nc);
            const uint32_t d_outl = tmem_base + outl_col0 + static_cast<uint32_t>(b * R);
            const uint64_t bo = static_cast<uint64_t>(c * (R >> 1) * 128) >> 4;
            if (elect_one()) {
              for (int kbo = 0; kbo < nko; ++kbo) {
                int so = s_o + kbo;
                if (so >= nstages) so -= nstages;
                const uint64_t da = make_sw128_kmajor_desc(smem_u32(stage_a(so)));
                const uint64_t db = make_sw128_kmajor_desc(smem_u32(stage_b(so))) + bo;
                int ksteps = (p.n_out - kbo * 64 + 15) / 16;
                if (ksteps > 4) ksteps = 4;
                for (int k = 0; k < ksteps; ++k)   // K = 16 fp16 = 32 B
                  umma_f16_2cta(d_outl, da + 2 * k, db + 2 * k, idesc_f16, (kbo | k) != 0);
              }
              umma_commit_2cta(&bar_tfull[b], 0x3);
              if (trace && blockIdx.x == 0 && i == 0 && c < 16) p.trace[1600 + c] = globaltimer_ns();
              if (c == P - 1)
                for (int kbo = 0; kbo < nko; ++kbo) {
                  int so = s_o + kbo;
                  if (so >= nstages) so -= nstages;
                  umma_commit_2cta(&bar_empty[so], 0x3);
                }
            }
            __syncwarp();
          }
          s += nko;
          if (s >= nstages) { s -= nstages; ph ^= 1; }
        }
      }
      if (trace && lane == 0) trace[4] = globaltimer_ns();
    }
  } else if (warp >= 4) {
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = (warp - 4) >> 2;       // which half of the tile's output columns
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t stage_sa = smem_u32(smem + Cfg::PIPE_BYTES + 256 + 512 + (warp - 4) * kEpiStageBytes);
    __half* s_scale = reinterpret_cast<__half*>(smem + Cfg::PIPE_BYTES + 256 + 512 + Cfg::EPI_WARPS * kEpiStageBytes) + half * 256;
    const uint32_t scale_sa = smem_u32(s_scale);
    const uint32_t scale_g_sa = scale_sa - static_cast<uint32_t>(half) * 512u, scale_u_sa = scale_g_sa + 512u;   // pair: gate | up scales
    const uint32_t tempty0 = mapa_u32(smem_u32(&bar_tempty[0]), 0), tempty1 = mapa_u32(smem_u32(&bar_tempty[1]), 0);
    const int h1 = n1 >> 1;                 // output columns of this half that live in MMA chunk 1
    const int mode = (p.outl != nullptr || p.bias != nullptr || p.act == 1) ? 2 : (p.residual != nullptr ? 1 : 0);
    // W4: these eight warps are idle while a tile's int8 k-blocks stream (one accumulator slot, no epilogue to overlap), so they
    // expand the packed nibbles: each stage's 64-byte weight rows (low nibble = even k, linear.py:14-18) become the
    // SWIZZLE_128B int8 tile the UMMA descriptor expects (row r: 16-byte chunk c at c ^ (r & 7)).  256 threads walk the
    // (row, 16-byte packed chunk) pairs of the stage; the leader's bar_ready collects both CTAs' warps.
    int us = 0, ups = 0, uit = 0;
    uint32_t uph = 0, upph = 0;
    const uint32_t ready_leader = mapa_u32(smem_u32(&bar_ready[0]), 0);
    for (int i = 0; i < my_tiles; ++i) {
      if (w4) {
        // Two groups of four warps take alternate items, so that one item's unpack (~0.5 us: generic-proxy shared-memory
        // traffic competes with the tensor core's operand reads) overlaps the next one's.  Every group still WAITS on every
        // item's barriers in order (a parity wait may not skip a phase); it unpacks and arrives only for its own items.
        const int ugrp = (warp - 4) >> 2;          // 0: warps 4-7, 1: warps 8-11
        const int ut = threadIdx.x - 128 - ugrp * 128;   // 0 .. 127 within the group
        for (int kb = 0; kb < nkt; ++kb, ++uit) {
          const bool mine = (uit & 1) == ugrp;
          const bool utr = p.trace != nullptr && blockIdx.x == 0 && warp == 4 && lane == 0 && uit < 40;
          if (utr) p.trace[1200 + 4 * uit] = globaltimer_ns();
          // the main stage is free again (its previous MMAs have retired) — also orders this item's bar_ready tick after the
          // MMA warp has consumed the stage's previous one
          if (uit >= nstages) mbar_wait_lane0(&bar_empty[us], uph ^ 1, 17, us);
          if (utr) p.trace[1200 + 4 * uit + 1] = globaltimer_ns();
          if (kb < nk) {
            mbar_wait_lane0(&bar_bfull[ups], upph, 15, ups);
            if (utr) p.trace[1200 + 4 * uit + 2] = globaltimer_ns();
            if (mine) {
              // 32-bit shared-window addresses + ld/st.shared (generic 64-bit pointers compile to LD.E / ST.E: slower path)
              const uint32_t src_sa = smem_u32(packed_slot(ups)), dst_sa = smem_u32(stage_b(us));
#pragma unroll 2
              for (int it = ut; it < bh * 4; it += 128) {
                const uint32_t r = static_cast<uint32_t>(it) >> 2, v = static_cast<uint32_t>(it) & 3u;
                const uint4 pk = lds128(src_sa + r * 64u + v * 16u);
                const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
                uint32_t o[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint32_t lo = nib_lo_s8x4(w[j]);
                  const uint32_t hi = nib_hi_s8x4(w[j]);
                  o[2 * j] = __byte_perm(lo, hi, 0x5140);
                  o[2 * j + 1] = __byte_perm(lo, hi, 0x7362);
                }
                const uint32_t drow = dst_sa + r * 128u;
                sts128(drow + (((2u * v) ^ (r & 7u)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
                sts128(drow + (((2u * v + 1u) ^ (r & 7u)) << 4), make_uint4(o[4], o[5], o[6], o[7]));
              }
              fence_proxy_async_smem();             // generic-proxy smem writes -> visible to the pair's tcgen05.mma (async proxy)
              __syncwarp();
            }
            if (lane == 0) mbar_arrive(&bar_bempty[ups]);     // the packed slot may be refilled once BOTH groups are past it
            if (++ups == NP) { ups = 0; upph ^= 1; }
            if (utr) p.trace[1200 + 4 * uit + 3] = globaltimer_ns();
          }
          // CTA-scope release, like the TMEM hand-back arrive of the epilogue: the unpacked tile sits in THIS CTA's shared memory
          // (fence.proxy.async above made it visible to the async proxy); a cluster-scope release costs ~0.8 us per arrive
          // (in-kernel trace, profiles/r02_trace_w4*) and paced the whole W4 mainloop at 1.5 us per k-block
          // (both groups arrive — the other one right after its waits — so that no barrier of the item can advance two phases
          // before a group has looked at it)
          if (lane == 0) mbar_arrive_cluster(ready_leader + static_cast<uint32_t>(us) * 8u);
          if (++us == nstages) { us = 0; uph ^= 1; }
        }
      }
      const int tile = pair + i * npairs;
      const int m0 = (tile % MP) * 256 + static_cast<int>(rank) * 128 + q * 32;   // first row of this warp
      const int n0 = pairm ? (tile / MP) * bh : (tile / MP) * W + half * bh;     // first output column of this warp
      const int row = m0 + lane;
      float xs = 0.f;
      if (p.epilogue == EPI_DEQUANT_F16 && row < p.M) xs = __half2float(p.x_scale[row]);
      // scale_col of this half -> smem, once per tile, by the first of the four warps that share it (SwiGLU pair: every warp
      // reads both halves' scales, so all eight warps meet)
      named_bar_sync(pairm ? 13 : 13 + half, pairm ? 256 : 128);
      if (q == 0 && p.epilogue == EPI_DEQUANT_F16)
        for (int j = lane * 8; j < bh; j += 256)
          *reinterpret_cast<uint4*>(s_scale + j) =
              (n0 + j < p.N) ? __ldg(reinterpret_cast<const uint4*>(((pairm && half == 1) ? p.scale_col2 : p.scale_col) + n0 + j))
                             : make_uint4(0, 0, 0, 0);
      named_bar_sync(pairm ? 13 : 13 + half, pairm ? 256 : 128);

      const uint32_t int_col0 = static_cast<uint32_t>((SLOTS == 2 ? (i & 1) : 0) * W);   // this tile's int32 accumulator slot
      for (int c = 0; c < P; ++c) {
        const int G = i * P + c;
        const int buf = G % NB;
        bool pre_staged = false;
        if (c == 0 && mode == 1 && !pairm && p.epilogue == EPI_DEQUANT_F16) {
          // the residual tile of the first 64 columns comes in while the accumulator is still being computed
          const int nc0 = W < R ? W : R;
          const int b0 = (nc0 >> 1) < h1 ? (nc0 >> 1) : h1;
          epi_stage_in(stage_sa, p.residual, p.ld_res, m0, n0, (b0 < 64 ? b0 : 64) >> 3, p.M, p.N, lane);
          __syncwarp();
          pre_staged = true;
        }
#ifdef MIXQ_EPI_TRACE
        if (trace && blockIdx.x == 0 && warp == 4 && lane == 0 && i == 0 && c < 16) p.trace[2000 + 2 * c] = static_cast<unsigned long long>(clock64());
#endif
        mbar_wait_warp(&bar_tfull[buf], (G / NB) & 1, 5, c);
        tc_fence_after();
#ifdef MIXQ_EPI_TRACE
        if (trace && blockIdx.x == 0 && warp == 4 && lane == 0 && i == 0 && c < 16) p.trace[2000 + 2 * c + 1] = static_cast<unsigned long long>(clock64());
#endif
        if (trace && !p.fused_prologue && warp == 4 && lane == 0 && i == 0 && c == 0) trace[7] = globaltimer_ns();
        if (trace && blockIdx.x == 0 && warp == 4 && lane == 0 && i == 0 && c < 16) p.trace[1536 + 2 * c] = globaltimer_ns();
        const int x0 = c * (R >> 1);                                   // this pass: output columns [x0, x1) of the half
        const int nc = (W - c * R) < R ? (W - c * R) : R;
        const int x1 = x0 + (nc >> 1);
        const uint32_t t_outl = lane_base + outl_col0 + static_cast<uint32_t>(buf * R + half * (nc >> 1));
        // split the run where the int32 accumulator switches from MMA chunk 1 to chunk 2
        for (int part = 0; part < 2; ++part) {
          const int a = part == 0 ? x0 : (x0 > h1 ? x0 : h1);
          const int b = part == 0 ? (x1 < h1 ? x1 : h1) : x1;
          if (a >= b) continue;
          const uint32_t t_int = lane_base + int_col0 + static_cast<uint32_t>(part == 0 ? half * h1 + a : n1 + half * (n2 >> 1) + (a - h1));
          const uint32_t t_o = t_outl + (a - x0);
          const uint32_t sc = scale_sa + a * 2;
          if (pairm) {
            // output columns [a, b) of the tile: gate accumulator | up accumulator (| their outlier accumulators of this pass)
            const uint32_t t_g = lane_base + int_col0 + static_cast<uint32_t>(part == 0 ? a : n1 + (a - h1));
            const uint32_t t_u = lane_base + int_col0 + static_cast<uint32_t>(part == 0 ? h1 + a : n1 + (n2 >> 1) + (a - h1));
            const uint32_t t_og = lane_base + outl_col0 + static_cast<uint32_t>(buf * R + (a - x0));
            const uint32_t t_ou = t_og + static_cast<uint32_t>(nc >> 1);
            if (has_o) epilogue_run_swiglu<true>(p, stage_sa, t_g, t_u, t_og, t_ou, m0, n0 + a, b - a, xs, scale_g_sa + a * 2, scale_u_sa + a * 2, lane, half);
            else epilogue_run_swiglu<false>(p, stage_sa, t_g, t_u, 0u, 0u, m0, n0 + a, b - a, xs, scale_g_sa + a * 2, scale_u_sa + a * 2, lane, half);
          } else if (p.epilogue == EPI_DEQUANT_F16) {
            if (has_o) {
              if (mode == 0) epilogue_run_coalesced<true, 0>(p, stage_sa, t_int, t_o, m0, n0 + a, b - a, xs, sc, lane);
              else if (mode == 1) epilogue_run_coalesced<true, 1>(p, stage_sa, t_int, t_o, m0, n0 + a, b - a, xs, sc, lane, pre_staged && part == 0);
              else epilogue_run_coalesced<true, 2>(p, stage_sa, t_int, t_o, m0, n0 + a, b - a, xs, sc, lane);
            } else {
              if (mode == 0) epilogue_run_coalesced<false, 0>(p, stage_sa, t_int, 0u, m0, n0 + a, b - a, xs, sc, lane);
              else if (mode == 1) epilogue_run_coalesced<false, 1>(p, stage_sa, t_int, 0u, m0, n0 + a, b - a, xs, sc, lane, pre_staged && part == 0);
              else epilogue_run_coalesced<false, 2>(p, stage_sa, t_int, 0u, m0, n0 + a, b - a, xs, sc, lane);
            }
          } else {   // raw int32 accumulators (mixlib.gemm)
            epilogue_span<false>(p, t_int, 0u, row, row < p.M, n0 + a, b - a, xs, nullptr, 0);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (trace && blockIdx.x == 0 && warp == 4 && lane == 0 && i == 0 && c < 16) p.trace[1536 + 2 * c + 1] = globaltimer_ns();
        if (lane == 0) mbar_arrive_cluster(buf ? tempty1 : tempty0);
      }
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

template __global__ void mixq_linear2_kernel<false>(const __grid_constant__ LinearParams);
template __global__ void mixq_linear2_kernel<true>(const __grid_constant__ LinearParams);

}  // namespace mixq
