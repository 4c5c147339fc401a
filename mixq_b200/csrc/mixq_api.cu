// extern "C" surface of libmixq_sm100 (see include/mixq.h).  Host-side only: argument checks,
// TMA tensor-map encoding, launch configuration.  No allocation, no synchronisation.
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>

#include "../../include/mixq.h"
#include "mixq_gemm.cuh"
#include "mixq_kernels.cuh"

namespace {

using namespace mixq;

thread_local std::string g_err;
std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_tile_n{0};
std::atomic<unsigned long long*> g_trace{nullptr};
// Programmatic dependent launch (ptx.cuh: pdl_wait / pdl_launch_dependents): on unless MIXQ_PDL=0 or mixq_set_pdl(0).
std::atomic<int> g_pdl{[] { const char* e = getenv("MIXQ_PDL"); return (e && e[0] == '0') ? 0 : 1; }()};
bool pdl_on() { return g_pdl.load(std::memory_order_relaxed) != 0; }
// How long the tensor-parallel exchange waits for a silent peer before it reports a stall (printf + trap).  Ranks legitimately
// drift apart by seconds around host-synchronising phases (outlier discovery, graph capture, rank-0-only work), so the default
// is generous; MIXQ_PEER_TIMEOUT_MS / mixq_set_peer_timeout_ms change it.
std::atomic<unsigned long long> g_peer_timeout_ms{[] {
  const char* e = getenv("MIXQ_PEER_TIMEOUT_MS");
  const long long v = e ? atoll(e) : 0;
  return static_cast<unsigned long long>(v > 0 ? v : 120000);
}()};

// How the kernels with a grid barrier (fused activation prologue) are launched — mixq_set_grid_barrier_mode / MIXQ_GRID_BARRIER:
//   0 "pdl"   (default) programmatic dependent launch, one CTA per SM: co-residency follows from the GPU being ours (every CTA of
//             the previous kernel leaves without waiting for anybody).  For a process that owns the GPU (the benchmark, a serving
//             worker with one stream of work).
//   1 "coop"  cudaLaunchAttributeCooperative instead of PDL on those launches: the driver guarantees co-residency (and refuses the
//             launch otherwise); the kernel boundary is no longer overlapped.
//   2 "split" no grid barrier at all: the activation prologue runs as its own launch (rowquant_kernel), then the GEMM with
//             skip_prologue — bit-identical results; for processes that share the GPU with other streams / MPS clients.
std::atomic<int> g_barrier_mode{[] {
  const char* e = getenv("MIXQ_GRID_BARRIER");
  if (e == nullptr) return 0;
  if (!strcmp(e, "coop")) return 1;
  if (!strcmp(e, "split")) return 2;
  return 0;
}()};
int barrier_mode() { return g_barrier_mode.load(std::memory_order_relaxed); }

// MIXQ_DEBUG_* tuning knobs: read once when the library loads (and again by mixq_reload_debug_env, for tests that change them),
// never on the launch path.
struct DebugEnv {
  std::atomic<int> katoms{0}, stage_bytes{0}, stages{0}, ablate{0}, splits{0};
  void load() {
    auto rd = [](const char* n) { const char* e = getenv(n); return e ? atoi(e) : 0; };
    katoms.store(rd("MIXQ_DEBUG_KATOMS"));
    stage_bytes.store(rd("MIXQ_DEBUG_STAGE_BYTES"));
    stages.store(rd("MIXQ_DEBUG_STAGES"));
    ablate.store(rd("MIXQ_DEBUG_ABLATE"));
    splits.store(rd("MIXQ_DEBUG_SPLITS"));     // cap of the split-K factor (1 = never split)
  }
  DebugEnv() { load(); }
};
DebugEnv g_dbg;

// Launch attributes shared by every kernel of the library.  A kernel with a grid barrier needs all of its CTAs
// co-resident: a cooperative launch guarantees it; under PDL the grid is exactly one CTA per SM and the CTAs of the
// previous kernel leave without waiting for anybody, so ours all become resident as those exit.
struct LaunchAttrs {
  cudaLaunchAttribute a[2];
  int n = 0;
  LaunchAttrs(bool cooperative, bool pdl) {
    if (cooperative && barrier_mode() == 1) pdl = false;
    if (pdl) {
      a[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      a[n].val.programmaticStreamSerializationAllowed = 1;
      ++n;
    } else if (cooperative) {
      a[n].id = cudaLaunchAttributeCooperative;
      a[n].val.cooperative = 1;
      ++n;
    }
  }
};
template <typename... KArgs, typename... Args>
cudaError_t launch_small(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  LaunchAttrs la(false, pdl_on());
  cfg.attrs = la.a;
  cfg.numAttrs = la.n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* where) {
  g_err = std::string(where) + ": " + cudaGetErrorString(e);
  return static_cast<int>(e);
}
#define MIXQ_CUDA(call)                                   \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

struct DeviceInfo {
  int sms = 0;
  int cc_major = 0;
  bool ok = false;
};
int device_info(DeviceInfo* out) {
  static std::mutex mu;
  static std::unordered_map<int, DeviceInfo> cache;
  int dev = 0;
  MIXQ_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(dev);
  if (it == cache.end()) {
    DeviceInfo d;
    MIXQ_CUDA(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
    MIXQ_CUDA(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    d.ok = true;
    it = cache.emplace(dev, d).first;
  }
  *out = it->second;
  if (out->cc_major != 10) return fail(MIXQ_EARCH, "libmixq_sm100 needs a compute-capability 10.x device (B200)");
  return 0;
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// Encoded tensor maps are pure functions of (base address, geometry): a module calling with the same buffers gets its maps from
// this cache instead of 4-7 cuTensorMapEncodeTiled calls per launch (the eager call pattern of benchbitsand.py; under a CUDA
// graph the launch path does not run at all).
struct MapKey {
  uint64_t v[8];
  bool operator==(const MapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (uint64_t x : k.v) h = (h ^ x) * 0xFF51AFD7ED558CCDull + (h >> 29);
    return static_cast<size_t>(h);
  }
};
std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
constexpr size_t kMapCacheCap = 8192;      // ~1 MB; cleared when full (a 70B model has 80 x 4 modules x <= 7 maps)
bool map_cache_get(const MapKey& k, CUtensorMap* m) {
  std::lock_guard<std::mutex> lk(g_map_mu);
  auto it = g_maps.find(k);
  if (it == g_maps.end()) return false;
  *m = it->second;
  return true;
}
void map_cache_put(const MapKey& k, const CUtensorMap& m) {
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (g_maps.size() >= kMapCacheCap) g_maps.clear();
  g_maps.emplace(k, m);
}

// 2-D row-major tensor [rows, cols] of `elt` bytes, row pitch `pitch_bytes`; box = box_cols x box_rows.
int make_map(CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int elt, long long cols, long long rows,
             long long pitch_bytes, int box_cols, int box_rows, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(MIXQ_EDRIVER, "cuTensorMapEncodeTiled not available from the driver");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(MIXQ_EINVAL, "TMA operand not 16-byte aligned");
  if (pitch_bytes % 16 != 0) return fail(MIXQ_EINVAL, "TMA operand row pitch not a multiple of 16 bytes");
  const MapKey key{{reinterpret_cast<uintptr_t>(ptr), static_cast<uint64_t>(dt) | (static_cast<uint64_t>(sw) << 32), static_cast<uint64_t>(cols),
                    static_cast<uint64_t>(rows), static_cast<uint64_t>(pitch_bytes), static_cast<uint64_t>(box_cols),
                    static_cast<uint64_t>(box_rows), 2}};
  if (map_cache_get(key, m)) return 0;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(pitch_bytes)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  (void)elt;
  CUresult r = fn(m, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MIXQ_EDRIVER, "cuTensorMapEncodeTiled failed (CUresult " + std::to_string(r) + ")");
  map_cache_put(key, *m);
  return 0;
}

// The [byte in k-atom (128), row, k-atom] view of a row-major int8 matrix [rows, K] (K % 128 == 0): a box of
// 128 B x box_rows x atoms lands in shared memory as `atoms` consecutive SWIZZLE_128B tiles of box_rows x 128 B.
int make_map_katoms(CUtensorMap* m, const void* ptr, long long K, long long rows, int box_rows, int atoms) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(MIXQ_EDRIVER, "cuTensorMapEncodeTiled not available from the driver");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(MIXQ_EINVAL, "TMA operand not 16-byte aligned");
  const MapKey key{{reinterpret_cast<uintptr_t>(ptr), static_cast<uint64_t>(K), static_cast<uint64_t>(rows), static_cast<uint64_t>(box_rows),
                    static_cast<uint64_t>(atoms), 0, 0, 3}};
  if (map_cache_get(key, m)) return 0;
  cuuint64_t gdim[3] = {128, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(K / 128)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(K), 128};
  cuuint32_t box[3] = {128, static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(atoms)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MIXQ_EDRIVER, "cuTensorMapEncodeTiled (3-D) failed (CUresult " + std::to_string(r) + ")");
  map_cache_put(key, *m);
  return 0;
}

template <int BN, bool W4>
int launch_linear(const LinearParams& p, int grid, bool cooperative, cudaStream_t st) {
  using Cfg = GemmCfg<BN, W4>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(mixq_linear_kernel<BN, W4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::SMEM_BYTES);
  });
  // the attribute is per device; set it again cheaply if we are on another device
  static thread_local int last_dev = -1;
  int dev = 0;
  MIXQ_CUDA(cudaGetDevice(&dev));
  if (dev != last_dev) {
    MIXQ_CUDA(cudaFuncSetAttribute(mixq_linear_kernel<BN, W4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg::SMEM_BYTES));
    last_dev = dev;
  }
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(smem)");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  LaunchAttrs la(cooperative, pdl_on());
  cfg.attrs = la.a;
  cfg.numAttrs = la.n;
  MIXQ_CUDA(cudaLaunchKernelEx(&cfg, mixq_linear_kernel<BN, W4>, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

template <bool W4>
int launch_linear2_t(const LinearParams& p, int grid, bool cooperative, cudaStream_t st) {
  using Cfg = Gemm2Cfg;
  static thread_local int last_dev = -1;
  static thread_local int max_clusters = 0;
  int dev = 0;
  MIXQ_CUDA(cudaGetDevice(&dev));
  if (dev != last_dev) {
    MIXQ_CUDA(cudaFuncSetAttribute(mixq_linear2_kernel<W4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    // the grid barrier needs every CTA pair resident at once: ask the driver how many clusters of this kernel fit the device
    cudaLaunchConfig_t q{};
    q.gridDim = dim3(2);
    q.blockDim = dim3(Cfg::NUM_THREADS);
    q.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = 2;
    qa[0].val.clusterDim.y = qa[0].val.clusterDim.z = 1;
    q.attrs = qa;
    q.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, mixq_linear2_kernel<W4>, &q) != cudaSuccess) {
      (void)cudaGetLastError();
      max_clusters = 0;   // unknown: do not block the launch on a failed query
    }
    last_dev = dev;
  }
  if (cooperative && max_clusters > 0 && grid / 2 > max_clusters)
    return fail(MIXQ_EINVAL, "fused prologue: the device cannot hold all " + std::to_string(grid / 2) +
                                 " CTA pairs of the grid barrier at once (max " + std::to_string(max_clusters) +
                                 "); use mixq_set_grid_barrier_mode(2)");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  LaunchAttrs la(cooperative, pdl_on());
  cfg.attrs = la.a;
  cfg.numAttrs = la.n;
  MIXQ_CUDA(cudaLaunchKernelEx(&cfg, mixq_linear2_kernel<W4>, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int launch_linear2(const LinearParams& p, int grid, bool cooperative, cudaStream_t st) {
  return p.w4 ? launch_linear2_t<true>(p, grid, cooperative, st) : launch_linear2_t<false>(p, grid, cooperative, st);
}

// Tile width of the 2-CTA kernel.  The widest strip that still gives every pair one tile: the `npairs / MP` pairs that
// share a 256-row block split the N columns evenly; if that is more than one TMEM-full (512 columns) they take several
// tiles each.  Wide tiles minimise the bytes each SM must land per MMA cycle (8192/W + 32).
int pick_w2(int requested, int M, int N, int n_out, int npairs) {
  const int forced = g_tile_n.load(std::memory_order_relaxed);
  int fixed = 0;
  if (requested >= 32 && requested <= 512 && requested % 32 == 0) fixed = requested;
  else if (forced >= 32 && forced <= 512 && forced % 32 == 0) fixed = forced;
  if (fixed) {
    const int nko_f = (n_out + 63) / 64;
    if (nko_f > 2 && fixed > 256) fixed = 256;
    if (nko_f > 0 && fixed > 448) fixed = 448;
    return fixed;
  }
  const int mp = (M + 255) / 256;
  int strips = npairs / mp;
  if (strips < 1) strips = 1;
  int w = ((N + strips - 1) / strips + 31) / 32 * 32;
  if (w > 512) {
    const int t = (w + 511) / 512;
    w = ((N + strips * t - 1) / (strips * t) + 31) / 32 * 32;
  }
  if (w < 128) w = 128;   // an MMA of N <= 128 costs the same 75 cycles (tools/mma_bw.cu): narrower tiles only add traffic
  // outlier k-blocks stay resident in their stages while the epilogue passes run: at most 2 of them, else the whole
  // fp32 outlier accumulator must fit beside the int32 one (W <= 256, a single pass, blocks stream normally)
  const int nko = (n_out + 63) / 64;
  if (nko > 2 && w > 256) w = 256;
  if (nko > 0 && w > 448) w = 448;
  return w;
}

int pick_tile_n(int requested, int M, int N, int sms, bool tmem_outliers) {
  if (requested == 128 || requested == 256) return requested;
  const int forced = g_tile_n.load(std::memory_order_relaxed);
  if (forced == 128 || forced == 256) return forced;
  // Fewest "waves x tile width" wins; 256-wide tiles lose accumulator double-buffering when the
  // fp32 outlier accumulator also lives in TMEM.
  const int mb = (M + 127) / 128;
  auto cost = [&](int bn) {
    const int tiles = mb * ((N + bn - 1) / bn);
    const int waves = (tiles + sms - 1) / sms;
    double c = static_cast<double>(waves) * bn;
    if (bn == 256 && tmem_outliers) c *= 1.15;
    if (bn == 128) c *= 1.08;  // narrower tile: more smem traffic per MMA
    return c;
  };
  return cost(256) < cost(128) ? 256 : 128;
}

int pick_row_groups(RowQuantArgs* a, int grid, int warps_per_cta, long long smem_budget);

// Launch plan of one MixLinear GEMM: which kernel, tile width, k-atoms per TMA op, pipeline depth, TMEM plan.  Pure host
// arithmetic (exported as mixq_plan_linear so that the heuristics are testable without a GPU).
struct GemmPlan {
  int two_cta, tile_w, k_atoms, stage_bytes, nstages, tiles, tiles_per_unit, units;
  int splits = 1;    // 1-CTA kernel: CTAs per tile along K (needs a workspace)
  int npacked = 0;   // W4 on the 2-CTA kernel: slots of the packed-row ring behind the main stages
  TmemPlan tmem;
};
constexpr long long kSplitKCounterBytes = 4096;     // per-tile counters at the head of the split-K workspace
int plan_gemm(int M, int N, int K, int bit, int n_out, bool pair, int tile_req, int sms, GemmPlan* g, long long splitk_ws_bytes = 0) {
  if (M < 1 || N < 8 || K < 16 || sms < 1) return fail(MIXQ_EINVAL, "M>=1, N>=8, K>=16 required");
  const bool w4 = bit == 4;
  // M > 128: CTA pairs (cta_group::2) own 256 x bn tiles — half the L2 -> SM bytes per MMA cycle (mixq_gemm2.cu)
  bool two_cta = M > 128 && sms >= 2;   // W4 too: the epilogue warps unpack the nibbles during the mainloop (mixq_gemm2.cu)
  const int npairs = sms / 2;
  if (pair && (!two_cta || N % 16 != 0)) return fail(MIXQ_EINVAL, "SwiGLU pair needs M > 128 and N % 16 == 0");
  // pair: a tile of width W holds W/2 gate columns (staged by CTA 0) and the SAME W/2 up columns (staged by CTA 1)
  int bn = two_cta ? pick_w2(tile_req, M, pair ? 2 * N : N, n_out, npairs) : pick_tile_n(tile_req, M, N, sms, n_out > 0);
  // narrow tiles are paced by the TMA op count (one op ~340 clocks of the SM's TMA unit whatever its size): two k-atoms
  // (256 bytes of K) per op and pipeline stage whenever at least three such stages fit
  g->splits = 1;
  if (!two_cta && splitk_ws_bytes > kSplitKCounterBytes && tile_req == 0 && g_tile_n.load(std::memory_order_relaxed) == 0) {
    // few 128-row tiles: one SM lands ~63 B/clk, a 128 x 128 tile's k-block (32 KB) costs ~0.27 us whatever else idles.
    // Split K over S <= 4 CTAs per tile (single wave: t128 * S <= SMs; every split keeps >= 4 k-blocks) when the model
    //   t(S) = nk / S * 0.27 us + (S > 1 ? 5.0 + 1.0 (S - 1) : 0)      [partial store + counter + bulk copies + fold]
    // says it pays (constants from profiles/r02_trace_splitk_v2.log: 4096 x 4096 at M = 32 must NOT split, 17.8 vs 19.9 us).
    const int t128 = ((M + 127) / 128) * ((N + 127) / 128);
    const int nk = (K + 127) / 128;
    int smax = sms / t128;
    if (smax > 4) smax = 4;
    if (smax > nk / 4) smax = nk / 4;
    if (const int cap = g_dbg.splits.load(std::memory_order_relaxed); cap >= 1 && smax > cap) smax = cap;   // MIXQ_DEBUG_SPLITS
    while (smax > 1 && kSplitKCounterBytes + static_cast<long long>(t128) * (smax - 1) * 65536 > splitk_ws_bytes) --smax;
    int S = 1;
    double best = nk * 0.27;
    for (int c = 2; c <= smax; ++c) {
      const double t = static_cast<double>((nk + c - 1) / c) * 0.27 + 5.0 + 1.0 * (c - 1);
      if (t < best - 0.5) { best = t; S = c; }
    }
    if (g_dbg.splits.load(std::memory_order_relaxed) >= 2 && smax >= 2) S = smax;      // forced (tests, tuning)
    if (S >= 2 && t128 * 4 <= kSplitKCounterBytes) {
      bn = 128;
      g->splits = S;
    }
  }
  int k_atoms = (two_cta && !w4 && K % 128 == 0 && K >= 256 && bn <= 256) ? 2 : 1;
  if (g_dbg.katoms.load(std::memory_order_relaxed) == 1) k_atoms = 1;
  int stage2 = 0, nstages2 = 0;
  for (;;) {
    stage2 = k_atoms * (Gemm2Cfg::A_BYTES + (bn / 2) * 128);
    stage2 = (stage2 + 1023) / 1024 * 1024;
    if (const int v = g_dbg.stage_bytes.load(std::memory_order_relaxed); v >= stage2 && v % 1024 == 0) stage2 = v;
    nstages2 = Gemm2Cfg::PIPE_BYTES / stage2;
    if (nstages2 > Gemm2Cfg::MAX_STAGES) nstages2 = Gemm2Cfg::MAX_STAGES;
    if (w4 && two_cta && nstages2 > 3) nstages2 = 3;   // W4: three main stages cover the unpack -> MMA -> commit chain; the rest of
                                                      // the pipeline memory is the packed-row ring that covers HBM latency
    if (const int v = g_dbg.stages.load(std::memory_order_relaxed); v >= 2 && v < nstages2) nstages2 = v;
    // the outlier k-blocks of a tile stay resident in the ring during the epilogue passes: with the big stages they may not fit
    if (k_atoms == 2 && (nstages2 < 3 || (n_out + 63) / 64 > nstages2 - 1)) { k_atoms = 1; continue; }
    break;
  }
  // the outlier k-blocks of a tile stay resident in the ring during the epilogue passes: they must all fit
  if (two_cta && (n_out + 63) / 64 > nstages2 - 1) {
    if (pair) return fail(MIXQ_EINVAL, "SwiGLU pair: too many outlier columns for the resident outlier stages");
    two_cta = false;
    bn = pick_tile_n(tile_req, M, N, sms, n_out > 0);
  }
  g->two_cta = two_cta ? 1 : 0;
  g->npacked = 0;
  if (two_cta && w4) {
    const int pitch = ((bn / 2) * 64 + 127) / 128 * 128;
    int np = (Gemm2Cfg::PIPE_BYTES - nstages2 * stage2) / pitch;
    if (np > Gemm2Cfg::MAX_STAGES) np = Gemm2Cfg::MAX_STAGES;
    if (np < 2) return fail(MIXQ_EINVAL, "W4: no room for the packed-row ring at this tile width");
    g->npacked = np;
  }
  g->tile_w = bn;
  g->k_atoms = two_cta ? k_atoms : 1;
  if (two_cta) {
    g->stage_bytes = stage2;
    g->nstages = nstages2;
    const int wout = pair ? bn / 2 : bn;
    g->tiles = ((M + 255) / 256) * ((N + wout - 1) / wout);
    g->units = npairs;
    g->tiles_per_unit = (g->tiles + npairs - 1) / npairs;
    const bool single = (g_dbg.ablate.load(std::memory_order_relaxed) & 16) != 0;
    g->tmem = plan_tmem(bn, n_out > 0, w4 ? 1 : g->tiles_per_unit, single);
  } else {
    g->stage_bytes = w4 ? (bn == 256 ? GemmCfg<256, true>::STAGE_BYTES : GemmCfg<128, true>::STAGE_BYTES)
                        : (bn == 256 ? GemmCfg<256, false>::STAGE_BYTES : GemmCfg<128, false>::STAGE_BYTES);
    g->nstages = w4 ? (bn == 256 ? GemmCfg<256, true>::STAGES : GemmCfg<128, true>::STAGES)
                    : (bn == 256 ? GemmCfg<256, false>::STAGES : GemmCfg<128, false>::STAGES);
    g->tiles = ((M + 127) / 128) * ((N + bn - 1) / bn);
    g->units = sms;
    g->tiles_per_unit = (g->tiles * g->splits + sms - 1) / sms;
    const int acc_cols = bn * (n_out > 0 ? 2 : 1);       // mixq_gemm.cu: s32 accumulator [+ f32 outlier accumulator]
    g->tmem.slots = (2 * acc_cols <= 512) ? 2 : 1;
    g->tmem.passes = 1;
    g->tmem.pass_cols = bn;
    g->tmem.buffers = 1;
  }
  return 0;
}

struct GemmCall {
  const void* q_x;
  const void* q_w;
  const void* q_w_up = nullptr;        // SwiGLU pair
  const void* scale_col_up = nullptr;
  const void* weight_cache_up = nullptr;
  int bit;
  const void* x_scale;
  const void* scale_col;
  const void* bias;
  const void* outl;
  int ld_outl;
  const void* residual;
  int ld_res;
  const void* act_outliers;
  int ld_ao;
  const void* weight_cache;
  int ld_wc;
  int n_out;
  void* y;
  int32_t* y_i32;
  int M, N, K, act, epilogue, tile_n;
  const RowQuantArgs* rq;  // non-null => fused prologue
  uint32_t* grid_sync;
  void* const* y_peer = nullptr;   // tensor-parallel push (see mixq_linear_args.y_peer)
  int peer_cols = 0;
  int peer_bcast = 0;
  void* splitk_ws = nullptr;
  long long splitk_ws_bytes = 0;
};

int run_gemm(const GemmCall& c, cudaStream_t st) {
  DeviceInfo di;
  if (int r = device_info(&di)) return r;
  if (c.M < 1 || c.N < 8 || c.K < 16) return fail(MIXQ_EINVAL, "M>=1, N>=8, K>=16 required");
  if (c.K % 16 != 0 || c.N % 8 != 0) return fail(MIXQ_EINVAL, "K % 16 == 0 and N % 8 == 0 required");
  if (c.bit != 8 && c.bit != 4) return fail(MIXQ_EINVAL, "bit must be 8 or 4");
  if (c.bit == 4 && c.K % 32 != 0) return fail(MIXQ_EINVAL, "bit 4 needs K % 32 == 0");
  if (c.n_out > 0 && (c.ld_ao % 8 != 0 || c.ld_wc % 8 != 0 || c.ld_ao < c.n_out || c.ld_wc < c.n_out))
    return fail(MIXQ_EINVAL, "outlier buffers need ld % 8 == 0 and ld >= n_ind");
  const bool w4 = (c.bit == 4);
  const bool pair = c.q_w_up != nullptr;
  GemmPlan gp{};
  // split-K only for the plain dequant epilogue into y (raw int32 output and the tensor-parallel pushes keep one CTA per tile)
  const bool can_split = c.splitk_ws != nullptr && (reinterpret_cast<uintptr_t>(c.splitk_ws) & 15) == 0 && c.epilogue == EPI_DEQUANT_F16 &&
                         c.peer_cols <= 0 && c.peer_bcast <= 0 && c.N % 4 == 0;
  if (int r = plan_gemm(c.M, c.N, c.K, c.bit, c.n_out, pair, c.tile_n, di.sms, &gp, can_split ? c.splitk_ws_bytes : 0)) return r;
  if (c.peer_cols > 0) {
    // tensor-parallel push: a tile must not straddle two ranks' column slices
    if (pair || c.bias || c.outl || c.residual || c.epilogue != EPI_DEQUANT_F16 || c.y_peer == nullptr)
      return fail(MIXQ_EINVAL, "tensor-parallel push: no bias / residual / addend / SwiGLU pair, fp16 output only");
    if (c.peer_cols % 128 != 0 || c.N % c.peer_cols != 0 || c.N / c.peer_cols > 8)
      return fail(MIXQ_EINVAL, "tensor-parallel push: peer_cols must be a multiple of 128 dividing N into at most 8 slices");
    if (c.peer_cols % gp.tile_w != 0) {
      if (int r = plan_gemm(c.M, c.N, c.K, c.bit, c.n_out, pair, 128, di.sms, &gp)) return r;
      if (c.peer_cols % gp.tile_w != 0) return fail(MIXQ_EINVAL, "tensor-parallel push: no tile width divides peer_cols");
    }
    for (int j = 0; j < c.N / c.peer_cols; ++j)
      if (c.y_peer[j] == nullptr || (reinterpret_cast<uintptr_t>(c.y_peer[j]) & 15) != 0)
        return fail(MIXQ_EINVAL, "tensor-parallel push: missing or misaligned y_peer pointer");
  }
  if (c.peer_bcast > 0) {
    if (c.peer_cols > 0 || c.peer_bcast > 8 || pair || c.bias || c.outl || c.residual || c.epilogue != EPI_DEQUANT_F16 || c.y_peer == nullptr)
      return fail(MIXQ_EINVAL, "tensor-parallel broadcast push: 1..8 destinations, no bias / residual / addend / SwiGLU pair");
    for (int j = 0; j < c.peer_bcast; ++j)
      if (c.y_peer[j] == nullptr || (reinterpret_cast<uintptr_t>(c.y_peer[j]) & 15) != 0)
        return fail(MIXQ_EINVAL, "tensor-parallel push: missing or misaligned y_peer pointer");
  }
  const bool two_cta = gp.two_cta != 0;
  const int npairs = di.sms / 2;
  if (pair && (c.bias != nullptr || !c.scale_col_up || (c.n_out > 0 && !c.weight_cache_up) || c.epilogue != EPI_DEQUANT_F16 ||
               c.outl != nullptr || c.residual != nullptr))
    return fail(MIXQ_EINVAL, "SwiGLU pair needs both scale/weight_cache sets and no bias/residual/addend");
  const int bn = gp.tile_w, k_atoms = gp.k_atoms, stage2 = gp.stage_bytes, nstages2 = gp.nstages;
  const int b_rows = two_cta ? bn / 2 : bn;

  LinearParams p{};
  const bool ka2 = two_cta && k_atoms == 2;
  if (ka2) {
    if (int r = make_map_katoms(&p.tm_a, c.q_x, c.K, c.M, 128, 2)) return r;
    if (int r = make_map_katoms(&p.tm_b, c.q_w, c.K, c.N, b_rows, 2)) return r;
  } else if (int r = make_map(&p.tm_a, c.q_x, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, c.K, c.M, c.K, 128, 128,
                              CU_TENSOR_MAP_SWIZZLE_128B)) {
    return r;
  }
  if (ka2) {
  } else if (w4) {
    if (int r = make_map(&p.tm_b, c.q_w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, c.K / 2, c.N, c.K / 2, 64, b_rows,
                         CU_TENSOR_MAP_SWIZZLE_NONE))
      return r;
  } else {
    if (int r = make_map(&p.tm_b, c.q_w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, c.K, c.N, c.K, 128, b_rows,
                         CU_TENSOR_MAP_SWIZZLE_128B))
      return r;
  }
  if (c.n_out > 0) {
    if (int r = make_map(&p.tm_oa, c.act_outliers, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, c.n_out, c.M,
                         static_cast<long long>(c.ld_ao) * 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B))
      return r;
    if (int r = make_map(&p.tm_ob, c.weight_cache, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, c.n_out, c.N,
                         static_cast<long long>(c.ld_wc) * 2, 64, b_rows, CU_TENSOR_MAP_SWIZZLE_128B))
      return r;
  }
  if (pair) {
    if (ka2) {
      if (int r = make_map_katoms(&p.tm_b2, c.q_w_up, c.K, c.N, b_rows, 2)) return r;
    } else if (w4) {
      if (int r = make_map(&p.tm_b2, c.q_w_up, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, c.K / 2, c.N, c.K / 2, 64, b_rows,
                           CU_TENSOR_MAP_SWIZZLE_NONE))
        return r;
    } else if (int r = make_map(&p.tm_b2, c.q_w_up, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, c.K, c.N, c.K, 128, b_rows,
                                CU_TENSOR_MAP_SWIZZLE_128B)) {
      return r;
    }
    if (c.n_out > 0)
      if (int r = make_map(&p.tm_ob2, c.weight_cache_up, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, c.n_out, c.N,
                           static_cast<long long>(c.ld_wc) * 2, 64, b_rows, CU_TENSOR_MAP_SWIZZLE_128B))
        return r;
    p.scale_col2 = static_cast<const __half*>(c.scale_col_up);
    p.pair_swiglu = 1;
  }
  if (c.rq != nullptr) {
    p.rq = *c.rq;
    p.fused_prologue = 1;
    // phase A runs on the persistent grid (one CTA per SM) with every warp of the GEMM CTA
    {
      // row buffers live in the pipeline stages the weight prefetch does not use yet: all but the first stage
      const long long stage = w4 ? (bn == 256 ? GemmCfg<256, true>::STAGE_BYTES : GemmCfg<128, true>::STAGE_BYTES)
                                 : (bn == 256 ? GemmCfg<256, false>::STAGE_BYTES : GemmCfg<128, false>::STAGE_BYTES);
      const int stages = w4 ? (bn == 256 ? GemmCfg<256, true>::STAGES : GemmCfg<128, true>::STAGES)
                            : (bn == 256 ? GemmCfg<256, false>::STAGES : GemmCfg<128, false>::STAGES);
      if (two_cta) {
        if (int r = pick_row_groups(&p.rq, npairs * 2, Gemm2Cfg::NUM_THREADS / 32,
                                    static_cast<long long>(stage2) * (nstages2 - 1)))
          return r;
      } else if (int r = pick_row_groups(&p.rq, di.sms, w4 ? 12 : 8, stage * (stages - 1))) {
        return r;
      }
    }
    p.rq.trace = g_trace.load(std::memory_order_relaxed);
  }
  p.x_scale = static_cast<const __half*>(c.x_scale);
  p.scale_col = static_cast<const __half*>(c.scale_col);
  p.bias = static_cast<const __half*>(c.bias);
  p.outl = static_cast<const __half*>(c.outl);
  p.ld_outl = c.ld_outl;
  p.residual = static_cast<const __half*>(c.residual);
  p.ld_res = c.ld_res;
  p.y = static_cast<__half*>(c.y);
  p.peer_cols = c.peer_cols;
  p.peer_bcast = c.peer_bcast;
  p.splits = two_cta ? 1 : gp.splits;
  p.sk_cnt = static_cast<uint32_t*>(c.splitk_ws);
  p.sk_ws = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(c.splitk_ws) + kSplitKCounterBytes);
  if (c.peer_cols > 0)
    for (int j = 0; j < c.N / c.peer_cols; ++j) p.y_peer[j] = static_cast<__half*>(c.y_peer[j]);
  for (int j = 0; j < c.peer_bcast; ++j) p.y_peer[j] = static_cast<__half*>(c.y_peer[j]);
  p.y_i32 = c.y_i32;
  p.M = c.M;
  p.N = c.N;
  p.K = c.K;
  p.n_out = c.n_out;
  p.act = c.act;
  p.epilogue = c.epilogue;
  p.grid_sync = c.grid_sync;
  p.trace = g_trace.load(std::memory_order_relaxed);

  p.bn = bn;
  p.nstages = nstages2;
  p.stage_bytes = stage2;
  p.k_atoms = ka2 ? 2 : 1;
  p.w4 = (w4 && two_cta) ? 1 : 0;
  p.npacked = gp.npacked;
  p.ablate = g_dbg.ablate.load(std::memory_order_relaxed);
  p.q_w = static_cast<const uint8_t*>(c.q_w);
  p.q_w_pitch = w4 ? c.K / 2 : c.K;
  if (two_cta) {
    const int tiles2 = ((c.M + 255) / 256) * (((pair ? 2 * c.N : c.N) + bn - 1) / bn);
    const bool coop2 = p.fused_prologue != 0;
    const int grid2 = coop2 ? npairs * 2 : 2 * (tiles2 < npairs ? tiles2 : npairs);
    return launch_linear2(p, grid2, coop2, st);
  }
  const int tiles = ((c.M + 127) / 128) * ((c.N + bn - 1) / bn) * p.splits;    // work units
  const bool coop = p.fused_prologue != 0;
  const int grid = coop ? di.sms : (tiles < di.sms ? tiles : di.sms);
  if (w4) return bn == 256 ? launch_linear<256, true>(p, grid, coop, st) : launch_linear<128, true>(p, grid, coop, st);
  return bn == 256 ? launch_linear<256, false>(p, grid, coop, st) : launch_linear<128, false>(p, grid, coop, st);
}

int grid_for(long long work_items, int threads, int sms, int per_sm = 8) {
  long long blocks = (work_items + threads - 1) / threads;
  long long cap = static_cast<long long>(sms) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// Work split of the activation prologue: `ngroups` rows in flight per CTA (each needs K*2 bytes of shared memory),
// G = warps / ngroups warps on each.  All rows of the batch in flight at once when possible: a 4 MB pass is bounded
// by latency, not bandwidth.
int pick_row_groups(RowQuantArgs* a, int grid, int warps_per_cta, long long smem_budget) {
  if (a->K > kRowQuantMaxK) return fail(MIXQ_EINVAL, "activation prologue supports K <= 32768");
  const long long row_bytes = static_cast<long long>(a->K) * 2;
  const int rows_per_cta = (a->M + grid - 1) / grid;
  long long ng = rows_per_cta;
  if (ng > warps_per_cta) ng = warps_per_cta;
  if (ng > kRowQuantMaxGroups) ng = kRowQuantMaxGroups;
  if (ng > smem_budget / row_bytes) ng = smem_budget / row_bytes;
  if (ng < 1) return fail(MIXQ_EINVAL, "activation row does not fit the shared-memory row buffer");
  a->ngroups = static_cast<int>(ng);
  a->group_warps = warps_per_cta / a->ngroups;
  return 0;
}

int fill_rowquant(RowQuantArgs* a, void* x, const void* norm_w, void* norm_out, float eps, const int32_t* ind,
                  int n_ind, void* act_out, int ld_ao, void* q_x, void* x_scale, int M, int K, int bit, float sigma,
                  uint8_t* col_over, uint32_t* over_flag) {
  if (M < 1 || K < 8 || K % 8 != 0) return fail(MIXQ_EINVAL, "M >= 1 and K % 8 == 0 required");
  if (q_x != nullptr && bit != 8 && bit != 4) return fail(MIXQ_EINVAL, "bit must be 8 or 4");
  if (n_ind > 0 && (ind == nullptr || act_out == nullptr || ld_ao < n_ind))
    return fail(MIXQ_EINVAL, "outlier gather needs ind, act_out and ld_ao >= n_ind");
  a->x = static_cast<__half*>(x);
  a->norm_w = static_cast<const __half*>(norm_w);
  a->norm_out = static_cast<__half*>(norm_out);
  a->eps = eps;
  a->ind = ind;
  a->n_ind = n_ind;
  a->act_out = static_cast<__half*>(act_out);
  a->ld_ao = ld_ao;
  a->q_x = static_cast<int8_t*>(q_x);
  a->x_scale = static_cast<__half*>(x_scale);
  a->M = M;
  a->K = K;
  a->bit = bit;
  const float qmax = (bit == 4) ? 7.f : 127.f;
  a->sigma = __float2half_rn(sigma);
  // torch: fp16 tensor / python int -> computed in fp32, rounded to fp16 (linear.py:201)
  a->thr = __float2half_rn(__half2float(a->sigma) / qmax);
  a->col_over = col_over;
  a->over_flag = over_flag;
  return 0;
}

int launch_rowquant(RowQuantArgs a, cudaStream_t st) {
  DeviceInfo di;
  if (int r = device_info(&di)) return r;
  constexpr int kThreads = 256, kWarps = kThreads / 32;
  constexpr long long kMaxSmem = 200 * 1024;
  // G warps per row so that a lane sees ~8 vectors; ngroups = 8 / G rows per CTA
  const int nvec = a.K / 8;
  int G = 1;
  while (G < kWarps && nvec > 256 * G) G *= 2;
  int ngroups = kWarps / G;
  while (ngroups > 1 && static_cast<long long>(ngroups) * a.K * 2 > kMaxSmem) ngroups /= 2;
  if (static_cast<long long>(ngroups) * a.K * 2 > kMaxSmem) return fail(MIXQ_EINVAL, "K too large for the row buffer");
  a.group_warps = kWarps / ngroups;
  a.ngroups = ngroups;
  if (a.K > kRowQuantMaxK) return fail(MIXQ_EINVAL, "activation prologue supports K <= 32768");
  const size_t smem = static_cast<size_t>(ngroups) * a.K * 2;
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(rowquant_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMaxSmem));
  });
  static thread_local int last_dev = -1;
  int dev = 0;
  MIXQ_CUDA(cudaGetDevice(&dev));
  if (dev != last_dev) {
    MIXQ_CUDA(cudaFuncSetAttribute(rowquant_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMaxSmem)));
    last_dev = dev;
  }
  long long grid = (a.M + ngroups - 1) / ngroups;
  const long long cap = static_cast<long long>(di.sms) * 16;
  if (grid > cap) grid = cap;
  MIXQ_CUDA(launch_small(rowquant_kernel, static_cast<int>(grid), kThreads, smem, st, a));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // namespace

extern "C" {

const char* mixq_last_error(void) { return g_err.c_str(); }
int mixq_version(void) { return 100; }
unsigned long long mixq_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int mixq_set_tile_n(int tile_n) {
  if (tile_n != 0 && (tile_n < 32 || tile_n > 512 || tile_n % 32 != 0))
    return fail(MIXQ_EINVAL, "tile_n must be 0 or a multiple of 32 up to 512 (the 1-CTA kernel honours 128 and 256 only)");
  g_tile_n.store(tile_n, std::memory_order_relaxed);
  return 0;
}

int mixq_plan_linear(int M, int N, int K, int bit, int n_ind, int swiglu_pair, int tile_n, int sms, mixq_linear_plan* out) {
  if (out == nullptr) return fail(MIXQ_EINVAL, "null plan");
  if (bit != 8 && bit != 4) return fail(MIXQ_EINVAL, "bit must be 8 or 4");
  if (K % 16 != 0 || N % 8 != 0) return fail(MIXQ_EINVAL, "K % 16 == 0 and N % 8 == 0 required");
  if (sms <= 0) {
    DeviceInfo di;
    if (int r = device_info(&di)) return r;
    sms = di.sms;
  }
  GemmPlan g{};
  if (int r = plan_gemm(M, N, K, bit, n_ind, swiglu_pair != 0, tile_n, sms, &g)) return r;
  out->two_cta = g.two_cta;
  out->tile_w = g.tile_w;
  out->k_atoms = g.k_atoms;
  out->stage_bytes = g.stage_bytes;
  out->nstages = g.nstages;
  out->tiles = g.tiles;
  out->units = g.units;
  out->tiles_per_unit = g.tiles_per_unit;
  out->acc_slots = g.tmem.slots;
  out->passes = g.tmem.passes;
  out->pass_cols = g.tmem.pass_cols;
  out->pass_buffers = g.tmem.buffers;
  out->tmem_cols = g.two_cta ? g.tmem.columns(g.tile_w, n_ind > 0) : g.tmem.slots * g.tile_w * (n_ind > 0 ? 2 : 1);
  return 0;
}

int mixq_set_grid_barrier_mode(int mode) {
  if (mode < 0 || mode > 2) return fail(MIXQ_EINVAL, "grid barrier mode: 0 = pdl, 1 = cooperative, 2 = split (two launches)");
  g_barrier_mode.store(mode, std::memory_order_relaxed);
  return 0;
}

int mixq_plan_split_k(int M, int N, int K, int bit, int n_ind, int sms, long long splitk_ws_bytes) {
  if ((bit != 8 && bit != 4) || K % 16 != 0 || N % 8 != 0) {
    fail(MIXQ_EINVAL, "bit must be 8 or 4, K % 16 == 0 and N % 8 == 0 required");
    return -1;
  }
  if (sms <= 0) {
    DeviceInfo di;
    if (device_info(&di)) return -1;
    sms = di.sms;
  }
  GemmPlan g{};
  if (plan_gemm(M, N, K, bit, n_ind, false, 0, sms, &g, splitk_ws_bytes)) return -1;
  return g.two_cta ? 1 : g.splits;
}

int mixq_set_pdl(int on) {
  g_pdl.store(on ? 1 : 0, std::memory_order_relaxed);
  return 0;
}

int mixq_set_peer_timeout_ms(long long ms) {
  if (ms <= 0) return fail(MIXQ_EINVAL, "peer timeout must be positive");
  g_peer_timeout_ms.store(static_cast<unsigned long long>(ms), std::memory_order_relaxed);
  return 0;
}

int mixq_set_trace_buffer(void* buf) {
  g_trace.store(static_cast<unsigned long long*>(buf), std::memory_order_relaxed);
  return 0;
}

int mixq_find_row_scale(const void* x, void* x_scale, void* q_x, int M, int K, int bit, void* stream) {
  if (x == nullptr || x_scale == nullptr || q_x == nullptr) return fail(MIXQ_EINVAL, "null pointer");
  RowQuantArgs a{};
  if (int r = fill_rowquant(&a, const_cast<void*>(x), nullptr, nullptr, 0.f, nullptr, 0, nullptr, 0, q_x, x_scale,
                            M, K, bit, 0.f, nullptr, nullptr))
    return r;
  return launch_rowquant(a, static_cast<cudaStream_t>(stream));
}

int mixq_find_row_scale_scan(const void* x, void* x_scale, void* q_x, int M, int K, int bit, float sigma,
                             uint8_t* col_over, uint32_t* over_flag, void* stream) {
  if (x == nullptr || x_scale == nullptr || q_x == nullptr) return fail(MIXQ_EINVAL, "null pointer");
  RowQuantArgs a{};
  if (int r = fill_rowquant(&a, const_cast<void*>(x), nullptr, nullptr, 0.f, nullptr, 0, nullptr, 0, q_x, x_scale,
                            M, K, bit, sigma, col_over, over_flag))
    return r;
  return launch_rowquant(a, static_cast<cudaStream_t>(stream));
}

int mixq_extract_outliers_and_set_to_zeros(const int32_t* ind, int n_ind, void* x, void* out, int ld_out, int M,
                                           int K, void* stream) {
  if (n_ind == 0) return 0;
  if (ind == nullptr || x == nullptr || out == nullptr || n_ind < 0 || ld_out < n_ind || M < 1)
    return fail(MIXQ_EINVAL, "bad extract arguments");
  DeviceInfo di;
  if (int r = device_info(&di)) return r;
  const int grid = grid_for(static_cast<long long>(M) * n_ind, 256, di.sms);
  MIXQ_CUDA(launch_small(extract_outliers_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), ind, n_ind,
                         static_cast<__half*>(x), static_cast<__half*>(out), ld_out, M, K));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mixq_int8_fused_dequantize(const void* q_x, const void* q_w, const void* x_scale, const void* scale_col,
                               const void* outl, int ld_outl, void* y, int M, int N, int K, int act,
                               void* stream) {
  if (!q_x || !q_w || !x_scale || !scale_col || !y) return fail(MIXQ_EINVAL, "null pointer");
  GemmCall c{};
  c.q_x = q_x; c.q_w = q_w; c.bit = 8; c.x_scale = x_scale; c.scale_col = scale_col;
  c.outl = outl; c.ld_outl = ld_outl; c.y = y; c.M = M; c.N = N; c.K = K; c.act = act;
  c.epilogue = EPI_DEQUANT_F16;
  return run_gemm(c, static_cast<cudaStream_t>(stream));
}

int mixq_int4_fused_dequantize(const void* q_x, const void* q_w_packed, const void* x_scale,
                               const void* scale_col, const void* outl, int ld_outl, void* y, int M, int N,
                               int K, int act, void* stream) {
  if (!q_x || !q_w_packed || !x_scale || !scale_col || !y) return fail(MIXQ_EINVAL, "null pointer");
  GemmCall c{};
  c.q_x = q_x; c.q_w = q_w_packed; c.bit = 4; c.x_scale = x_scale; c.scale_col = scale_col;
  c.outl = outl; c.ld_outl = ld_outl; c.y = y; c.M = M; c.N = N; c.K = K; c.act = act;
  c.epilogue = EPI_DEQUANT_F16;
  return run_gemm(c, static_cast<cudaStream_t>(stream));
}

int mixq_gemm_i8(const void* q_x, const void* q_w, int32_t* y, int M, int N, int K, void* stream) {
  if (!q_x || !q_w || !y) return fail(MIXQ_EINVAL, "null pointer");
  GemmCall c{};
  c.q_x = q_x; c.q_w = q_w; c.bit = 8; c.y_i32 = y; c.M = M; c.N = N; c.K = K;
  c.epilogue = EPI_RAW_I32;
  return run_gemm(c, static_cast<cudaStream_t>(stream));
}

int mixq_dequantize_int8(const int32_t* acc, const void* x_scale, const void* scale_col, const void* outl,
                         int ld_outl, void* y, int M, int N, int act, void* stream) {
  if (!acc || !x_scale || !scale_col || !y || M < 1 || N % 8 != 0) return fail(MIXQ_EINVAL, "bad dequantize arguments");
  DeviceInfo di;
  if (int r = device_info(&di)) return r;
  const int grid = grid_for(static_cast<long long>(M) * (N / 8), 256, di.sms);
  MIXQ_CUDA(launch_small(dequant_i32_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), acc,
                         static_cast<const __half*>(x_scale), static_cast<const __half*>(scale_col),
                         static_cast<const __half*>(outl), ld_outl, static_cast<__half*>(y), M, N, act));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mixq_unpack_int4_to_fp16(const void* q_w_packed, const int32_t* ind, int n_ind, void* out, int ld_out, int N,
                             int K, void* stream) {
  if (n_ind == 0) return 0;
  if (!q_w_packed || !ind || !out || ld_out < n_ind) return fail(MIXQ_EINVAL, "bad unpack arguments");
  DeviceInfo di;
  if (int r = device_info(&di)) return r;
  const int grid = grid_for(static_cast<long long>(N) * n_ind, 256, di.sms);
  unpack_int4_cols_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(q_w_packed), ind, n_ind, static_cast<__half*>(out), ld_out, N, K);
  MIXQ_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mixq_rmsnorm(const void* x, const void* w, void* out, float eps, int M, int K, void* stream) {
  if (!x || !w || !out) return fail(MIXQ_EINVAL, "null pointer");
  RowQuantArgs a{};
  if (int r = fill_rowquant(&a, const_cast<void*>(x), w, out, eps, nullptr, 0, nullptr, 0, nullptr, nullptr, M, K,
                            8, 0.f, nullptr, nullptr))
    return r;
  return launch_rowquant(a, static_cast<cudaStream_t>(stream));
}

int mixq_rmsnorm_extract_outliers(const void* x, const void* w, void* out, float eps, const int32_t* ind,
                                  int n_ind, void* x_scale, void* act_out, int ld_ao, void* q_x, int M, int K,
                                  int bit, void* stream) {
  if (!x || !w || !out || !x_scale || !q_x) return fail(MIXQ_EINVAL, "null pointer");
  RowQuantArgs a{};
  if (int r = fill_rowquant(&a, const_cast<void*>(x), w, out, eps, ind, n_ind, act_out, ld_ao, q_x, x_scale, M, K,
                            bit, 0.f, nullptr, nullptr))
    return r;
  return launch_rowquant(a, static_cast<cudaStream_t>(stream));
}

int mixq_gather_weight_columns(const void* q_w, const void* scale_col, const int32_t* ind, int n_ind, void* wc,
                               int ld_wc, int col0, int N, int K, int bit, void* stream) {
  if (n_ind == 0) return 0;
  if (!q_w || !scale_col || !ind || !wc || col0 < 0 || ld_wc < col0 + n_ind || (bit != 8 && bit != 4))
    return fail(MIXQ_EINVAL, "bad gather arguments");
  DeviceInfo di;
  if (int r = device_info(&di)) return r;
  const int grid = grid_for(static_cast<long long>(N) * n_ind, 256, di.sms);
  gather_weight_cols_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      q_w, static_cast<const __half*>(scale_col), ind, n_ind, static_cast<__half*>(wc), ld_wc, col0, N, K, bit);
  MIXQ_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mixq_compact_outlier_columns(uint8_t* col_over, int K, int32_t* ind_out, int max_new, int32_t* n_new,
                                 void* stream) {
  if (!col_over || !ind_out || !n_new || K < 1 || max_new < 0) return fail(MIXQ_EINVAL, "bad compact arguments");
  compact_cols_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(col_over, K, ind_out, max_new, n_new);
  MIXQ_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mixq_linear_fused(const mixq_linear_args* a, void* stream) {
  if (a == nullptr) return fail(MIXQ_EINVAL, "null args");
  if (!a->q_weight || !a->scale_col || !a->q_x || !a->x_scale || (!a->y && a->peer_cols <= 0 && a->peer_bcast <= 0))
    return fail(MIXQ_EINVAL, "null pointer");
  if (a->n_ind > 0 && (!a->ind || !a->weight_cache || !a->act_outliers))
    return fail(MIXQ_EINVAL, "n_ind > 0 needs ind, weight_cache and act_outliers");
  RowQuantArgs rq{};
  if (!a->skip_prologue) {
    if (!a->x || !a->grid_sync) return fail(MIXQ_EINVAL, "prologue needs x and grid_sync");
    if (int r = fill_rowquant(&rq, a->x, a->norm_weight, a->norm_out, a->eps, a->ind, a->n_ind, a->act_outliers,
                              a->ld_ao, a->q_x, a->x_scale, a->M, a->K, a->bit, a->sigma, a->col_over,
                              a->over_flag))
      return r;
  }
  GemmCall c{};
  c.q_x = a->q_x; c.q_w = a->q_weight; c.bit = a->bit; c.x_scale = a->x_scale; c.scale_col = a->scale_col;
  c.bias = a->bias; c.act_outliers = a->act_outliers; c.ld_ao = a->ld_ao; c.weight_cache = a->weight_cache;
  c.ld_wc = a->ld_wc; c.n_out = a->n_ind; c.y = a->y; c.M = a->M; c.N = a->N; c.K = a->K; c.act = a->act;
  c.epilogue = EPI_DEQUANT_F16; c.tile_n = a->tile_n;
  c.residual = a->residual; c.ld_res = a->ld_res;
  c.q_w_up = a->q_weight_up; c.scale_col_up = a->scale_col_up; c.weight_cache_up = a->weight_cache_up;
  if (a->residual && a->ld_res < a->N) return fail(MIXQ_EINVAL, "ld_res < N");
  c.rq = a->skip_prologue ? nullptr : &rq;
  if (c.rq != nullptr && barrier_mode() == 2) {
    // "split": the activation prologue as its own launch, then the GEMM without a grid barrier (bit-identical)
    if (int r = launch_rowquant(rq, static_cast<cudaStream_t>(stream))) return r;
    c.rq = nullptr;
  }
  c.grid_sync = a->grid_sync;
  c.y_peer = a->y_peer;
  c.peer_cols = a->peer_cols;
  c.peer_bcast = a->peer_bcast;
  c.splitk_ws = a->splitk_ws;
  c.splitk_ws_bytes = a->splitk_ws_bytes;
  return run_gemm(c, static_cast<cudaStream_t>(stream));
}

int mixq_rope_attention_decode(const void* qkv, void* k_cache, void* v_cache, int cache_cap, int past_len, void* out,
                               int M, int H, int Hkv, int D, float theta, void* stream) {
  if (!qkv || !out || M < 1 || H < 1 || Hkv < 1 || H % Hkv != 0 || (D != 64 && D != 128) || past_len < 0)
    return fail(MIXQ_EINVAL, "bad attention arguments (head_dim must be 64 or 128)");
  if (past_len > 0 && (!k_cache || !v_cache || cache_cap <= past_len))
    return fail(MIXQ_EINVAL, "past_len > 0 needs k/v caches with capacity > past_len");
  MIXQ_CUDA(launch_rope_attn_decode(static_cast<const __half*>(qkv), static_cast<__half*>(k_cache),
                                    static_cast<__half*>(v_cache), cache_cap, past_len, static_cast<__half*>(out), M, H,
                                    Hkv, D, theta, pdl_on(), static_cast<cudaStream_t>(stream)));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mixq_rope_attention_decode_quant(const void* qkv, void* k_cache, void* v_cache, int cache_cap, int past_len, void* out,
                                     int M, int H, int Hkv, int D, float theta, const int32_t* ind, int n_ind,
                                     void* act_outliers, int ld_ao, void* q_x, void* x_scale, int bit, void* stream) {
  if (!qkv || !q_x || !x_scale || M < 1 || H < 1 || Hkv < 1 || H % Hkv != 0 || (D != 64 && D != 128) || past_len < 0)
    return fail(MIXQ_EINVAL, "bad attention arguments (head_dim must be 64 or 128)");
  if (past_len > 0 && (!k_cache || !v_cache || cache_cap <= past_len))
    return fail(MIXQ_EINVAL, "past_len > 0 needs k/v caches with capacity > past_len");
  if (static_cast<long long>(2 * H + 2 * Hkv) * D * 2 > 96 * 1024)
    return fail(MIXQ_EINVAL, "attention row does not fit the shared-memory row buffer");
  RowQuantArgs rq{};
  if (int r = fill_rowquant(&rq, out, nullptr, nullptr, 0.f, ind, n_ind, act_outliers, ld_ao, q_x, x_scale, M, H * D, bit,
                            0.f, nullptr, nullptr))
    return r;
  rq.group_warps = 4;   // the whole CTA (kAttnQuantWarps warps) owns the row
  rq.ngroups = 1;
  static thread_local int last_dev = -1;
  int dev = 0;
  MIXQ_CUDA(cudaGetDevice(&dev));
  if (dev != last_dev) {
    MIXQ_CUDA(cudaFuncSetAttribute(rope_attn_decode_quant_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    MIXQ_CUDA(cudaFuncSetAttribute(rope_attn_decode_quant_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    last_dev = dev;
  }
  MIXQ_CUDA(launch_rope_attn_decode(static_cast<const __half*>(qkv), static_cast<__half*>(k_cache), static_cast<__half*>(v_cache),
                                    cache_cap, past_len, static_cast<__half*>(out), M, H, Hkv, D, theta, pdl_on(),
                                    static_cast<cudaStream_t>(stream), &rq));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mixq_peer_alloc(unsigned long long bytes, void** ptr) {
  if (!ptr || bytes == 0) return fail(MIXQ_EINVAL, "bad peer_alloc arguments");
  MIXQ_CUDA(cudaMalloc(ptr, bytes));            // its own allocation: the IPC handle maps exactly this buffer
  MIXQ_CUDA(cudaMemset(*ptr, 0, bytes));
  MIXQ_CUDA(cudaDeviceSynchronize());
  return 0;
}
int mixq_peer_free(void* ptr) {
  MIXQ_CUDA(cudaFree(ptr));
  return 0;
}
int mixq_ipc_get_handle(const void* ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!ptr || !handle64) return fail(MIXQ_EINVAL, "null pointer");
  MIXQ_CUDA(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), const_cast<void*>(ptr)));
  return 0;
}
int mixq_ipc_open_handle(const void* handle64, void** ptr) {
  if (!ptr || !handle64) return fail(MIXQ_EINVAL, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  MIXQ_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int mixq_ipc_close_handle(void* ptr) {
  MIXQ_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

int mixq_allreduce_residual(const mixq_allreduce_args* a, void* stream) {
  if (!a || a->world < 2 || a->world > kMaxPeers || a->rank < 0 || a->rank >= a->world || !a->out || !a->epoch || !a->done ||
      a->n < 8 || (a->n & 7))
    return fail(MIXQ_EINVAL, "bad all-reduce arguments (2 <= world <= 8, n % 8 == 0)");
  AllReduceArgs k{};
  for (int p = 0; p < a->world; ++p) {
    if (!a->partial0[p] || !a->partial1[p] || !a->flags[p]) return fail(MIXQ_EINVAL, "missing peer pointer");
    k.partial[p][0] = static_cast<const __half*>(a->partial0[p]);
    k.partial[p][1] = static_cast<const __half*>(a->partial1[p]);
    k.flags[p] = static_cast<uint32_t*>(a->flags[p]);
    k.result[p][0] = static_cast<__half*>(a->result0[p]);
    k.result[p][1] = static_cast<__half*>(a->result1[p]);
    if ((a->result0[p] == nullptr) != (a->result0[0] == nullptr) || (a->result1[p] == nullptr) != (a->result0[0] == nullptr))
      return fail(MIXQ_EINVAL, "two-shot needs the result buffers of every rank");
  }
  if (a->result0[0] != nullptr && a->out != a->result0[a->rank] && a->out != a->result1[a->rank])
    return fail(MIXQ_EINVAL, "two-shot: out must be this rank's result buffer of the exchange");
  k.epoch = static_cast<uint32_t*>(a->epoch);
  k.done = static_cast<uint32_t*>(a->done);
  k.residual = static_cast<const __half*>(a->residual);
  k.out = static_cast<__half*>(a->out);
  k.n = a->n;
  k.world = a->world;
  k.rank = a->rank;
  k.buf = a->buf & 1;
  k.timeout_ns = g_peer_timeout_ms.load(std::memory_order_relaxed) * 1000000ull;
  DeviceInfo di;
  if (int r = device_info(&di)) return r;
  const int grid = grid_for(a->n / 8, 256, di.sms, 4);
  MIXQ_CUDA(launch_small(allreduce_residual_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), k));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mixq_allreduce_multicast(const mixq_mc_allreduce_args* a, void* stream) {
  if (!a || a->world < 2 || a->world > kMaxPeers || a->rank < 0 || a->rank >= a->world || !a->mc || !a->local || !a->epoch ||
      !a->done || a->n < 8 || (a->n % (8ll * a->world)) != 0)
    return fail(MIXQ_EINVAL, "bad multicast all-reduce arguments (2 <= world <= 8, n % (8 * world) == 0)");
  for (int b = 0; b < 2; ++b)
    if ((a->partial_off[b] & 15) || (a->result_off[b] & 15)) return fail(MIXQ_EINVAL, "buffer offsets must be 16-byte aligned");
  if (a->flags_off & 15) return fail(MIXQ_EINVAL, "flags offset must be 16-byte aligned");
  McAllReduceArgs k{};
  k.mc = static_cast<uint8_t*>(a->mc);
  k.local = static_cast<uint8_t*>(a->local);
  for (int b = 0; b < 2; ++b) {
    k.partial_off[b] = a->partial_off[b];
    k.result_off[b] = a->result_off[b];
  }
  k.flags_off = a->flags_off;
  k.epoch = static_cast<uint32_t*>(a->epoch);
  k.done = static_cast<uint32_t*>(a->done);
  k.residual = static_cast<const __half*>(a->residual);
  k.n = a->n;
  k.world = a->world;
  k.rank = a->rank;
  k.buf = a->buf & 1;
  k.timeout_ns = g_peer_timeout_ms.load(std::memory_order_relaxed) * 1000000ull;
  DeviceInfo di;
  if (int r = device_info(&di)) return r;
  const int grid = grid_for(a->n / 8 / a->world, 256, di.sms, 2);
  MIXQ_CUDA(launch_small(allreduce_multicast_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), k));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mixq_exchange_finish(const mixq_exchange_finish_args* a, void* stream) {
  if (!a || a->world < 2 || a->world > kMaxPeers || a->rank < 0 || a->rank >= a->world || !a->recv || !a->epoch || !a->done ||
      a->M < 1 || a->N < 8 || a->N % (8 * a->world) != 0)
    return fail(MIXQ_EINVAL, "bad exchange_finish arguments (2 <= world <= 8, N % (8 * world) == 0)");
  XchgFinishArgs k{};
  k.recv = static_cast<const __half*>(a->recv);
  for (int p = 0; p < a->world; ++p) {
    if (!a->flags[p] || (!a->mc_result && !a->result[p] && !(a->one_shot && p != a->rank))) return fail(MIXQ_EINVAL, "missing peer pointer");
    k.result[p] = static_cast<__half*>(a->result[p]);
    k.flags[p] = static_cast<uint32_t*>(a->flags[p]);
  }
  k.mc_result = static_cast<__half*>(a->mc_result);
  k.mc_flags = static_cast<uint32_t*>(a->mc_flags);
  k.epoch = static_cast<uint32_t*>(a->epoch);
  k.done = static_cast<uint32_t*>(a->done);
  k.residual = static_cast<const __half*>(a->residual);
  k.M = a->M;
  k.N = a->N;
  k.world = a->world;
  k.rank = a->rank;
  k.one_shot = a->one_shot ? 1 : 0;
  k.timeout_ns = g_peer_timeout_ms.load(std::memory_order_relaxed) * 1000000ull;
  static const int xf_ablate = [] { const char* e = getenv("MIXQ_DEBUG_XF"); return e ? atoi(e) : 0; }();
  k.ablate = xf_ablate;
  k.trace = g_trace.load(std::memory_order_relaxed);
  DeviceInfo di;
  if (int r = device_info(&di)) return r;
  const int grid = grid_for(static_cast<long long>(a->M) * (a->N / (a->one_shot ? 1 : a->world) / 8), 256, di.sms, 2);
  MIXQ_CUDA(launch_small(exchange_finish_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), k));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mixq_quik_quantize(const void* x, const int64_t* int_indices, int n_int, const int64_t* fp_indices, int n_fp, int bits,
                       void* q, void* meta, void* fp_x, int M, int K, void* stream) {
  if (!x || !int_indices || !q || !meta || M < 1 || n_int < 1 || n_int + n_fp > K || (bits != 4 && bits != 8) ||
      (n_fp > 0 && (!fp_indices || !fp_x)))
    return fail(MIXQ_EINVAL, "bad quik_quantize arguments");
  DeviceInfo di;
  if (int r = device_info(&di)) return r;
  const int grid = M < di.sms * 8 ? M : di.sms * 8;
  MIXQ_CUDA(launch_small(quik_quantize_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), static_cast<const __half*>(x),
                         int_indices, n_int, fp_indices, n_fp, bits, static_cast<int8_t*>(q), static_cast<__half*>(meta),
                         static_cast<__half*>(fp_x), M, K));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mixq_quik_addend(const void* meta, const void* reduced_w, const void* fp_result, int ld_fp, void* out, int M, int N,
                     int bits, void* stream) {
  if (!meta || !reduced_w || !out || M < 1 || N < 8 || N % 8 != 0 || (bits != 4 && bits != 8) || (fp_result && ld_fp < N))
    return fail(MIXQ_EINVAL, "bad quik_addend arguments");
  DeviceInfo di;
  if (int r = device_info(&di)) return r;
  const int grid = grid_for(static_cast<long long>(M) * (N / 8), 256, di.sms);
  MIXQ_CUDA(launch_small(quik_addend_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), static_cast<const __half*>(meta),
                         static_cast<const __half*>(reduced_w), static_cast<const __half*>(fp_result), ld_fp,
                         static_cast<__half*>(out), M, N, bits));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // extern "C"
namespace {
int fill_xchg_poll(XchgPollArgs* kp, const mixq_exchange_poll_args* a) {
  XchgPollArgs& k = *kp;
  if (!a || a->world < 2 || a->world > kMaxPeers || a->rank < 0 || a->rank >= a->world || !a->recv || a->M < 1 || a->N < 8 ||
      a->N % (8 * a->world) != 0)
    return fail(MIXQ_EINVAL, "bad exchange_finish_poll arguments (2 <= world <= 8, N % (8 * world) == 0)");
  k.recv = static_cast<__half*>(a->recv);
  for (int p = 0; p < a->world; ++p) {
    if (!a->mc_result && !a->result[p] && !(a->one_shot && p != a->rank)) return fail(MIXQ_EINVAL, "missing peer pointer");
    k.result[p] = static_cast<__half*>(a->result[p]);
  }
  if (!a->result[a->rank]) return fail(MIXQ_EINVAL, "missing local result buffer");
  k.mc_result = static_cast<__half*>(a->mc_result);
  k.reset = static_cast<__half*>(a->reset);
  k.residual = static_cast<const __half*>(a->residual);
  k.M = a->M;
  k.N = a->N;
  k.world = a->world;
  k.rank = a->rank;
  k.one_shot = a->one_shot ? 1 : 0;
  k.timeout_ns = g_peer_timeout_ms.load(std::memory_order_relaxed) * 1000000ull;
  k.trace = g_trace.load(std::memory_order_relaxed);
  return 0;
}
}  // namespace
extern "C" {

int mixq_exchange_finish_poll(const mixq_exchange_poll_args* a, void* stream) {
  XchgPollArgs k{};
  if (int r = fill_xchg_poll(&k, a)) return r;
  DeviceInfo di;
  if (int r = device_info(&di)) return r;
  const int grid = grid_for(static_cast<long long>(a->M) * (a->N / (a->one_shot ? 1 : a->world) / 8), 256, di.sms, 2);
  MIXQ_CUDA(launch_small(exchange_finish_poll_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), k));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mixq_exchange_finish_poll_quant(const mixq_exchange_poll_args* a, const void* norm_weight, float eps, const int32_t* ind,
                                    int n_ind, void* act_outliers, int ld_ao, void* q_x, void* x_scale, int bit, void* stream) {
  XchgPollArgs k{};
  if (int r = fill_xchg_poll(&k, a)) return r;
  if (!norm_weight || !q_x || !x_scale) return fail(MIXQ_EINVAL, "exchange_finish_poll_quant needs norm_weight, q_x and x_scale");
  if (static_cast<long long>(a->N) * 2 > 48 * 1024) return fail(MIXQ_EINVAL, "row does not fit the shared-memory row buffer");
  RowQuantArgs rq{};
  if (int r = fill_rowquant(&rq, nullptr, norm_weight, nullptr, eps, ind, n_ind, act_outliers, ld_ao, q_x, x_scale, a->M, a->N, bit,
                            0.f, nullptr, nullptr))
    return r;
  rq.group_warps = 4;   // the whole CTA (kXchgQuantThreads) owns the row
  rq.ngroups = 1;
  MIXQ_CUDA(launch_exchange_finish_rowquant(k, rq, a->M, pdl_on(), static_cast<cudaStream_t>(stream)));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

void mixq_reload_debug_env(void) { g_dbg.load(); }

int mixq_debug_pingpong(void* mine, void* peer, void* mc, int iters, int rank, void* out_ns, void* stream) {
  pingpong_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<uint32_t*>(mine), static_cast<uint32_t*>(peer),
                                                                    static_cast<uint32_t*>(mc), iters, rank,
                                                                    static_cast<unsigned long long*>(out_ns));
  MIXQ_CUDA(cudaGetLastError());
  return 0;
}

int mixq_mul_inplace(void* a, const void* b, long long n, void* stream) {
  if (!a || !b || n < 0 || (n & 1)) return fail(MIXQ_EINVAL, "bad mul arguments");
  DeviceInfo di;
  if (int r = device_info(&di)) return r;
  const int grid = grid_for(n / 2, 256, di.sms);
  MIXQ_CUDA(launch_small(mul_inplace_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), static_cast<__half2*>(a),
                         static_cast<const __half2*>(b), n / 2));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // extern "C"
