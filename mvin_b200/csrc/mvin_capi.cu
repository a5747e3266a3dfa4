// mvin_capi.cu -- libmvin_b200.so: C ABI (include/mvin_b200.h) and the host-side orchestration of the kernels.
//
// Forward  = model.py:125-159 of the reference (src/model/MVIN/), backward = TF autodiff of model.py:378-412,
// Adam = model.py:414.  The factorisation the kernels implement is spelled out in DESIGN.md section 3 and has a
// CPU twin in tests/fused_model.py.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include "../../include/mvin_b200.h"
#include "gemm.cuh"
#include "level.cuh"
#include "level_tc.cuh"
#include "misc.cuh"
#include "umma.cuh"
#include "user.cuh"

using namespace mvin;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void prof_mark(mvin_handle_t h, cudaStream_t st, const char* name);

// Every kernel goes through here.  With programmatic stream serialisation the kernel may start while its predecessor
// in the stream is still running; each kernel's first statement is pdl_enter() (common.cuh), which restores the
// dependency on the device.  MVIN_B200_PDL=0 launches normally.
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("MVIN_B200_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <typename... KArgs, typename... Args>
void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);   // error picked up by LAUNCH_CHECK
}
#define MVIN_LAUNCH(kernel, grid, block, smem, st, ...) launch_kernel(kernel, grid, block, smem, st, ##__VA_ARGS__)

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) return fail(MVIN_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                       __FILE__, __LINE__);                                         \
  } while (0)

#define LAUNCH_CHECK(h, name)                                                                       \
  do {                                                                                              \
    (h)->launches++;                                                                                \
    prof_mark((h), st, name);                                                                       \
    cudaError_t _e = cudaGetLastError();                                                            \
    if (_e != cudaSuccess) return fail(MVIN_ERR_CUDA, "launch %s: %s", name, cudaGetErrorString(_e)); \
  } while (0)

constexpr int MAX_L = 3;
constexpr int MAX_SHARDS = 16;

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Workspace layout for batch size B (all offsets in bytes from the workspace base).
//   V[j][h]   output of aggregator iteration j-1 at level h (V[0][h] = T[h], the user-oriented transform); the
//             level-0 slices V[0..H][0] are contiguous (Vtop) so the mix layer is one batch-reduce GEMM
//   Y[i][h]   GEMM input (self + agg) of aggregator iteration i at level h
//   DC[j][h]  gradient of V[j][h] arriving from its parent's dchild (h >= 1) or from the mix layer (h = 0)
//   DS[j][h]  gradient of V[j][h] arriving from its own aggregator step (iteration j, level h)
struct Layout {
  size_t ent[MAX_L];                      // int32 [B K^h], h < L
  size_t Vbuf, Q, probs, O, u, s;         // user side + relation scores
  size_t SU;                              // leaf: S + u
  size_t Y[MAX_L][MAX_L];                 // Y[i][h], i < H, h < H - i
  size_t V[MAX_L + 1][MAX_L];             // V[j][h]
  size_t item, scores;
  // backward
  size_t DC[MAX_L + 1][MAX_L], DS[MAX_L][MAX_L];
  size_t du, ditem, dO, wT;               // wT: [H + H + 1][D][D] transposed weights
  size_t zero_begin, ds, cnt, acc, zero_mid, dQ, dv, GSe, zero_end;   // cleared at the start of every backward:
                                          // [begin, mid) on the launch stream, [mid, end) on a side stream
  size_t stamp, Se;                       // entity mode of the leaf level (stamp is cleared by every forward)
  bool entity_leaf;
  size_t total;
  long rows[MAX_L + 1];
};

}  // namespace

struct mvin_handle_s {
  mvin_config_t cfg;
  mvin_params_t P, G;
  bool has_params = false, has_grads = false;
  const int32_t* adj = nullptr;
  const int32_t* uts = nullptr;    // device-resident ripple sets [n_user, max(1,p), 3, m] (mvin_bind_user_triplets)
  int device = 0, sm_count = 148;
  int64_t launches = 0;
  // batch of the last forward (pointers owned by the caller, must stay valid until backward)
  const int64_t* user = nullptr;
  const int64_t* item = nullptr;
  const int32_t *mem_h = nullptr, *mem_r = nullptr, *mem_t = nullptr;
  int B = 0;
  void* fwd_workspace = nullptr;
  // optional per-category kernel timing (mvin_profile_enable / mvin_profile_read)
  // entity table / gradient accessors (single table, or row-sharded over n_shards peers)
  ETab etab{};
  GTab gtab{};
  int n_shards = 1;
  long n_local_rows = 0;           // rows of the local entity shard
  void** d_shard_tab = nullptr;    // device array [2][MAX_SHARDS] of shard base pointers (allocated in mvin_create)
  int* d_sched = nullptr;          // tile-scheduler counter pairs of the aggregator kernels (level.cuh), zero between launches
  int global_batch = 0;            // 0: the batch of the call
  float dense_l2_scale = 1.f;
  // fork/join helpers: independent kernels of a step run on two internal side streams (disabled while profiling)
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr}, ev_mid = nullptr;
  bool use_streams = true;
  bool in_host_step = false;       // set for the duration of a host-step entry point (guards the two flags below)
  bool early_init = false;         // backward part 0 of this step was already enqueued (host-step entry points)
  cudaEvent_t ev_early = nullptr, ev_item = nullptr;   // its completion; 'item ids are on the device'
  cudaStream_t copy_stream = nullptr;   // mvin_feed_prefetch: H2D copies of the NEXT batch while this one computes
  cudaEvent_t ev_feed[2] = {nullptr, nullptr}, ev_free = nullptr;   // one 'feed landed' event per staging slot
  const void* prefetched_staging[2] = {nullptr, nullptr};
  int prefetched_B[2] = {0, 0};
  bool pre_fork = false;           // forward: the side stream starts from ev_item instead of the launch stream's tail
  int entity_leaf_mode = -1;       // -1 auto, 0 off, 1 on (env MVIN_B200_ENTITY_LEAF, read in mvin_create)
  int user_pb_fwd = 4;             // max pairs per CTA of the user-side forward kernel (env MVIN_B200_USER_PB_FWD)
  int stream_mode = -1;            // -1 auto, 0 never, 1 always (env MVIN_B200_STREAM)
  int tc_mode = 1;                 // tcgen05 forward row kernels for d in {32, 64}: 0 never, 1 auto, 2 always (env MVIN_B200_TC)
  int max_ctas_per_sm = 4;         // cap on resident CTAs per SM of the persistent row kernels (env MVIN_B200_CTAS_PER_SM)
  bool prof_on = false;
  struct ProfRec { const char* name; cudaEvent_t ev; };
  std::vector<ProfRec> prof;
};

namespace {

// Kernel timing: when enabled, one CUDA event is recorded on the launch stream after every kernel launch (and one
// marker at each API entry); the stream is in-order, so consecutive events bracket one kernel.
void prof_mark(mvin_handle_t h, cudaStream_t st, const char* name) {
  if (!h->prof_on) return;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, st);
  h->prof.push_back({name, ev});
}

// M of fastdiv (level.cuh): floor(2^64 / d) + 1, 0 for d = 1
inline unsigned long long div_magic(long d) { return d <= 1 ? 0ull : ~0ull / (unsigned long long)d + 1ull; }

// Fork/join of the launch stream onto the handle's side streams.  Plain stream/event calls, so a step can also be
// stream-captured into a CUDA graph with the side work as parallel branches.
struct Par {
  mvin_handle_t h;
  cudaStream_t main;
  bool on;
  cudaStream_t s(int i) const { return on ? h->side[i] : main; }
  void fork(int i) const {
    if (!on) return;
    cudaEventRecord(h->ev_fork[i], main);
    cudaStreamWaitEvent(h->side[i], h->ev_fork[i], 0);
  }
  void join(int i) const {
    if (!on) return;
    cudaEventRecord(h->ev_join[i], h->side[i]);
    cudaStreamWaitEvent(main, h->ev_join[i], 0);
  }
  // partial join: the launch stream waits for what side stream 0 has been given so far, the side stream carries on
  void mark_mid() const { if (on) cudaEventRecord(h->ev_mid, h->side[0]); }
  void wait_mid() const { if (on) cudaStreamWaitEvent(main, h->ev_mid, 0); }
};

inline bool has_agg(int H, int i, int h) { return i < H && h < H - i; }          // aggregator step (i, h) exists
inline bool has_V(int H, int j, int h) { return j == 0 ? h < H : h <= H - j; }   // buffer V[j][h] exists

// Entity mode of the leaf level (level.cuh, leaf_entity_kernel) pays off when the depth-(L-1) nodes of a batch
// re-use entities: enabled when there are at least n_entity / 4 of them and the two per-entity buffers are small.
bool use_entity_leaf(const mvin_config_t& c, long B, int n_shards, int mode) {
  if (n_shards != 1 || mode == 0) return false;
  if (mode == 1) return true;
  long rows = B;
  for (int h = 1; h < c.h_hop; ++h) rows *= c.neighbor_sample_size;
  return n_shards == 1 && rows * 4 >= (long)c.n_entity && (long)c.n_entity * c.dim * 8 <= (2L << 30);
}

Layout make_layout(const mvin_config_t& c, long B, bool entity_leaf) {
  Layout L;
  memset(&L, 0, sizeof(L));
  L.entity_leaf = entity_leaf;
  const long D = c.dim, K = c.neighbor_sample_size, H = c.h_hop, p = c.p_hop, m = c.n_memory, nr = c.n_relation;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes);
    return o;
  };
  long r = B;
  for (int h = 0; h <= H; ++h) { L.rows[h] = r; r *= K; }
  for (int h = 0; h < H; ++h) L.ent[h] = take(sizeof(int32_t) * L.rows[h]);
  const size_t f = sizeof(float);
  L.Vbuf = take(f * B * D);
  L.Q = take(f * B * nr * D);
  L.probs = take(f * (p + 1) * B * m);
  L.O = take(f * B * (p + 1) * D);
  L.u = take(f * B * D);
  L.s = take(f * H * nr);
  L.SU = take(f * L.rows[H - 1] * D);
  const size_t vtop = take(f * (H + 1) * B * D);
  for (int j = 0; j <= H; ++j)
    for (int h = 0; h < MAX_L; ++h)
      if (has_V(H, j, h)) L.V[j][h] = h == 0 ? vtop + f * j * B * D : take(f * L.rows[h] * D);
  for (int i = 0; i < H; ++i)
    for (int h = 0; h < H - i; ++h) L.Y[i][h] = take(f * L.rows[h] * D);
  L.item = take(f * B * D);
  L.scores = take(f * B);
  const size_t dtop = take(f * (H + 1) * B * D);
  for (int j = 0; j <= H; ++j)
    for (int h = 0; h < MAX_L; ++h)
      if (has_V(H, j, h)) L.DC[j][h] = h == 0 ? dtop + f * j * B * D : take(f * L.rows[h] * D);
  for (int i = 0; i < H; ++i)
    for (int h = 0; h < H - i; ++h) L.DS[i][h] = take(f * L.rows[h] * D);
  L.du = take(f * B * D);
  L.ditem = take(f * B * D);
  L.dO = take(f * B * (p + 1) * D);
  L.wT = take(f * (2 * H + 1) * D * D);
  L.zero_begin = off;
  L.ds = take(f * H * nr);
  L.cnt = take(f * nr);
  L.acc = take(f * 8);
  L.zero_mid = off;
  L.dQ = take(f * B * nr * D);
  L.dv = take(f * B * D);
  if (entity_leaf) L.GSe = take(f * (size_t)c.n_entity * D);
  L.zero_end = off;
  if (entity_leaf) {
    L.stamp = take(sizeof(int32_t) * (size_t)c.n_entity);
    L.Se = take(f * (size_t)c.n_entity * D);
  }
  L.total = off;
  return L;
}

template <typename T>
T* at(void* ws, size_t off) { return reinterpret_cast<T*>(static_cast<char*>(ws) + off); }

template <int BM, int BN, int BK>
void launch_gemm_tile(const GemmArgs& g, cudaStream_t st) {
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, (g.reduce ? 1 : g.nbatch) * g.ksplit);
  MVIN_LAUNCH((gemm_kernel<BM, BN, BK>), grid, GEMM_THREADS, 0, st, g);
}

// Tile choice: these GEMMs are tall and skinny and tiny next to the gather kernels, so pick the tile that yields
// enough CTAs to cover the SMs rather than the one with the best reuse.
int run_gemm(mvin_handle_t h, cudaStream_t st, const GemmArgs& g, const char* name = "gemm") {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return MVIN_OK;
  const long z = (long)(g.reduce ? 1 : g.nbatch) * g.ksplit;
  if (g.N <= 32) {
    launch_gemm_tile<32, 32, 32>(g, st);
  } else {
    const long ctas64 = (long)((g.M + 63) / 64) * ((g.N + 63) / 64) * z;
    if (ctas64 < 2L * h->sm_count) launch_gemm_tile<16, 64, 32>(g, st); else launch_gemm_tile<64, 64, 16>(g, st);
  }
  LAUNCH_CHECK(h, name);
  return MVIN_OK;
}

GemmArgs gemm_args() {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.nbatch = 1;
  g.ksplit = 1;
  g.alpha = 1.f;
  return g;
}

int pick_ksplit(long K) {
  long s = K / 128;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return (int)s;
}

template <typename KernelT>
int set_smem(KernelT k, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail(MVIN_ERR_CUDA, "cudaFuncSetAttribute(%zu B): %s", bytes, cudaGetErrorString(e));
  }
  // the persistent row kernels want as many co-resident CTAs as shared memory allows: without this hint the driver
  // sizes the L1 / shared split for ONE block of a large-footprint kernel
  if (bytes > 16 * 1024)
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  return MVIN_OK;
}

// Resident CTAs per SM of a kernel at a given dynamic shared-memory size (registers, threads and shared memory all
// taken into account by the occupancy calculator), capped by the handle's limit: the row kernels are persistent, so
// CTAs beyond the resident set only add prologue / epilogue work (weight loads, dW flushes).
template <typename KernelT>
int resident_ctas(mvin_handle_t h, KernelT k, int threads, size_t smem_bytes) {
  struct Key { const void* f; size_t s; int n; };
  static thread_local std::vector<Key> cache;
  for (const Key& e : cache)
    if (e.f == (const void*)k && e.s == smem_bytes) return e.n < h->max_ctas_per_sm ? e.n : h->max_ctas_per_sm;
  int n = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, threads, smem_bytes) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = 1;
  }
  cache.push_back({(const void*)k, smem_bytes, n});
  return n < h->max_ctas_per_sm ? n : h->max_ctas_per_sm;
}

// Split a grid of at most `cap` CTAs between levels in proportion to their tile counts (>= 1 CTA per level), then
// shrink each level's share so that all its CTAs walk the same number of tiles (+-1).  Returns the grid size and
// fills cta_end[].
int partition_grid(const long* rows, int nlev, int tile_rows, int cap, int* cta_end) {
  long tiles[MAX_LV], tot = 0;
  for (int l = 0; l < nlev; ++l) {
    tiles[l] = (rows[l] + tile_rows - 1) / tile_rows;
    if (tiles[l] < 1) tiles[l] = 1;
    tot += tiles[l];
  }
  int end = 0;
  for (int l = 0; l < nlev; ++l) {
    long n = tot <= cap ? tiles[l] : (long)((double)cap * (double)tiles[l] / (double)tot);
    if (n < 1) n = 1;
    if (n > tiles[l]) n = tiles[l];
    const long rounds = (tiles[l] + n - 1) / n;
    n = (tiles[l] + rounds - 1) / rounds;
    end += (int)n;
    cta_end[l] = end;
  }
  return end;
}

// tcgen05 versions of the forward row kernels (level_tc.cuh): 128-row tiles and a heavier prologue pay off only on
// large levels (measured: +8 % on transform_fwd at C3, parity at C4, slower at C2); env MVIN_B200_TC=0 / 2 = never / always
inline bool use_tc_path(mvin_handle_t h, long leaf_rows) {
  if (h->tc_mode == 0) return false;
  if (h->tc_mode == 2) return true;
  return leaf_rows >= 131072;
}

// activation buffers of a level that dwarf L2 (126 MB) are accessed with streaming hints (common.cuh, ld4a / st4a)
inline int stream_level(mvin_handle_t h, long rows, int D) {
  if (h->stream_mode >= 0) return h->stream_mode;
  return (size_t)rows * D * sizeof(float) >= ((size_t)48 << 20) ? 1 : 0;
}

// Tile list of one aggregator launch (level.cuh, TileList): returns the grid size (every CTA is resident).
int make_tile_list(TileList& tl, const long* rows, int nlev, int tile_rows, int cap, int* ctr) {
  long end = 0;
  for (int l = 0; l < nlev; ++l) {
    end += (rows[l] + tile_rows - 1) / tile_rows;
    tl.tile_end[l] = end;
  }
  tl.nlev = nlev;
  tl.ctr = ctr;
  return (int)(end < cap ? end : cap);
}

// ------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------
template <int D>
int forward_impl(mvin_handle_t h, const int64_t* item, const int32_t* mem_h, const int32_t* mem_r,
                 const int32_t* mem_t, int B, float* scores, float* scores_norm, void* ws, cudaStream_t st) {
  using C = TC<D>;
  const mvin_config_t& c = h->cfg;
  const int K = c.neighbor_sample_size, H = c.h_hop, p = c.p_hop, m = c.n_memory, nr = c.n_relation;
  const Layout L = make_layout(c, B, use_entity_leaf(c, B, h->n_shards, h->entity_leaf_mode));
  const mvin_params_t& P = h->P;
  int rc;
  prof_mark(h, st, nullptr);

  // side stream: integer expansion (model.py:243-256; level L ids are never materialised), relation scores and the
  // per-entity leaf aggregate are independent of the user side that runs on the launch stream meanwhile
  const Par par{h, st, h->use_streams && !h->prof_on};
  int32_t* stamp = L.entity_leaf ? at<int32_t>(ws, L.stamp) : nullptr;
  if (par.on && h->pre_fork && h->in_host_step) {
    CUDA_TRY(cudaStreamWaitEvent(h->side[0], h->ev_item, 0));   // only the item ids are needed on this branch
    h->pre_fork = false;
  } else {
    par.fork(0);
  }
  {
    cudaStream_t st = par.s(0);
    if (stamp) CUDA_TRY(cudaMemsetAsync(stamp, 0, sizeof(int32_t) * (size_t)c.n_entity, st));
    MVIN_LAUNCH((seed_kernel), (unsigned)((B + 255) / 256), 256, 0, st, item, B, at<int32_t>(ws, L.ent[0]), H == 1 ? stamp : nullptr);
    LAUNCH_CHECK(h, "seed");
    for (int lv = 0; lv + 1 < H; ++lv) {
      const long n = L.rows[lv] * K;
      MVIN_LAUNCH((expand_kernel), (unsigned)((n + 255) / 256), 256, 0, st, at<int32_t>(ws, L.ent[lv]), h->adj, L.rows[lv], K,
                                                                 at<int32_t>(ws, L.ent[lv + 1]),
                                                                 lv + 1 == H - 1 ? stamp : nullptr);
      LAUNCH_CHECK(h, "expand");
    }
    // relation scores of every aggregator
    {
      const int warps = H * nr;
      MVIN_LAUNCH((rel_scores_kernel), (warps * 32 + 255) / 256, 256, 0, st, P.relation_emb, P.agg_urh_w, nr, D, H,
                                                                 at<float>(ws, L.s));
      LAUNCH_CHECK(h, "rel_scores");
    }
    // entity mode: S_e for every distinct depth-(L-1) entity of the batch
    if (L.entity_leaf) {
      LeafEntArgs a;
      memset(&a, 0, sizeof(a));
      a.stamp = stamp; a.adj = h->adj; a.s = at<float>(ws, L.s); a.E = h->etab; a.Se = at<float>(ws, L.Se);
      a.n_entity = c.n_entity; a.K = K; a.n_rel = nr;
      const size_t sm = leaf_entity_smem(nr);
      if ((rc = set_smem(leaf_entity_kernel<D, false>, sm))) return rc;
      const long want = ((long)c.n_entity + LEAF_NW * 32 - 1) / (LEAF_NW * 32);
      const long cap = (long)h->sm_count * 8;
      MVIN_LAUNCH((leaf_entity_kernel<D, false>), (unsigned)(want < cap ? want : cap), LEAF_NT, sm, st, a);
      LAUNCH_CHECK(h, "leaf_entity_fwd");
    }
  }
  // the user side in one launch: seeds v = E[item], Q = RK^T v, ripple attention, user MLP -> user_o
  // (model.py:125-134, :161-240).  With a large relation-KGE table Q comes from a batched GEMM instead.
  const bool q_fused = user_q_fused(D, nr);
  if (!q_fused && p > 0) {
    const long n = (long)B * C::LPR;
    MVIN_LAUNCH((prep_items_kernel<D>), (unsigned)((n + 255) / 256), 256, 0, st, item, h->etab, B, nullptr,
                                                                      at<float>(ws, L.Vbuf), nullptr);
    LAUNCH_CHECK(h, "prep_items");
  }
  // Q[b, r, :] = RK[r]^T v_b      (model.py:211-220 refactored)
  if (!q_fused && p > 0) {
    GemmArgs g = gemm_args();
    g.A = at<float>(ws, L.Vbuf); g.sa_m = D; g.sa_k = 1; g.bsA = 0;
    g.B = P.relation_kge; g.sb_k = D; g.sb_n = 1; g.bsB = (long)D * D;
    g.C = at<float>(ws, L.Q); g.ldc = (long)nr * D; g.bsC = D;
    g.M = B; g.N = D; g.K = D; g.nbatch = nr;
    if ((rc = run_gemm(h, st, g, "gemm_q"))) return rc;
  }
  {
    UserArgs a;
    memset(&a, 0, sizeof(a));
    a.E = h->etab; a.item = item; a.RK = P.relation_kge; a.w_hi = P.h_item_w;
    a.W_user = P.user_mlp_w; a.b_user = P.user_mlp_b;
    a.mem_h = mem_h; a.mem_r = mem_r; a.mem_t = mem_t;
    a.Vbuf = at<float>(ws, L.Vbuf); a.Q = at<float>(ws, L.Q); a.probs = at<float>(ws, L.probs);
    a.O = at<float>(ws, L.O); a.u = at<float>(ws, L.u);
    a.B = B; a.m = m; a.p = p; a.n_rel = nr; a.q_ready = q_fused ? 0 : 1;
    const int PB = user_pairs_per_cta(D, nr, p, m, h->user_pb_fwd);
    const size_t sm = user_fwd_smem(D, PB, nr, p, m);
    const int nt = 32 * user_warps(PB, p);
    const unsigned grid = (unsigned)((B + PB - 1) / PB);
    switch (PB) {
      case 4:
        if ((rc = set_smem(user_fwd_kernel<D, 4>, sm))) return rc;
        MVIN_LAUNCH((user_fwd_kernel<D, 4>), grid, nt, sm, st, a);
        break;
      case 2:
        if ((rc = set_smem(user_fwd_kernel<D, 2>, sm))) return rc;
        MVIN_LAUNCH((user_fwd_kernel<D, 2>), grid, nt, sm, st, a);
        break;
      default:
        if ((rc = set_smem(user_fwd_kernel<D, 1>, sm))) return rc;
        MVIN_LAUNCH((user_fwd_kernel<D, 1>), grid, nt, sm, st, a);
    }
    LAUNCH_CHECK(h, "user_fwd");
  }
  par.join(0);
  // user-oriented transform of levels 0..L-1, one launch   (model.py:270-283)
  {
    const size_t sm = transform_fwd_smem<D>();
    if ((rc = set_smem(transform_fwd_kernel<D>, sm))) return rc;
    TransformArgs a;
    memset(&a, 0, sizeof(a));
    long rows[MAX_LV];
    for (int lv = 0; lv < H; ++lv) {
      TransformLevel& t = a.lv[lv];
      t.ent = at<int32_t>(ws, L.ent[lv]);
      t.W = P.transfer_w + (long)lv * D * D; t.b = P.transfer_b + (long)lv * D;
      t.T = at<float>(ws, L.V[0][lv]);
      t.rows = rows[lv] = L.rows[lv]; t.rpp = (int)(L.rows[lv] / B); t.rpp_magic = div_magic(t.rpp);
        t.stream = stream_level(h, L.rows[lv], D);
    }
    a.nlev = H; a.E = h->etab; a.u = at<float>(ws, L.u);
    bool done = false;
    if constexpr (D == 32 || D == 64) {
      if (use_tc_path(h, L.rows[H - 1])) {
        const size_t smt = transform_fwd_tc_smem<D>();
        if ((rc = set_smem(transform_fwd_tc_kernel<D>, smt))) return rc;
        const int grid = partition_grid(rows, H, TT<D>::R,
                                        h->sm_count * resident_ctas(h, transform_fwd_tc_kernel<D>, TT<D>::NT, smt), a.cta_end);
        MVIN_LAUNCH((transform_fwd_tc_kernel<D>), grid, TT<D>::NT, smt, st, a);
        done = true;
      }
    }
    if (!done) {
      const int grid = partition_grid(rows, H, C::R, h->sm_count * resident_ctas(h, transform_fwd_kernel<D>, C::NT, sm), a.cta_end);
      MVIN_LAUNCH((transform_fwd_kernel<D>), grid, C::NT, sm, st, a);
    }
    LAUNCH_CHECK(h, "transform_fwd");
  }
  // aggregation iterations (model.py:286-307): one launch per iteration, every level of it
  {
    const size_t sm_leaf = agg_fwd_smem<D, true>(K, nr), sm_in = agg_fwd_smem<D, false>(K, nr);
    if ((rc = set_smem(agg_fwd_kernel<D, true>, sm_leaf))) return rc;
    if ((rc = set_smem(agg_fwd_kernel<D, false>, sm_in))) return rc;
    static const char* names[MAX_L] = {"agg_fwd_0", "agg_fwd_1", "agg_fwd_2"};
    for (int i = 0; i < H; ++i) {
      AggArgs a;
      memset(&a, 0, sizeof(a));
      long rows[MAX_LV];
      const int nlev = H - i;
      // tile order: the levels with the most expensive tiles first (the per-entity leaf mode has the cheapest)
      const bool leaf_last = (i == 0 && L.entity_leaf);
      for (int q = 0; q < nlev; ++q) {
        const int lv = leaf_last ? q : nlev - 1 - q;
        AggLevel& t = a.lv[q];
        t.ent = at<int32_t>(ws, L.ent[lv]);
        t.self = at<float>(ws, L.V[i][lv]);
        t.Y = at<float>(ws, L.Y[i][lv]); t.V = at<float>(ws, L.V[i + 1][lv]);
        t.rows = rows[q] = L.rows[lv]; t.rpp = (int)(L.rows[lv] / B); t.rpp_magic = div_magic(t.rpp);
        t.stream = stream_level(h, L.rows[lv], D);
        t.leaf = (i == 0 && lv == H - 1);
        if (t.leaf) t.SU = at<float>(ws, L.SU); else t.child = at<float>(ws, L.V[i][lv + 1]);
      }
      a.adj = h->adj; a.s = at<float>(ws, L.s) + (long)i * nr;
      a.Wa = P.agg_w + (long)i * D * D; a.ba = P.agg_b + (long)i * D;
      a.K = K; a.n_rel = nr;
      if (i == 0) {
        a.E = h->etab; a.u = at<float>(ws, L.u);
        a.Se = L.entity_leaf ? at<float>(ws, L.Se) : nullptr;
        a.Wt = P.transfer_w + (long)H * D * D; a.bt = P.transfer_b + (long)H * D;
      }
      bool done = false;
      if constexpr (D == 32 || D == 64) {
        if (i == 0 && use_tc_path(h, L.rows[H - 1])) {     // leaf iteration; inner-only iterations are faster on mma.sync
          const size_t smt = agg_fwd_tc_smem<D>(i == 0, K, nr);
          if (i == 0) {
            if ((rc = set_smem(agg_fwd_tc_kernel<D, true>, smt))) return rc;
            const int grid = make_tile_list(a.tl, rows, nlev, TT<D>::R,
                                            h->sm_count * resident_ctas(h, agg_fwd_tc_kernel<D, true>, TT<D>::NT, smt), h->d_sched);
            MVIN_LAUNCH((agg_fwd_tc_kernel<D, true>), grid, TT<D>::NT, smt, st, a);
          } else {
            if ((rc = set_smem(agg_fwd_tc_kernel<D, false>, smt))) return rc;
            const int grid = make_tile_list(a.tl, rows, nlev, TT<D>::R,
                                            h->sm_count * resident_ctas(h, agg_fwd_tc_kernel<D, false>, TT<D>::NT, smt), h->d_sched);
            MVIN_LAUNCH((agg_fwd_tc_kernel<D, false>), grid, TT<D>::NT, smt, st, a);
          }
          done = true;
        }
      }
      if (done) {
      } else if (i == 0) {
        const int grid = make_tile_list(a.tl, rows, nlev, C::R,
                                        h->sm_count * resident_ctas(h, agg_fwd_kernel<D, true>, C::NT, sm_leaf), h->d_sched);
        MVIN_LAUNCH((agg_fwd_kernel<D, true>), grid, C::NT, sm_leaf, st, a);
      } else {
        const int grid = make_tile_list(a.tl, rows, nlev, C::R,
                                        h->sm_count * resident_ctas(h, agg_fwd_kernel<D, false>, C::NT, sm_in), h->d_sched);
        MVIN_LAUNCH((agg_fwd_kernel<D, false>), grid, C::NT, sm_in, st, a);
      }
      LAUNCH_CHECK(h, names[i]);
    }
  }
  // wide&deep mix + score (model.py:309-315, :158-159) in one launch: item = concat(V[0][0] .. V[H][0]) . W_mix + b_mix,
  // score = u . item   (the level-0 slices V[0..H][0] are contiguous)
  {
    constexpr int RT = 256 / C::LPR;
    const size_t sm = sizeof(float) * ((size_t)RT * (H + 1) * D + (D <= 64 ? (size_t)(H + 1) * D * D : 0));
    if ((rc = set_smem(mix_score_kernel<D>, sm))) return rc;
    MVIN_LAUNCH((mix_score_kernel<D>), (unsigned)((B + RT - 1) / RT), 256, sm, st, at<float>(ws, L.V[0][0]), P.mix_w, P.mix_b,
                                                                        at<float>(ws, L.u), B, H + 1, at<float>(ws, L.item),
                                                                        at<float>(ws, L.scores), scores_norm);
    LAUNCH_CHECK(h, "mix_score");
    if (scores) CUDA_TRY(cudaMemcpyAsync(scores, at<float>(ws, L.scores), sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
  }
  return MVIN_OK;
}

// ------------------------------------------------------------------------------------------------------
// backward, part 0: everything that depends on the parameters only -- zeroed accumulators, gradient buffers
// initialised with their dense L2 terms, transposed weights.  `mid` (optional) is recorded once the part the first
// backward kernels need is enqueued.  The host-step entry points run it on a side stream while the feed is still
// crossing the bus.
// ------------------------------------------------------------------------------------------------------
template <int D>
int backward_init(mvin_handle_t h, int B, void* ws, cudaStream_t st, cudaEvent_t mid, bool zero_small) {
  const mvin_config_t& c = h->cfg;
  const int H = c.h_hop, p = c.p_hop, nr = c.n_relation;
  const Layout L = make_layout(c, B, use_entity_leaf(c, B, h->n_shards, h->entity_leaf_mode));
  const mvin_params_t& P = h->P;
  const mvin_params_t& G = h->G;
  const float l2w = c.l2_weight, l2a = c.l2_agg_weight;
  float* acc = at<float>(ws, L.acc);
  float* wT = at<float>(ws, L.wT);
  if (zero_small) CUDA_TRY(cudaMemsetAsync(at<char>(ws, L.zero_begin), 0, L.zero_mid - L.zero_begin, st));
  // dense L2 terms: initialise every other gradient buffer with coef * param (model.py:388-410)
  {
    L2Segments sg;
    memset(&sg, 0, sizeof(sg));
    int n = 0;
    auto add = [&](const float* prm, float* grd, long cnt, float coef, float mult, int which) {
      sg.param[n] = prm; sg.grad[n] = grd; sg.n[n] = cnt; sg.coef[n] = coef * mult * h->dense_l2_scale;
      sg.mult[n] = mult * h->dense_l2_scale; sg.which[n] = which;
      ++n;
    };
    const float pm = p > 0 ? 1.f : 0.f;
    add(P.user_emb, G.user_emb, (long)c.n_user * D, l2a, 1.f, 1);                 // model.py:392
    add(P.relation_emb, G.relation_emb, (long)nr * D, l2w, 1.f, 0);               // :388
    add(P.relation_kge, G.relation_kge, (long)nr * D * D, 0.f, 0.f, 0);
    add(P.mix_w, G.mix_w, (long)(H + 1) * D * D, l2a, 1.f, 1);                    // :400-401
    add(P.mix_b, G.mix_b, D, l2a, 1.f, 1);
    add(P.user_mlp_w, G.user_mlp_w, (long)(p + 1) * D * D, l2w, pm, 0);           // :404
    add(P.user_mlp_b, G.user_mlp_b, D, l2w, pm, 0);
    if (H > 0) {
      add(P.transfer_w, G.transfer_w, (long)H * D * D, l2w, pm, 0);               // :407-408
      add(P.transfer_b, G.transfer_b, (long)H * D, l2w, pm, 0);
    }
    add(P.transfer_w + (long)H * D * D, G.transfer_w + (long)H * D * D, (long)D * D, l2w, 2.f * pm, 0);   // :405 + :408
    add(P.transfer_b + (long)H * D, G.transfer_b + (long)H * D, D, l2w, 2.f * pm, 0);
    add(P.h_item_w, G.h_item_w, 2 * D, l2w, 1.f, 0);                              // :410
    add(P.h_item_b, G.h_item_b, 1, l2w, 1.f, 0);
    add(P.agg_w, G.agg_w, (long)H * D * D, l2a, 1.f, 1);                          // :394-396
    add(P.agg_b, G.agg_b, (long)H * D, 0.f, 0.f, 1);
    add(P.agg_urh_w, G.agg_urh_w, (long)H * 3 * D, l2a, 1.f, 1);
    add(P.agg_urh_b, G.agg_urh_b, H, 0.f, 0.f, 1);
    sg.count = n;
    MVIN_LAUNCH((l2_dense_kernel), h->sm_count * 2, 256, 0, st, sg, acc);
    LAUNCH_CHECK(h, "l2_dense");
  }
  // transposed weights: wT[i] = W_a[i]^T (i < H), wT[H + e] = W_t[e]^T (e <= H)
  MVIN_LAUNCH((transpose_kernel), dim3(2 * H + 1), 256, 0, st, P.agg_w, P.transfer_w, H, D, wT);
  LAUNCH_CHECK(h, "transpose");
  if (mid) cudaEventRecord(mid, st);
  CUDA_TRY(cudaMemsetAsync(at<char>(ws, L.zero_mid), 0, L.zero_end - L.zero_mid, st));
  // sharded mode: peers scatter into this rank's shard, so the CALLER zeroes it (and synchronises the ranks)
  if (h->n_shards == 1) CUDA_TRY(cudaMemsetAsync(G.entity_emb, 0, sizeof(float) * (size_t)c.n_entity * D, st));
  prof_mark(h, st, "memset");
  return MVIN_OK;
}

// ------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------
template <int D>
int launch_dw(mvin_handle_t h, cudaStream_t st, const DwArgs& a, int groups, const char* name) {
  using C = TC<D>;
  const size_t sm = dw_smem<D>();
  int rc;
  if ((rc = set_smem(dw_kernel<D>, sm))) return rc;
  long tiles = (a.rows + C::R - 1) / C::R;
  int gx = (int)(tiles < h->sm_count ? tiles : h->sm_count);
  MVIN_LAUNCH((dw_kernel<D>), dim3(gx, groups), C::NT, sm, st, a);
  LAUNCH_CHECK(h, name);
  return MVIN_OK;
}

template <int D>
int backward_impl(mvin_handle_t h, const float* labels, int B, float* losses_out, void* ws, cudaStream_t st) {
  using C = TC<D>;
  const mvin_config_t& c = h->cfg;
  const int K = c.neighbor_sample_size, H = c.h_hop, p = c.p_hop, m = c.n_memory, nr = c.n_relation;
  const Layout L = make_layout(c, B, use_entity_leaf(c, B, h->n_shards, h->entity_leaf_mode));
  const mvin_params_t& P = h->P;
  const mvin_params_t& G = h->G;
  const float l2w = c.l2_weight, l2a = c.l2_agg_weight;
  int rc;
  float* acc = at<float>(ws, L.acc);
  float* wT = at<float>(ws, L.wT);
  prof_mark(h, st, nullptr);

  // part 0 (parameters only) on side stream 0 -- unless a host-step entry point already ran it during the feed copy
  const Par par{h, st, h->use_streams && !h->prof_on};
  if (h->early_init && h->in_host_step) {
    CUDA_TRY(cudaStreamWaitEvent(st, h->ev_early, 0));
    h->early_init = false;
    par.fork(0);
    par.mark_mid();
  } else {
    CUDA_TRY(cudaMemsetAsync(at<char>(ws, L.zero_begin), 0, L.zero_mid - L.zero_begin, st));   // loss accumulators
    par.fork(0);
    if ((rc = backward_init<D>(h, B, ws, par.s(0), par.on ? h->ev_mid : nullptr, false))) return rc;
  }
  {
  cudaStream_t st = par.s(0);
  // ripple-memory relation histogram (feeds the un-normalised L2 over gathered RK matrices, model.py:386)
  if (p > 0) {
    const long n = (long)p * B * m;
    MVIN_LAUNCH((hist_r_kernel), h->sm_count * 4, 256, sizeof(float) * nr, st, h->mem_r, n, nr, at<float>(ws, L.cnt));
    LAUNCH_CHECK(h, "hist_r");
  }
  }

  float* du = at<float>(ws, L.du);
  float* ditem = at<float>(ws, L.ditem);
  const float invB = 1.f / (float)(h->global_batch > 0 ? h->global_batch : B);
  const bool fused_mix_bwd = D <= 64;          // W_mix^T ((H+1) d^2 floats) is staged in shared memory
  if (fused_mix_bwd) {
    // loss gradient + mix backward in one launch: ditem, du, DC[j][0] = ditem . W_mix[j]^T
    const size_t sm = sizeof(float) * ((size_t)D * (H + 1) * D + 16 * D);
    if ((rc = set_smem(loss_mix_bwd_kernel<D>, sm))) return rc;
    const int grid = (B + 15) / 16 < 2 * h->sm_count ? (B + 15) / 16 : 2 * h->sm_count;
    MVIN_LAUNCH((loss_mix_bwd_kernel<D>), grid, 256, sm, st, at<float>(ws, L.scores), labels, at<float>(ws, L.u), at<float>(ws, L.item),
                                                  P.mix_w, B, H + 1, invB, ditem, du, at<float>(ws, L.DC[0][0]), acc);
    LAUNCH_CHECK(h, "loss_mix_bwd");
  } else {
    const long n = (long)B * C::LPR;
    MVIN_LAUNCH((loss_bwd_kernel<D>), (unsigned)((n + 255) / 256), 256, 0, st, at<float>(ws, L.scores), labels, at<float>(ws, L.u),
                                                                    at<float>(ws, L.item), B, invB, ditem, du, acc);
    LAUNCH_CHECK(h, "loss_bwd");
  }
  par.wait_mid();
  // mix backward: dW_mix[j] = V[j][0]^T ditem (grouped), DC[j][0] = ditem . W_mix[j]^T (batched)
  {
    DwArgs a;
    memset(&a, 0, sizeof(a));
    for (int j = 0; j <= H; ++j) { a.A[j] = at<float>(ws, L.V[j][0]); a.lda[j] = D; a.dW[j] = G.mix_w + (long)j * D * D; }
    a.G = ditem; a.db = G.mix_b; a.rows = B;
    par.fork(1);                                   // side stream 1: weight gradient of the mix layer
    if ((rc = launch_dw<D>(h, par.s(1), a, H + 1, "dw_mix"))) return rc;
    if (!fused_mix_bwd) {
      GemmArgs g = gemm_args();
      g.A = ditem; g.sa_m = D; g.sa_k = 1; g.bsA = 0;
      g.B = P.mix_w; g.sb_k = 1; g.sb_n = D; g.bsB = (long)D * D;   // W_mix[jD + n][k] -> transposed use
      g.C = at<float>(ws, L.DC[0][0]); g.ldc = D; g.bsC = (long)B * D;
      g.M = B; g.N = D; g.K = D; g.nbatch = H + 1;
      if ((rc = run_gemm(h, st, g, "gemm_mix_bwd"))) return rc;
    }
  }
  // aggregation iterations, reversed; one launch per iteration
  {
    const size_t sm_leaf = agg_bwd_smem<D, true>(K, nr), sm_in = agg_bwd_smem<D, false>(K, nr);
    if ((rc = set_smem(agg_bwd_kernel<D, true>, sm_leaf))) return rc;
    if ((rc = set_smem(agg_bwd_kernel<D, false>, sm_in))) return rc;
    static const char* names[MAX_L] = {"agg_bwd_0", "agg_bwd_1", "agg_bwd_2"};
    for (int i = H - 1; i >= 0; --i) {
      AggBwdArgs a;
      memset(&a, 0, sizeof(a));
      long rows[MAX_LV];
      const int nlev = H - i;
      const bool leaf_last = (i == 0 && L.entity_leaf);
      for (int q = 0; q < nlev; ++q) {
        const int lv = leaf_last ? q : nlev - 1 - q;
        AggBwdLevel& t = a.lv[q];
        t.ent = at<int32_t>(ws, L.ent[lv]);
        t.V = at<float>(ws, L.V[i + 1][lv]); t.Y = at<float>(ws, L.Y[i][lv]);
        t.g1 = at<float>(ws, L.DC[i + 1][lv]);
        t.g2 = has_agg(H, i + 1, lv) ? at<float>(ws, L.DS[i + 1][lv]) : nullptr;
        t.dself = at<float>(ws, L.DS[i][lv]);
        t.rows = rows[q] = L.rows[lv]; t.rpp = (int)(L.rows[lv] / B); t.rpp_magic = div_magic(t.rpp);
        t.stream = stream_level(h, L.rows[lv], D);
        t.leaf = (i == 0 && lv == H - 1);
        if (t.leaf) {
          t.SU = at<float>(ws, L.SU);
        } else {
          t.child = at<float>(ws, L.V[i][lv + 1]);
          t.dchild = at<float>(ws, L.DC[i][lv + 1]);
        }
      }
      a.adj = h->adj; a.s = at<float>(ws, L.s) + (long)i * nr;
      a.WaT = wT + (long)i * D * D;
      a.dWa = G.agg_w + (long)i * D * D; a.dba = G.agg_b + (long)i * D;
      a.ds = at<float>(ws, L.ds) + (long)i * nr;
      a.K = K; a.n_rel = nr;
      if (i == 0) {
        par.join(0);       // zeroed dE / GSe / dQ / cnt are first needed here
        a.E = h->etab; a.WtT = wT + (long)(H + H) * D * D;
        a.dWt = G.transfer_w + (long)H * D * D; a.dbt = G.transfer_b + (long)H * D;
        a.dE = h->gtab; a.du = du;
        a.GSe = L.entity_leaf ? at<float>(ws, L.GSe) : nullptr;
        const int grid = make_tile_list(a.tl, rows, nlev, C::R,
                                        h->sm_count * resident_ctas(h, agg_bwd_kernel<D, true>, C::NT, sm_leaf), h->d_sched + 2);
        MVIN_LAUNCH((agg_bwd_kernel<D, true>), grid, C::NT, sm_leaf, st, a);
      } else {
        const int grid = make_tile_list(a.tl, rows, nlev, C::R,
                                        h->sm_count * resident_ctas(h, agg_bwd_kernel<D, false>, C::NT, sm_in), h->d_sched + 2);
        MVIN_LAUNCH((agg_bwd_kernel<D, false>), grid, C::NT, sm_in, st, a);
      }
      LAUNCH_CHECK(h, names[i]);
    }
  }
  // side stream 0: per-entity leaf backward + relation-score gradients, while the user-oriented transform backward
  // runs on the launch stream (both only add into dE)
  par.fork(0);
  {
  cudaStream_t st = par.s(0);
  if (L.entity_leaf) {
    LeafEntArgs a;
    memset(&a, 0, sizeof(a));
    a.stamp = at<int32_t>(ws, L.stamp); a.adj = h->adj; a.s = at<float>(ws, L.s); a.E = h->etab;
    a.GSe = at<float>(ws, L.GSe); a.dE = h->gtab; a.ds = at<float>(ws, L.ds);
    a.n_entity = c.n_entity; a.K = K; a.n_rel = nr;
    const size_t sm = leaf_entity_smem(nr);
    if ((rc = set_smem(leaf_entity_kernel<D, true>, sm))) return rc;
    const long want = ((long)c.n_entity + LEAF_NW * 32 - 1) / (LEAF_NW * 32);
    const long cap = (long)h->sm_count * 8;
    MVIN_LAUNCH((leaf_entity_kernel<D, true>), (unsigned)(want < cap ? want : cap), LEAF_NT, sm, st, a);
    LAUNCH_CHECK(h, "leaf_entity_bwd");
  }
  MVIN_LAUNCH((rel_scores_bwd_kernel), H, 128, 0, st, P.relation_emb, P.agg_urh_w, at<float>(ws, L.ds), nr, D, G.relation_emb,
                                           G.agg_urh_w);
  LAUNCH_CHECK(h, "rel_scores_bwd");
  }
  // user-oriented transform backward, levels 0..L-1, one launch
  {
    const size_t sm = transform_bwd_smem<D>();
    if ((rc = set_smem(transform_bwd_kernel<D>, sm))) return rc;
    TransformArgs a;
    memset(&a, 0, sizeof(a));
    long rows[MAX_LV];
    for (int q = 0; q < H; ++q) {
      const int lv = H - 1 - q;
      TransformLevel& t = a.lv[q];
      t.ent = at<int32_t>(ws, L.ent[lv]);
      t.W = wT + (long)(H + lv) * D * D;
      t.g1 = at<float>(ws, L.DC[0][lv]); t.g2 = at<float>(ws, L.DS[0][lv]);
      t.dW = G.transfer_w + (long)lv * D * D; t.db = G.transfer_b + (long)lv * D;
      t.rows = rows[q] = L.rows[lv]; t.rpp = (int)(L.rows[lv] / B); t.rpp_magic = div_magic(t.rpp);
        t.stream = stream_level(h, L.rows[lv], D);
    }
    a.nlev = H; a.E = h->etab; a.u = at<float>(ws, L.u); a.dE = h->gtab; a.du = du;
    const int grid = partition_grid(rows, H, C::R, h->sm_count * resident_ctas(h, transform_bwd_kernel<D>, C::NT, sm), a.cta_end);
    MVIN_LAUNCH((transform_bwd_kernel<D>), grid, C::NT, sm, st, a);
    LAUNCH_CHECK(h, "transform_bwd");
  }
  // user_o = O . W_user + b  backward
  {
    DwArgs a;
    memset(&a, 0, sizeof(a));
    for (int s = 0; s <= p; ++s) {
      a.A[s] = at<float>(ws, L.O) + (long)s * D; a.lda[s] = (long)(p + 1) * D; a.dW[s] = G.user_mlp_w + (long)s * D * D;
    }
    a.G = du; a.db = G.user_mlp_b; a.rows = B;
    par.fork(1);                                   // side stream 1: weight gradient of the user MLP
    if ((rc = launch_dw<D>(h, par.s(1), a, p + 1, "dw_user"))) return rc;
    // dO = du . W_user^T as a batched GEMM (a per-warp matvec inside the ripple kernel re-reads W_user per warp and
    // measured slower: +9 us at C2, +78 us at C3)
    GemmArgs g = gemm_args();
    g.A = du; g.sa_m = D; g.sa_k = 1;
    g.B = P.user_mlp_w; g.sb_k = 1; g.sb_n = D;
    g.C = at<float>(ws, L.dO); g.ldc = (long)(p + 1) * D;
    g.M = B; g.N = (p + 1) * D; g.K = D;
    if ((rc = run_gemm(h, st, g, "gemm_user_bwd"))) return rc;
  }
  // ripple backward
  {
    RippleBwdArgs a;
    a.E = h->etab; a.Q = at<float>(ws, L.Q); a.w_hi = P.h_item_w;
    a.mem_h = h->mem_h; a.mem_r = h->mem_r; a.mem_t = h->mem_t;
    a.probs = at<float>(ws, L.probs); a.dO = at<float>(ws, L.dO);
    a.dE = h->gtab; a.dQ = at<float>(ws, L.dQ); a.dw_hi = G.h_item_w; a.l2_acc = acc + 1;
    a.l2_weight = l2w; a.B = B; a.m = m; a.p = p; a.n_rel = nr;
    const size_t sm = ripple_bwd_smem(m, D);
    if ((rc = set_smem(ripple_bwd_kernel<D>, sm))) return rc;
    a.ctr = h->d_sched + 4;
    const long warps = (long)B * (p + 1);
    long grid = (warps + RIPPLE_NW - 1) / RIPPLE_NW;
    const long resident = (long)h->sm_count * resident_ctas(h, ripple_bwd_kernel<D>, RIPPLE_NT, sm) * 2;   // cap 4 -> 8
    if (grid > resident) grid = resident;
    MVIN_LAUNCH((ripple_bwd_kernel<D>), (unsigned)grid, RIPPLE_NT, sm, st, a);
    LAUNCH_CHECK(h, "ripple_bwd");
  }
  if (p > 0) {
    // three independent consumers of dQ / cnt: RK L2 term (side 1), dRK (side 0), dE[item] (launch stream)
    par.fork(0);
    par.fork(1);
    MVIN_LAUNCH((rk_l2_kernel), nr, 256, 0, par.s(1), P.relation_kge, at<float>(ws, L.cnt), D * D, 2.f * l2w, G.relation_kge, acc + 1);
    LAUNCH_CHECK(h, "rk_l2");
    // dRK[r][i][j] += sum_b v[b][i] dQ[b][r][j]
    GemmArgs g = gemm_args();
    g.A = at<float>(ws, L.Vbuf); g.sa_m = 1; g.sa_k = D; g.bsA = 0;
    g.B = at<float>(ws, L.dQ); g.sb_k = (long)nr * D; g.sb_n = 1; g.bsB = D;
    g.C = G.relation_kge; g.ldc = D; g.bsC = (long)D * D;
    g.M = D; g.N = D; g.K = B; g.nbatch = nr; g.ksplit = pick_ksplit(B); g.accumulate = 1;
    if ((rc = run_gemm(h, par.s(0), g, "gemm_drk"))) return rc;
    // dv[b][i] = sum_r sum_j dQ[b][r][j] RK[r][i][j]  (reduced over r inside the CTA);  dE[item_b] += dv[b]
    GemmArgs g2 = gemm_args();
    g2.A = at<float>(ws, L.dQ); g2.sa_m = (long)nr * D; g2.sa_k = 1; g2.bsA = D;
    g2.B = P.relation_kge; g2.sb_k = 1; g2.sb_n = D; g2.bsB = (long)D * D;
    g2.M = B; g2.N = D; g2.K = D; g2.nbatch = nr; g2.reduce = 1;
    g2.ksplit = nr >= 8 ? 4 : 1;
    if (h->n_shards == 1) {
      // accumulate straight into the entity-table gradient rows of the items
      g2.C = G.entity_emb; g2.ldc = D; g2.bsC = 0; g2.c_rows = at<int32_t>(ws, L.ent[0]); g2.accumulate = 1;
      if ((rc = run_gemm(h, st, g2, "gemm_dv"))) return rc;
    } else {
      g2.C = at<float>(ws, L.dv); g2.ldc = D; g2.bsC = 0; g2.accumulate = g2.ksplit > 1;   // dv lives in the zeroed region
      if ((rc = run_gemm(h, st, g2, "gemm_dv"))) return rc;
      const long n = (long)B * C::LPR;
      MVIN_LAUNCH((scatter_rows_kernel<D>), (unsigned)((n + 255) / 256), 256, 0, st, at<float>(ws, L.dv), at<int32_t>(ws, L.ent[0]), B,
                                                                          h->gtab);
      LAUNCH_CHECK(h, "scatter_dv");
    }
  }
  par.join(0);
  par.join(1);
  MVIN_LAUNCH((finalize_loss_kernel), 1, 32, 0, st, acc, l2w, l2a, losses_out);
  LAUNCH_CHECK(h, "finalize_loss");
  return MVIN_OK;
}

int check_supported(const mvin_config_t* c) {
  if (c->flags != MVIN_FLAGS_ALL)
    return fail(MVIN_ERR_UNSUPPORTED, "only --ablation all (flags 0x1f) is supported, got 0x%x", c->flags);
  if (c->n_mix_hop != 1) return fail(MVIN_ERR_UNSUPPORTED, "n_mix_hop must be 1, got %d", c->n_mix_hop);
  if (c->h_hop < 1 || c->h_hop > MAX_L) return fail(MVIN_ERR_UNSUPPORTED, "h_hop must be in 1..3, got %d", c->h_hop);
  const int d = c->dim;
  if (!(d == 8 || d == 16 || d == 32 || d == 64 || d == 128))
    return fail(MVIN_ERR_UNSUPPORTED, "dim must be one of 8,16,32,64,128, got %d", d);
  if (c->neighbor_sample_size < 1 || c->neighbor_sample_size > MAX_K)
    return fail(MVIN_ERR_UNSUPPORTED, "neighbor_sample_size must be in 1..64, got %d", c->neighbor_sample_size);
  if (c->p_hop < 0 || c->p_hop > 8) return fail(MVIN_ERR_UNSUPPORTED, "p_hop must be in 0..8, got %d", c->p_hop);
  if (c->n_memory < 1 || c->n_memory > 4096)
    return fail(MVIN_ERR_UNSUPPORTED, "n_memory must be in 1..4096, got %d", c->n_memory);
  if (c->n_user < 1 || c->n_entity < 1 || c->n_relation < 1 || c->n_relation > 4096)
    return fail(MVIN_ERR_INVALID, "bad table sizes (n_user %d, n_entity %d, n_relation %d)", c->n_user, c->n_entity,
                c->n_relation);
  if (c->max_batch < 1) return fail(MVIN_ERR_INVALID, "max_batch must be >= 1");
  if (user_pairs_per_cta(d, c->n_relation, c->p_hop, c->n_memory) == 0)
    return fail(MVIN_ERR_UNSUPPORTED,
                "n_relation x dim (%d x %d) with n_memory %d does not fit the user-side kernel's shared memory",
                c->n_relation, d, c->n_memory);
  return MVIN_OK;
}

#define DISPATCH_D(d, CALL)                                           \
  switch (d) {                                                        \
    case 8: { constexpr int DD = 8; return CALL; }                    \
    case 16: { constexpr int DD = 16; return CALL; }                  \
    case 32: { constexpr int DD = 32; return CALL; }                  \
    case 64: { constexpr int DD = 64; return CALL; }                  \
    case 128: { constexpr int DD = 128; return CALL; }                \
    default: return fail(MVIN_ERR_UNSUPPORTED, "dim %d", d);          \
  }

int dispatch_forward(mvin_handle_t h, const int64_t* item, const int32_t* mh, const int32_t* mr, const int32_t* mt,
                     int B, float* scores, float* sn, void* ws, cudaStream_t st) {
  DISPATCH_D(h->cfg.dim, (forward_impl<DD>(h, item, mh, mr, mt, B, scores, sn, ws, st)));
}
struct HostStepGuard {
  mvin_handle_t h;
  explicit HostStepGuard(mvin_handle_t hh) : h(hh) { h->in_host_step = true; h->early_init = false; h->pre_fork = false; }
  ~HostStepGuard() { h->in_host_step = false; h->early_init = false; h->pre_fork = false; }
};

int dispatch_backward_init(mvin_handle_t h, int B, void* ws, cudaStream_t st) {
  DISPATCH_D(h->cfg.dim, (backward_init<DD>(h, B, ws, st, nullptr, true)));
}
// Host-step entry points: once the item ids are on the device the KG side of the forward can start, and the
// parameter-only part of the backward can run right away -- both overlap the H2D copy of the ripple memories.
int host_step_overlap(mvin_handle_t h, int B, void* ws, cudaStream_t st) {
  h->early_init = false;
  h->pre_fork = false;
  if (!h->use_streams || h->prof_on || !h->has_grads) return MVIN_OK;
  CUDA_TRY(cudaEventRecord(h->ev_item, st));
  h->pre_fork = true;
  CUDA_TRY(cudaStreamWaitEvent(h->side[1], h->ev_item, 0));
  int rc = dispatch_backward_init(h, B, ws, h->side[1]);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(h->ev_early, h->side[1]));
  h->early_init = true;
  return MVIN_OK;
}
int dispatch_backward(mvin_handle_t h, const float* labels, int B, float* losses, void* ws, cudaStream_t st) {
  DISPATCH_D(h->cfg.dim, (backward_impl<DD>(h, labels, B, losses, ws, st)));
}

}  // namespace

// ------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------
extern "C" {

int mvin_abi_version(void) { return MVIN_ABI_VERSION; }
const char* mvin_last_error(void) { return g_err; }

int mvin_create(const mvin_config_t* cfg, mvin_handle_t* out) {
  if (!cfg || !out) return fail(MVIN_ERR_INVALID, "null argument");
  int rc = check_supported(cfg);
  if (rc) return rc;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  if (prop.major < 10)
    return fail(MVIN_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major,
                prop.minor);
  mvin_handle_t h = new mvin_handle_s();
  h->cfg = *cfg;
  h->device = dev;
  h->sm_count = prop.multiProcessorCount;
  h->n_local_rows = cfg->n_entity;
  if (const char* ev = getenv("MVIN_B200_STREAMS")) h->use_streams = atoi(ev) != 0;
  for (int i = 0; i < 2; ++i) {
    if (cudaStreamCreateWithFlags(&h->side[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming) != cudaSuccess ||
        (i == 0 && (cudaEventCreateWithFlags(&h->ev_mid, cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&h->ev_early, cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&h->ev_item, cudaEventDisableTiming) != cudaSuccess))) {
      return fail(MVIN_ERR_CUDA, "stream / event creation: %s", cudaGetErrorString(cudaGetLastError()));
    }
  }
  if (const char* ev = getenv("MVIN_B200_ENTITY_LEAF")) h->entity_leaf_mode = atoi(ev) != 0 ? 1 : 0;
  if (const char* ev = getenv("MVIN_B200_USER_PB_FWD")) { const int n = atoi(ev); if (n == 1 || n == 2 || n == 4) h->user_pb_fwd = n; }
  if (const char* ev = getenv("MVIN_B200_TC")) { const int n = atoi(ev); if (n >= 0 && n <= 2) h->tc_mode = n; }
  if (const char* ev = getenv("MVIN_B200_STREAM")) h->stream_mode = atoi(ev) != 0 ? 1 : 0;
  if (const char* ev = getenv("MVIN_B200_CTAS_PER_SM")) { const int n = atoi(ev); if (n >= 1 && n <= 32) h->max_ctas_per_sm = n; }
  if (cudaMalloc(&h->d_shard_tab, sizeof(void*) * 2 * MAX_SHARDS) != cudaSuccess ||
      cudaMalloc(&h->d_sched, sizeof(int) * 16) != cudaSuccess ||
      cudaMemset(h->d_sched, 0, sizeof(int) * 16) != cudaSuccess) {
    delete h;
    return fail(MVIN_ERR_CUDA, "cudaMalloc(shard table / scheduler counters): %s", cudaGetErrorString(cudaGetLastError()));
  }
  *out = h;
  return MVIN_OK;
}

int mvin_destroy(mvin_handle_t h) {
  if (h && h->d_shard_tab) cudaFree(h->d_shard_tab);
  if (h && h->d_sched) cudaFree(h->d_sched);
  if (h && h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (int i = 0; h && i < 2; ++i) if (h->ev_feed[i]) cudaEventDestroy(h->ev_feed[i]);
  if (h && h->ev_free) cudaEventDestroy(h->ev_free);
  for (int i = 0; h && i < 2; ++i) {
    if (h->side[i]) cudaStreamDestroy(h->side[i]);
    if (h->ev_fork[i]) cudaEventDestroy(h->ev_fork[i]);
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    if (i == 0 && h->ev_mid) cudaEventDestroy(h->ev_mid);
    if (i == 0 && h->ev_early) cudaEventDestroy(h->ev_early);
    if (i == 0 && h->ev_item) cudaEventDestroy(h->ev_item);
  }
  delete h;
  return MVIN_OK;
}

int mvin_bind_params(mvin_handle_t h, const mvin_params_t* params) {
  if (!h || !params) return fail(MVIN_ERR_INVALID, "null argument");
  h->P = *params;
  h->has_params = true;
  if (h->n_shards == 1) h->etab = ETab{params->entity_emb, nullptr, 0, 0};
  return MVIN_OK;
}
int mvin_bind_grads(mvin_handle_t h, const mvin_params_t* grads) {
  if (!h || !grads) return fail(MVIN_ERR_INVALID, "null argument");
  h->G = *grads;
  h->has_grads = true;
  if (h->n_shards == 1) h->gtab = GTab{grads->entity_emb, nullptr, 0, 0};
  return MVIN_OK;
}
int mvin_bind_adjacency(mvin_handle_t h, const int32_t* adj_packed) {
  if (!h || !adj_packed) return fail(MVIN_ERR_INVALID, "null argument");
  h->adj = adj_packed;
  return MVIN_OK;
}

int mvin_pack_adjacency(const int64_t* adj_entity, const int64_t* adj_relation, int32_t n_entity, int32_t K,
                        int32_t* adj_packed, void* stream) {
  if (!adj_entity || !adj_relation || !adj_packed || n_entity < 1 || K < 1) return fail(MVIN_ERR_INVALID, "bad argument");
  const long n = (long)n_entity * K;
  MVIN_LAUNCH((pack_adj_kernel), (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, adj_entity, adj_relation, n_entity, K,
                                                                                adj_packed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MVIN_ERR_CUDA, "launch pack_adj: %s", cudaGetErrorString(e));
  return MVIN_OK;
}

int mvin_bind_entity_shards(mvin_handle_t h, int32_t n_shards, const float* const* entity_shards,
                            float* const* grad_shards) {
  if (!h || !entity_shards || !grad_shards) return fail(MVIN_ERR_INVALID, "null argument");
  if (n_shards < 1 || n_shards > MAX_SHARDS || (n_shards & (n_shards - 1)))
    return fail(MVIN_ERR_INVALID, "n_shards must be a power of two in 1..%d, got %d", MAX_SHARDS, n_shards);
  int shift = 0;
  while ((1 << shift) < n_shards) ++shift;
  void* host[2 * MAX_SHARDS] = {nullptr};
  for (int i = 0; i < n_shards; ++i) {
    if (!entity_shards[i] || !grad_shards[i]) return fail(MVIN_ERR_INVALID, "null shard pointer %d", i);
    host[i] = const_cast<float*>(entity_shards[i]);
    host[MAX_SHARDS + i] = grad_shards[i];
  }
  CUDA_TRY(cudaMemcpy(h->d_shard_tab, host, sizeof(host), cudaMemcpyHostToDevice));
  h->n_shards = n_shards;
  h->n_local_rows = ((long)h->cfg.n_entity + n_shards - 1) / n_shards;
  h->etab = ETab{nullptr, reinterpret_cast<const float* const*>(h->d_shard_tab), shift, n_shards - 1};
  h->gtab = GTab{nullptr, reinterpret_cast<float* const*>(h->d_shard_tab + MAX_SHARDS), shift, n_shards - 1};
  return MVIN_OK;
}

// ---- CUDA IPC plumbing for the peers' shards ---------------------------------------------------------------
namespace {
struct IpcOpened { cudaIpcMemHandle_t handle; void* base; };
std::vector<IpcOpened> g_ipc_opened;
std::mutex g_ipc_mutex;
}  // namespace

int mvin_ipc_export(const void* dev_ptr, void* handle_out, int64_t* offset_out) {
  if (!dev_ptr || !handle_out || !offset_out) return fail(MVIN_ERR_INVALID, "null argument");
  // base of the containing allocation (the caching allocator of the caller may sub-allocate): driver entry point
  // resolved at run time so that the library carries no link-time dependency on libcuda
  typedef int (*GetRangeFn)(unsigned long long*, size_t*, unsigned long long);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CUDA_TRY(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) return fail(MVIN_ERR_CUDA, "cuMemGetAddressRange not available");
  unsigned long long base = 0;
  size_t size = 0;
  const int drc = reinterpret_cast<GetRangeFn>(fn)(&base, &size, (unsigned long long)(uintptr_t)dev_ptr);
  if (drc != 0) return fail(MVIN_ERR_CUDA, "cuMemGetAddressRange failed (%d)", drc);
  cudaIpcMemHandle_t hd;
  CUDA_TRY(cudaIpcGetMemHandle(&hd, reinterpret_cast<void*>((uintptr_t)base)));
  memcpy(handle_out, &hd, sizeof(hd));
  *offset_out = (int64_t)((unsigned long long)(uintptr_t)dev_ptr - base);
  return MVIN_OK;
}

int mvin_ipc_open(const void* handle, int64_t offset, void** ptr_out) {
  if (!handle || !ptr_out || offset < 0) return fail(MVIN_ERR_INVALID, "bad argument");
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle, sizeof(hd));
  std::lock_guard<std::mutex> lock(g_ipc_mutex);
  for (auto& o : g_ipc_opened)
    if (memcmp(&o.handle, &hd, sizeof(hd)) == 0) { *ptr_out = static_cast<char*>(o.base) + offset; return MVIN_OK; }
  void* base = nullptr;
  CUDA_TRY(cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
  g_ipc_opened.push_back({hd, base});
  *ptr_out = static_cast<char*>(base) + offset;
  return MVIN_OK;
}

int mvin_set_batch_scale(mvin_handle_t h, int32_t global_batch, float dense_l2_scale) {
  if (!h || global_batch < 0) return fail(MVIN_ERR_INVALID, "bad argument");
  h->global_batch = global_batch;
  h->dense_l2_scale = dense_l2_scale;
  return MVIN_OK;
}

size_t mvin_workspace_bytes(mvin_handle_t h, int32_t B) {
  if (!h || B < 1) return 0;
  return make_layout(h->cfg, B, use_entity_leaf(h->cfg, B, h->n_shards, h->entity_leaf_mode)).total;
}

int mvin_get_neighbors(mvin_handle_t h, const int64_t* item_indices, int32_t B, int32_t n_levels,
                       int64_t* const* entities, int64_t* const* relations, void* stream) {
  if (!h || !item_indices || !entities || (n_levels > 0 && !relations) || B < 1 || n_levels < 0)
    return fail(MVIN_ERR_INVALID, "bad argument");
  if (!h->adj) return fail(MVIN_ERR_STATE, "adjacency not bound");
  cudaStream_t st = (cudaStream_t)stream;
  const int K = h->cfg.neighbor_sample_size;
  MVIN_LAUNCH((copy_i64_kernel), (B + 255) / 256, 256, 0, st, item_indices, B, entities[0]);
  LAUNCH_CHECK(h, "copy_i64");
  long rows = B;
  for (int i = 0; i < n_levels; ++i) {
    const long n = rows * K;
    MVIN_LAUNCH((expand_i64_kernel), (unsigned)((n + 255) / 256), 256, 0, st, entities[i], h->adj, rows, K, entities[i + 1],
                                                                   relations[i]);
    LAUNCH_CHECK(h, "expand_i64");
    rows = n;
  }
  return MVIN_OK;
}

int mvin_forward(mvin_handle_t h, const int64_t* user_indices, const int64_t* item_indices, const int32_t* mem_h,
                 const int32_t* mem_r, const int32_t* mem_t, int32_t B, float* scores, float* scores_normalized,
                 void* workspace, void* stream) {
  if (!h || !item_indices || !mem_h || !mem_r || !mem_t || !workspace) return fail(MVIN_ERR_INVALID, "null argument");
  if (B < 1 || B > h->cfg.max_batch) return fail(MVIN_ERR_INVALID, "B = %d outside 1..max_batch (%d)", B, h->cfg.max_batch);
  if (!h->has_params || !h->adj) return fail(MVIN_ERR_STATE, "parameters / adjacency not bound");
  h->user = user_indices;   // inert under --ablation all; kept for the feed contract
  h->item = item_indices;
  h->mem_h = mem_h; h->mem_r = mem_r; h->mem_t = mem_t;
  h->B = B;
  h->fwd_workspace = workspace;
  return dispatch_forward(h, item_indices, mem_h, mem_r, mem_t, B, scores, scores_normalized, workspace,
                          (cudaStream_t)stream);
}

int mvin_importance(mvin_handle_t h, float* imp0, float* imp1, void* workspace, void* stream) {
  if (!h || !imp0 || !workspace) return fail(MVIN_ERR_INVALID, "null argument");
  if (h->fwd_workspace != workspace || h->B < 1) return fail(MVIN_ERR_STATE, "no forward pass on this workspace");
  cudaStream_t st = (cudaStream_t)stream;
  const Layout L = make_layout(h->cfg, h->B, use_entity_leaf(h->cfg, h->B, h->n_shards, h->entity_leaf_mode));
  const int K = h->cfg.neighbor_sample_size;
  float* outs[2] = {imp0, imp1};
  for (int lv = 0; lv < 2 && lv < h->cfg.h_hop; ++lv) {
    if (!outs[lv]) continue;
    const long rows = L.rows[lv];
    MVIN_LAUNCH((importance_kernel), (unsigned)((rows * 32 + 255) / 256), 256, 0, st, at<int32_t>(workspace, L.ent[lv]), h->adj,
                                                                           at<float>(workspace, L.s), rows, K, outs[lv]);
    LAUNCH_CHECK(h, "importance");
  }
  return MVIN_OK;
}

int mvin_backward(mvin_handle_t h, const float* labels, int32_t B, float* losses_out, void* workspace, void* stream) {
  if (!h || !labels || !losses_out || !workspace) return fail(MVIN_ERR_INVALID, "null argument");
  if (!h->has_grads) return fail(MVIN_ERR_STATE, "gradient buffers not bound");
  if (h->fwd_workspace != workspace || h->B != B)
    return fail(MVIN_ERR_STATE, "mvin_backward must follow mvin_forward on the same workspace and batch size");
  return dispatch_backward(h, labels, B, losses_out, workspace, (cudaStream_t)stream);
}

int mvin_adam_step(mvin_handle_t h, const mvin_params_t* m, const mvin_params_t* v, float lr, float beta1, float beta2,
                   float eps, int32_t step, void* stream) {
  if (!h || !m || !v || step < 1) return fail(MVIN_ERR_INVALID, "bad argument");
  if (!h->has_params || !h->has_grads) return fail(MVIN_ERR_STATE, "parameters / gradients not bound");
  const mvin_config_t& c = h->cfg;
  const long D = c.dim, H = c.h_hop, p = c.p_hop, nr = c.n_relation;
  AdamSegments sg;
  memset(&sg, 0, sizeof(sg));
  int n = 0;
  auto add = [&](float* prm, const float* grd, float* mm, float* vv, long cnt) {
    sg.param[n] = prm; sg.grad[n] = grd; sg.m[n] = mm; sg.v[n] = vv; sg.n[n] = cnt;
    ++n;
  };
#define SEG(field, cnt) add(h->P.field, h->G.field, m->field, v->field, (cnt))
  SEG(user_emb, (long)c.n_user * D);
  SEG(entity_emb, h->n_local_rows * D);
  SEG(relation_emb, nr * D);
  SEG(relation_kge, nr * D * D);
  SEG(mix_w, (H + 1) * D * D);
  SEG(mix_b, D);
  SEG(user_mlp_w, (p + 1) * D * D);
  SEG(user_mlp_b, D);
  SEG(transfer_w, (H + 1) * D * D);
  SEG(transfer_b, (H + 1) * D);
  SEG(h_item_w, 2 * D);
  SEG(h_item_b, 1);
  SEG(agg_w, H * D * D);
  SEG(agg_b, H * D);
  SEG(agg_urh_w, H * 3 * D);
  SEG(agg_urh_b, H);
#undef SEG
  sg.count = n;
  const float lr_t = (float)((double)lr * std::sqrt(1.0 - std::pow((double)beta2, step)) /
                             (1.0 - std::pow((double)beta1, step)));
  cudaStream_t st = (cudaStream_t)stream;
  prof_mark(h, st, nullptr);
  MVIN_LAUNCH((adam_kernel), h->sm_count * 4, 256, 0, st, sg, lr_t, beta1, beta2, eps);
  LAUNCH_CHECK(h, "adam");
  return MVIN_OK;
}

size_t mvin_feed_bytes(mvin_handle_t h, int32_t B) {
  if (!h || B < 1) return 0;
  const size_t pm = (size_t)(h->cfg.p_hop > 0 ? h->cfg.p_hop : 1) * B * h->cfg.n_memory;
  return align_up(sizeof(int64_t) * B) * 2 + align_up(sizeof(float) * B) + 3 * align_up(sizeof(int32_t) * pm) +
         align_up(sizeof(float) * 4);
}

int mvin_train_step_host(mvin_handle_t h, const int64_t* user_indices, const int64_t* item_indices, const float* labels,
                         const int32_t* mem_h, const int32_t* mem_r, const int32_t* mem_t, int32_t B, void* staging,
                         void* workspace, const mvin_params_t* adam_m, const mvin_params_t* adam_v, float lr, int32_t step,
                         float* losses_host, void* stream) {
  if (!h || !user_indices || !item_indices || !labels || !mem_h || !mem_r || !mem_t || !staging || !workspace ||
      !losses_host)
    return fail(MVIN_ERR_INVALID, "null argument");
  if (B < 1 || B > h->cfg.max_batch) return fail(MVIN_ERR_INVALID, "B = %d outside 1..max_batch (%d)", B, h->cfg.max_batch);
  HostStepGuard guard(h);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t pm = (size_t)(h->cfg.p_hop > 0 ? h->cfg.p_hop : 1) * B * h->cfg.n_memory;
  char* base = static_cast<char*>(staging);
  size_t off = 0;
  auto stage = [&](const void* src, size_t bytes) -> void* {
    void* dst = base + off;
    off += align_up(bytes);
    cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
    return dst;
  };
  const int64_t* d_user = (const int64_t*)stage(user_indices, sizeof(int64_t) * B);
  const int64_t* d_item = (const int64_t*)stage(item_indices, sizeof(int64_t) * B);
  const float* d_lab = (const float*)stage(labels, sizeof(float) * B);
  int rc = host_step_overlap(h, B, workspace, st);
  if (rc) return rc;
  const int32_t* d_mh = (const int32_t*)stage(mem_h, sizeof(int32_t) * pm);
  const int32_t* d_mr = (const int32_t*)stage(mem_r, sizeof(int32_t) * pm);
  const int32_t* d_mt = (const int32_t*)stage(mem_t, sizeof(int32_t) * pm);
  float* d_loss = (float*)(base + off);
  CUDA_TRY(cudaGetLastError());
  rc = mvin_forward(h, d_user, d_item, d_mh, d_mr, d_mt, B, nullptr, nullptr, workspace, stream);
  if (rc) return rc;
  rc = mvin_backward(h, d_lab, B, d_loss, workspace, stream);
  if (rc) return rc;
  if (adam_m && adam_v) {
    rc = mvin_adam_step(h, adam_m, adam_v, lr, 0.9f, 0.999f, 1e-8f, step, stream);
    if (rc) return rc;
  }
  CUDA_TRY(cudaMemcpyAsync(losses_host, d_loss, sizeof(float) * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return MVIN_OK;
}

int mvin_feed_prefetch(mvin_handle_t h, const int64_t* user_indices, const int64_t* item_indices, const float* labels,
                       const int32_t* mem_h, const int32_t* mem_r, const int32_t* mem_t, int32_t B, void* staging,
                       int32_t slot, void* stream) {
  if (!h || !user_indices || !item_indices || !labels || !mem_h || !mem_r || !mem_t || !staging || slot < 0 || slot > 1)
    return fail(MVIN_ERR_INVALID, "null argument / slot must be 0 or 1");
  if (B < 1 || B > h->cfg.max_batch) return fail(MVIN_ERR_INVALID, "B = %d outside 1..max_batch (%d)", B, h->cfg.max_batch);
  if (!h->copy_stream) {
    CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_feed[0], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_feed[1], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_free, cudaEventDisableTiming));
  }
  // the copy may start once everything enqueued so far on the caller's stream -- in particular the last step that read
  // this staging buffer -- has finished; it then runs beside whatever the caller enqueues next
  CUDA_TRY(cudaEventRecord(h->ev_free, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_free, 0));
  const size_t pm = (size_t)(h->cfg.p_hop > 0 ? h->cfg.p_hop : 1) * B * h->cfg.n_memory;
  char* base = static_cast<char*>(staging);
  size_t off = 0;
  auto stage = [&](const void* src, size_t bytes) {
    cudaMemcpyAsync(base + off, src, bytes, cudaMemcpyHostToDevice, h->copy_stream);
    off += align_up(bytes);
  };
  stage(user_indices, sizeof(int64_t) * B);
  stage(item_indices, sizeof(int64_t) * B);
  stage(labels, sizeof(float) * B);
  stage(mem_h, sizeof(int32_t) * pm);
  stage(mem_r, sizeof(int32_t) * pm);
  stage(mem_t, sizeof(int32_t) * pm);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(h->ev_feed[slot], h->copy_stream));
  h->prefetched_staging[slot] = staging;
  h->prefetched_B[slot] = B;
  return MVIN_OK;
}

int mvin_train_step_prefetched(mvin_handle_t h, int32_t B, void* staging, int32_t slot, void* workspace,
                               const mvin_params_t* adam_m, const mvin_params_t* adam_v, float lr, int32_t step,
                               float* losses_host, void* stream) {
  if (!h || !staging || !workspace || !losses_host || slot < 0 || slot > 1) return fail(MVIN_ERR_INVALID, "bad argument");
  if (h->prefetched_staging[slot] != staging || h->prefetched_B[slot] != B)
    return fail(MVIN_ERR_STATE, "mvin_train_step_prefetched must follow mvin_feed_prefetch on the same slot, staging buffer and B");
  h->prefetched_staging[slot] = nullptr;
  HostStepGuard guard(h);
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaStreamWaitEvent(st, h->ev_feed[slot], 0));
  const size_t pm = (size_t)(h->cfg.p_hop > 0 ? h->cfg.p_hop : 1) * B * h->cfg.n_memory;
  char* base = static_cast<char*>(staging);
  size_t off = 0;
  auto carve = [&](size_t bytes) -> void* { void* p = base + off; off += align_up(bytes); return p; };
  const int64_t* d_user = (const int64_t*)carve(sizeof(int64_t) * B);
  const int64_t* d_item = (const int64_t*)carve(sizeof(int64_t) * B);
  const float* d_lab = (const float*)carve(sizeof(float) * B);
  const int32_t* d_mh = (const int32_t*)carve(sizeof(int32_t) * pm);
  const int32_t* d_mr = (const int32_t*)carve(sizeof(int32_t) * pm);
  const int32_t* d_mt = (const int32_t*)carve(sizeof(int32_t) * pm);
  float* d_loss = (float*)(base + off);
  int rc = host_step_overlap(h, B, workspace, st);
  if (rc) return rc;
  rc = mvin_forward(h, d_user, d_item, d_mh, d_mr, d_mt, B, nullptr, nullptr, workspace, stream);
  if (rc) return rc;
  rc = mvin_backward(h, d_lab, B, d_loss, workspace, stream);
  if (rc) return rc;
  if (adam_m && adam_v) {
    rc = mvin_adam_step(h, adam_m, adam_v, lr, 0.9f, 0.999f, 1e-8f, step, stream);
    if (rc) return rc;
  }
  CUDA_TRY(cudaMemcpyAsync(losses_host, d_loss, sizeof(float) * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return MVIN_OK;
}

int mvin_test_umma_gemm(const float* A, const float* W, float* C, int64_t M, int32_t D, void* stream) {
  if (!A || !W || !C || M < 1) return fail(MVIN_ERR_INVALID, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((M + 127) / 128);
  int rc;
  if (D == 32) {
    if ((rc = set_smem(umma_gemm_test_kernel<32>, umma_gemm_test_smem<32>()))) return rc;
    MVIN_LAUNCH((umma_gemm_test_kernel<32>), grid, 256, umma_gemm_test_smem<32>(), st, A, W, C, M);
  } else if (D == 64) {
    if ((rc = set_smem(umma_gemm_test_kernel<64>, umma_gemm_test_smem<64>()))) return rc;
    MVIN_LAUNCH((umma_gemm_test_kernel<64>), grid, 256, umma_gemm_test_smem<64>(), st, A, W, C, M);
  } else {
    return fail(MVIN_ERR_UNSUPPORTED, "tcgen05 path: dim must be 32 or 64, got %d", D);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MVIN_ERR_CUDA, "launch umma_gemm_test: %s", cudaGetErrorString(e));
  return MVIN_OK;
}

int mvin_test_umma_dw(const float* A, const float* G, float* dump, int64_t M, int32_t D, int32_t variant, void* stream) {
  if (!A || !G || !dump || M < 1) return fail(MVIN_ERR_INVALID, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (D == 32) {
    if ((rc = set_smem(umma_dw_test_kernel<32>, umma_dw_test_smem<32>()))) return rc;
    MVIN_LAUNCH((umma_dw_test_kernel<32>), 1, 256, umma_dw_test_smem<32>(), st, A, G, dump, M, variant);
  } else if (D == 64) {
    if ((rc = set_smem(umma_dw_test_kernel<64>, umma_dw_test_smem<64>()))) return rc;
    MVIN_LAUNCH((umma_dw_test_kernel<64>), 1, 256, umma_dw_test_smem<64>(), st, A, G, dump, M, variant);
  } else {
    return fail(MVIN_ERR_UNSUPPORTED, "tcgen05 path: dim must be 32 or 64, got %d", D);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MVIN_ERR_CUDA, "launch umma_dw_test: %s", cudaGetErrorString(e));
  return MVIN_OK;
}

int mvin_bind_user_triplets(mvin_handle_t h, const int32_t* user_triplet_set) {
  if (!h || !user_triplet_set) return fail(MVIN_ERR_INVALID, "null argument");
  h->uts = user_triplet_set;
  return MVIN_OK;
}

int mvin_gather_feed(mvin_handle_t h, const int64_t* user_indices, int32_t B, int32_t* mem_h, int32_t* mem_r,
                     int32_t* mem_t, void* stream) {
  if (!h || !user_indices || !mem_h || !mem_r || !mem_t || B < 1) return fail(MVIN_ERR_INVALID, "bad argument");
  if (!h->uts) return fail(MVIN_ERR_STATE, "user triplet sets not bound (mvin_bind_user_triplets)");
  cudaStream_t st = (cudaStream_t)stream;
  const int P = h->cfg.p_hop > 0 ? h->cfg.p_hop : 1, m = h->cfg.n_memory;
  const long n = (long)P * B * m;
  MVIN_LAUNCH((gather_feed_kernel), (unsigned)((n + 255) / 256), 256, 0, st, h->uts, user_indices, B, P, m, mem_h, mem_r, mem_t);
  LAUNCH_CHECK(h, "gather_feed");
  return MVIN_OK;
}

int mvin_train_step_users_host(mvin_handle_t h, const int64_t* user_indices, const int64_t* item_indices,
                               const float* labels, int32_t B, void* staging, void* workspace,
                               const mvin_params_t* adam_m, const mvin_params_t* adam_v, float lr, int32_t step,
                               float* losses_host, void* stream) {
  if (!h || !user_indices || !item_indices || !labels || !staging || !workspace || !losses_host)
    return fail(MVIN_ERR_INVALID, "null argument");
  if (B < 1 || B > h->cfg.max_batch) return fail(MVIN_ERR_INVALID, "B = %d outside 1..max_batch (%d)", B, h->cfg.max_batch);
  if (!h->uts) return fail(MVIN_ERR_STATE, "user triplet sets not bound (mvin_bind_user_triplets)");
  HostStepGuard guard(h);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t pm = (size_t)(h->cfg.p_hop > 0 ? h->cfg.p_hop : 1) * B * h->cfg.n_memory;
  char* base = static_cast<char*>(staging);
  size_t off = 0;
  auto carve = [&](size_t bytes) -> void* { void* p = base + off; off += align_up(bytes); return p; };
  int64_t* d_user = (int64_t*)carve(sizeof(int64_t) * B);
  int64_t* d_item = (int64_t*)carve(sizeof(int64_t) * B);
  float* d_lab = (float*)carve(sizeof(float) * B);
  int32_t* d_mh = (int32_t*)carve(sizeof(int32_t) * pm);
  int32_t* d_mr = (int32_t*)carve(sizeof(int32_t) * pm);
  int32_t* d_mt = (int32_t*)carve(sizeof(int32_t) * pm);
  float* d_loss = (float*)(base + off);
  CUDA_TRY(cudaMemcpyAsync(d_user, user_indices, sizeof(int64_t) * B, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_item, item_indices, sizeof(int64_t) * B, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_lab, labels, sizeof(float) * B, cudaMemcpyHostToDevice, st));
  int rc = host_step_overlap(h, B, workspace, st);
  if (rc) return rc;
  rc = mvin_gather_feed(h, d_user, B, d_mh, d_mr, d_mt, stream);
  if (rc) return rc;
  rc = mvin_forward(h, d_user, d_item, d_mh, d_mr, d_mt, B, nullptr, nullptr, workspace, stream);
  if (rc) return rc;
  rc = mvin_backward(h, d_lab, B, d_loss, workspace, stream);
  if (rc) return rc;
  if (adam_m && adam_v) {
    rc = mvin_adam_step(h, adam_m, adam_v, lr, 0.9f, 0.999f, 1e-8f, step, stream);
    if (rc) return rc;
  }
  CUDA_TRY(cudaMemcpyAsync(losses_host, d_loss, sizeof(float) * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return MVIN_OK;
}

int mvin_sample_adjacency(const int64_t* indptr, const int32_t* nbr, const int32_t* rel, int32_t n_entity, int32_t K,
                          uint64_t seed, int32_t* adj_packed, int64_t* adj_entity, int64_t* adj_relation,
                          int64_t* picked_edges, void* stream) {
  if (!indptr || !nbr || !rel || n_entity < 1 || K < 1 || K > MAX_K || (!adj_packed && !adj_entity))
    return fail(MVIN_ERR_INVALID, "bad argument (K must be in 1..%d)", MAX_K);
  MVIN_LAUNCH((sample_adjacency_kernel), (unsigned)((n_entity + 127) / 128), 128, 0, (cudaStream_t)stream, 
      indptr, nbr, rel, n_entity, K, (unsigned long long)seed, adj_packed, adj_entity, adj_relation, picked_edges);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MVIN_ERR_CUDA, "launch sample_adjacency: %s", cudaGetErrorString(e));
  return MVIN_OK;
}

int mvin_build_ripple_sets(const int64_t* indptr, const int32_t* nbr, const int32_t* rel, const int64_t* hist_ptr,
                           const int32_t* hist_items, int32_t n_user, int32_t p_hop, int32_t n_memory, int32_t n_neighbor,
                           uint64_t seed, int32_t* user_triplet_set, int64_t* slots, void* stream) {
  if (!indptr || !nbr || !rel || !hist_ptr || !hist_items || !user_triplet_set || n_user < 1)
    return fail(MVIN_ERR_INVALID, "null / bad argument");
  if (n_memory < 1 || n_memory > 64 || n_neighbor < 1 || n_neighbor > 16)
    return fail(MVIN_ERR_UNSUPPORTED, "device ripple sets: n_memory must be in 1..64 and n_neighbor in 1..16");
  const int P = p_hop > 0 ? p_hop : 1;
  MVIN_LAUNCH((ripple_sets_kernel), (unsigned)((n_user + 63) / 64), 64, 0, (cudaStream_t)stream, 
      indptr, nbr, rel, hist_ptr, hist_items, n_user, P, n_memory, n_neighbor, (unsigned long long)seed, user_triplet_set, slots);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MVIN_ERR_CUDA, "launch ripple_sets: %s", cudaGetErrorString(e));
  return MVIN_OK;
}

int mvin_topk_metrics(const float* scores, const uint8_t* relevant, const int32_t* n_cand, const int32_t* n_answers,
                      int32_t n_users, int32_t max_cand, const int32_t* k_list, int32_t nk, float* precision, float* recall,
                      float* ndcg, void* stream) {
  if (!scores || !relevant || !n_cand || !n_answers || !k_list || !precision || !recall || !ndcg || n_users < 1)
    return fail(MVIN_ERR_INVALID, "null / bad argument");
  if (nk < 1 || nk > TOPK_MAX_K || max_cand < 1) return fail(MVIN_ERR_INVALID, "nk must be in 1..%d", TOPK_MAX_K);
  TopkArgs a;
  memset(&a, 0, sizeof(a));
  a.scores = scores; a.rel = relevant; a.n_cand = n_cand; a.n_answers = n_answers; a.max_cand = max_cand; a.nk = nk;
  for (int i = 0; i < nk; ++i) {
    if (k_list[i] < 1 || k_list[i] > TOPK_MAX_RANK || (i && k_list[i] < k_list[i - 1]))
      return fail(MVIN_ERR_INVALID, "k_list must be ascending with entries in 1..%d", TOPK_MAX_RANK);
    a.k_list[i] = k_list[i];
  }
  a.precision = precision; a.recall = recall; a.ndcg = ndcg;
  const size_t sm = sizeof(float) * max_cand + (size_t)k_list[nk - 1] + 16;
  if (sm > 200 * 1024) return fail(MVIN_ERR_UNSUPPORTED, "too many candidates per user (%d)", max_cand);
  int rc;
  if ((rc = set_smem(topk_metrics_kernel, sm))) return rc;
  MVIN_LAUNCH((topk_metrics_kernel), n_users, 256, sm, (cudaStream_t)stream, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MVIN_ERR_CUDA, "launch topk_metrics: %s", cudaGetErrorString(e));
  return MVIN_OK;
}

int mvin_ctr_metrics(mvin_handle_t h, const float* scores_normalized, const float* labels, int32_t B, float* out3,
                     void* scratch40, void* stream) {
  if (!h || !scores_normalized || !labels || !out3 || !scratch40 || B < 1) return fail(MVIN_ERR_INVALID, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* acc64 = static_cast<unsigned long long*>(scratch40);
  CUDA_TRY(cudaMemsetAsync(acc64, 0, 5 * sizeof(unsigned long long), st));
  const int nt = 256;
  const int gx = (B + nt - 1) / nt;
  int gy = (4 * h->sm_count + gx - 1) / gx;            // split the j range so that ~4 CTAs per SM exist
  if (gy > gx) gy = gx;
  if (gy < 1) gy = 1;
  MVIN_LAUNCH((ctr_count_kernel), dim3(gx, gy), nt, 2 * nt * sizeof(float), st, scores_normalized, labels, B, acc64);
  LAUNCH_CHECK(h, "ctr_count");
  MVIN_LAUNCH((ctr_finalize_kernel), 1, 32, 0, st, acc64, B, out3);
  LAUNCH_CHECK(h, "ctr_finalize");
  return MVIN_OK;
}

int64_t mvin_launch_count(mvin_handle_t h) { return h ? h->launches : 0; }

int mvin_profile_enable(mvin_handle_t h, int32_t on) {
  if (!h) return fail(MVIN_ERR_INVALID, "null handle");
  h->prof_on = on != 0;
  return MVIN_OK;
}

int mvin_profile_read(mvin_handle_t h, char* buf, size_t buflen) {
  if (!h || !buf || buflen < 2) return fail(MVIN_ERR_INVALID, "bad argument");
  struct Row { const char* name; double ms; long n; };
  std::vector<Row> rows;
  if (!h->prof.empty()) CUDA_TRY(cudaEventSynchronize(h->prof.back().ev));
  for (size_t i = 1; i < h->prof.size(); ++i) {
    const char* name = h->prof[i].name;
    if (!name) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->prof[i - 1].ev, h->prof[i].ev) != cudaSuccess) continue;
    Row* r = nullptr;
    for (auto& x : rows)
      if (strcmp(x.name, name) == 0) r = &x;
    if (!r) { rows.push_back({name, 0.0, 0}); r = &rows.back(); }
    r->ms += ms;
    r->n += 1;
  }
  for (auto& e : h->prof) cudaEventDestroy(e.ev);
  h->prof.clear();
  size_t off = 0;
  buf[0] = 0;
  for (auto& r : rows) {
    int w = snprintf(buf + off, buflen - off, "%s:%.6f:%ld;", r.name, r.ms, r.n);
    if (w < 0 || (size_t)w >= buflen - off) break;
    off += (size_t)w;
  }
  return MVIN_OK;
}

}  // extern "C"
