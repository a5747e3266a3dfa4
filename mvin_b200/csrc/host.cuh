// host.cuh -- host-side state and helpers shared by the translation units of libmvin_b200.so: the handle, the
// workspace layout, launch / error macros, grid sizing.  mvin_capi.cu holds the C ABI; steps.cuh (compiled once per
// embedding dimension, mvin_steps.cu) holds the forward / backward orchestration.
#pragma once
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
#include "level_tcb.cuh"
#include "misc.cuh"
#include "table.cuh"
#include "exchange.cuh"
#include "group.cuh"
#include "umma.cuh"
#include "umma_bf.cuh"
#include "gemm_tc.cuh"
#include "user.cuh"

using namespace mvin;

namespace mvin_host {

char* err_buf();          // thread-local message buffer of mvin_last_error (defined in mvin_capi.cu), 512 bytes

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

inline void prof_mark(mvin_handle_t h, cudaStream_t st, const char* name);

// Every kernel goes through here.  With programmatic stream serialisation the kernel may start while its predecessor
// in the stream is still running; each kernel's first statement is pdl_enter() (common.cuh), which restores the
// dependency on the device.  MVIN_B200_PDL=0 launches normally.
inline bool pdl_enabled() {
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

constexpr int MAX_L = 4;            // deepest level count: h_hop <= 3 with one mix block, h_hop n_mix_hop <= 4 with several
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
  size_t RKs;                             // relation matrices pre-staged for the tcgen05 GEMMs (gemm_tc.cuh), 0 = none
  size_t ukg, user32, dukg;               // User_orient_kg_eh = 0: U[user] rows, int32 user ids, their gradient (zeroed region)
  size_t stamp, Se;                       // entity mode of the leaf level (stamp is cleared by every forward)
  bool entity_leaf;
  // table mode (table.cuh): composed maps, per-entity tables A_h, per-pair vectors C_h and their gradients
  bool table;
  size_t Mc, cst, Atab, Cp;               // [H][TBL_NM][D][D], [H][D], [H][n_entity][D], [H][B][D]
  size_t dA, dCs, dM, dcst;               // [H][n_entity][D], [H][B][D], [H][3][D][D], [H][D]   (inside the zeroed region)
  // entity-group evaluation of the table-gather level (group.cuh)
  bool group;
  size_t gcnt, goff, gtot, gorder, gesort, GP;   // int32 [n_entity] x 2, [4096 + 1], [rows(H-2)] x 2; fp32 [rows(H-2), D]
  // n_mix_hop > 1 (steps.cuh, forward_mix_impl / backward_mix_impl; model.py:286-315): X[n][h] = output of mix block n-1 at
  // level h (the input of block n >= 1); DM[n][h] = [h_hop + 1][rows(h), D] gradients of the h_hop + 1 inputs of mix block n
  // at level h; GV / GX / GT = summed gradients of V[j][h], X[n][h], T[h] where more than one consumer contributes
  size_t X[MAX_L][MAX_L], GX[MAX_L][MAX_L], DM[MAX_L][MAX_L], GV[MAX_L + 1][MAX_L], GT[MAX_L];
  // User_orient = 0 (generic step): the transform kernels run with the identity, a zero bias and a zero user vector, their
  // parameter / user gradients land in scratch:  eye [D, D], zb [D], zu [B, D], sdW [D, D], sdb [D], sdu [B, D]
  size_t eye, zb, zu, sdW, sdb, sdu;
  // PS_O_ft = 0: user_mlp_matrix [p D, D] padded with a zero block for the absent user_h_set slot, and its gradient
  size_t Wpad, dWpad;
  size_t total;
  long rows[MAX_L + 1];
};

}  // namespace mvin_host

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
  int table_mode = -1;             // entity-table form of aggregator iteration 0 (table.cuh): -1 auto, 0 off, 1 on
                                   // (env MVIN_B200_TABLE)
  int group_split = 1;             // warps sharing a window of the backward group kernel: 1 or 2 (env MVIN_B200_GROUP_SPLIT)
  int tcg_mode = 1;                // tcgen05 relation-batched GEMMs of the user side (gemm_tc.cuh): 0 never, 1 from 512 pairs on,
                                   // 2 always (env MVIN_B200_TCGEMM)
  int group_mode = 1;              // table-gather level per entity group (group.cuh): 0 off, 1 from 65 536 rows on, 2 whenever
                                   // supported (env MVIN_B200_GROUP)
  int ring_mode = 1;               // table-gather levels stage their rows with cp.async.bulk (level.cuh, RowRing): 0 off,
                                   // 1 backward kernel, 2 forward kernel too (env MVIN_B200_RING)
  int tcb_mode = 1;                // tcgen05 backward kernels of the deepest level (level_tcb.cuh): 0 never, 1 auto, 2 always
                                   // (env MVIN_B200_TCBWD)
  int max_ctas_per_sm = 4;         // cap on resident CTAs per SM of the persistent row kernels (env MVIN_B200_CTAS_PER_SM)
  // owner-side partial reduction of the leaf level (exchange.cuh; mvin_xchg_*): peer-visible buffers of every source rank
  struct Xchg {
    bool on = false;
    int n_src = 0, src_index = 0;    // source ranks taking part; this rank's index among them
    long rows = 0;                   // leaf-level parent nodes per source rank
    const int32_t* ids_all = nullptr;
    float* part[XCHG_MAX_RANKS] = {nullptr};
    float* gsu[XCHG_MAX_RANKS] = {nullptr};
    float* dot[XCHG_MAX_RANKS] = {nullptr};
  } xchg;
  void* shard_host[2 * 16] = {nullptr};   // host copy of the shard pointer table (entity shards, then gradient shards)
  bool prof_on = false;
  struct ProfRec { const char* name; cudaEvent_t ev; };
  std::vector<ProfRec> prof;
};

namespace mvin_host {

// Kernel timing: when enabled, one CUDA event is recorded on the launch stream after every kernel launch (and one
// marker at each API entry); the stream is in-order, so consecutive events bracket one kernel.
inline void prof_mark(mvin_handle_t h, cudaStream_t st, const char* name) {
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

// configurations served by the generic per-level step (steps.cuh: forward_mix_impl / backward_mix_impl) instead of the
// tuned single-block one: several mix blocks, no user-oriented transform, no relation attention
inline bool generic_step(const mvin_config_t& c) {
  return c.n_mix_hop > 1 || !(c.flags & MVIN_FLAG_USER_ORIENT) || !(c.flags & MVIN_FLAG_USER_ORIENT_RELA);
}

inline bool has_agg(int H, int i, int h) { return i < H && h < H - i; }          // aggregator step (i, h) exists
inline bool has_V(int H, int j, int h) { return j == 0 ? h < H : h <= H - j; }   // buffer V[j][h] exists

// Entity mode of the leaf level (level.cuh, leaf_entity_kernel) pays off when the depth-(L-1) nodes of a batch
// re-use entities: enabled when there are at least n_entity / 4 of them and the two per-entity buffers are small.
inline bool use_entity_leaf(const mvin_config_t& c, long B, int n_shards, int mode) {
  if (n_shards != 1 || mode == 0 || generic_step(c) || (c.flags & MVIN_FLAG_PS_ONLY)) return false;
  if (mode == 1) return true;
  long rows = B;
  for (int h = 1; h < c.h_hop; ++h) rows *= c.neighbor_sample_size;
  return n_shards == 1 && rows * 4 >= (long)c.n_entity && (long)c.n_entity * c.dim * 8 <= (2L << 30);
}

// Table mode (table.cuh) replaces the per-row evaluation of aggregator iteration 0 -- and with it the deepest level's
// buffers -- by per-entity tables.  Same applicability as the entity mode of the leaf level (one shard, the batch re-uses
// entities); MVIN_B200_TABLE=0 / 1 forces it off / on, an explicit MVIN_B200_ENTITY_LEAF selects the row kernels.
inline bool use_table(const mvin_config_t& c, long B, int n_shards, int table_mode, int entity_leaf_mode) {
  if (n_shards != 1 || table_mode == 0 || generic_step(c) || (c.flags & MVIN_FLAG_PS_ONLY)) return false;
  if ((long)c.n_entity * c.dim * 4 * (2 * c.h_hop + 2) > (8L << 30)) return false;
  if (table_mode == 1) return true;
  // automatic: when the deepest level has at least one row per entity of the graph (rows per entity at C3: 2.5, C4: 148;
  // measured: C3 1.15 vs 1.42 ms, C4 5.4 vs 38.9 ms per step.  At C2, 0.36 rows per entity, the table kernels' walk over
  // all 182 011 entities costs more than the 65 536 leaf rows it replaces: 0.49 vs 0.35 ms)
  long rows = B;
  for (int h = 1; h < c.h_hop; ++h) rows *= c.neighbor_sample_size;
  return entity_leaf_mode == -1 && rows >= (long)c.n_entity;
}

// buffer V[j][h] (and its gradient) exists in table mode: T[0] for the mix layer, the iteration-0 outputs of every
// level but the deepest, everything of the later iterations
inline bool tab_has_V(int H, int j, int h) {
  if (j == 0) return h == 0;
  if (j == 1) return h == 0 || h <= H - 2;
  return h <= H - j;
}

inline Layout make_layout(const mvin_config_t& c, long B, bool entity_leaf, bool table = false, bool group = false) {
  Layout L;
  memset(&L, 0, sizeof(L));
  if (table) entity_leaf = true;           // Se / GSe / stamp are shared with the entity mode of the leaf level
  L.entity_leaf = entity_leaf;
  L.table = table;
  // H = depth of the neighbour expansion = number of aggregators: h_hop iterations in each of n_mix_hop mix blocks
  const long D = c.dim, K = c.neighbor_sample_size, H = (long)c.h_hop * c.n_mix_hop, p = c.p_hop, m = c.n_memory, nr = c.n_relation;
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
  if (D == 32 || D == 64) L.RKs = take((size_t)nr * 4 * (D / 8) * ((D / 4) * 144));   // nr x rel_stage_bytes<D>()
  L.ukg = take(f * B * D);
  L.user32 = take(sizeof(int32_t) * B);
  L.s = take(f * H * nr);
  if (!table) L.SU = take(f * L.rows[H - 1] * D);
  auto exists = [&](int j, int h) { return table ? (h < H && tab_has_V(H, j, h)) : has_V(H, j, h); };
  const size_t vtop = take(f * (H + 1) * B * D);
  for (int j = 0; j <= H; ++j)
    for (int h = 0; h < MAX_L; ++h)
      if (exists(j, h)) L.V[j][h] = h == 0 ? vtop + f * j * B * D : take(f * L.rows[h] * D);
  for (int i = table ? 1 : 0; i < H; ++i)
    for (int h = 0; h < H - i; ++h) L.Y[i][h] = take(f * L.rows[h] * D);
  L.item = take(f * B * D);
  L.scores = take(f * B);
  const size_t dtop = take(f * (H + 1) * B * D);
  for (int j = 0; j <= H; ++j)
    for (int h = 0; h < MAX_L; ++h)
      if (exists(j, h)) L.DC[j][h] = h == 0 ? dtop + f * j * B * D : take(f * L.rows[h] * D);
  for (int i = table ? 1 : 0; i < H; ++i)
    for (int h = 0; h < H - i; ++h) L.DS[i][h] = take(f * L.rows[h] * D);
  if (table) {
    L.Mc = take(f * H * TBL_NM * D * D);
    L.cst = take(f * H * D);
    L.Atab = take(f * H * (size_t)c.n_entity * D);
    L.Cp = take(f * H * B * D);
    L.group = group && H >= 2;
    if (L.group) {
      L.gcnt = take(sizeof(int32_t) * (size_t)c.n_entity);
      L.goff = take(sizeof(int32_t) * (size_t)c.n_entity);
      L.gtot = take(sizeof(int32_t) * (2 * SCAN_PER_BLOCK + 2));
      L.gorder = take(sizeof(int32_t) * L.rows[H - 2]);
      L.gesort = take(sizeof(int32_t) * L.rows[H - 2]);
      L.GP = take(f * L.rows[H - 2] * D);
    }
  }
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
  L.dukg = take(f * B * D);
  if (entity_leaf) L.GSe = take(f * (size_t)c.n_entity * D);
  if (table) {
    L.dA = take(f * H * (size_t)c.n_entity * D);
    L.dCs = take(f * H * B * D);
    L.dM = take(f * H * 3 * D * D);
    L.dcst = take(f * H * D);
  }
  L.zero_end = off;
  if (entity_leaf) {
    L.stamp = take(sizeof(int32_t) * (size_t)c.n_entity);
    L.Se = take(f * (size_t)c.n_entity * D);
  }
  if (!(c.flags & MVIN_FLAG_PS_O_FT)) {
    L.Wpad = take(f * (p + 1) * D * D);
    L.dWpad = take(f * (p + 1) * D * D);
  }
  if (generic_step(c) && !(c.flags & MVIN_FLAG_USER_ORIENT)) {
    L.eye = take(f * D * D); L.zb = take(f * D); L.zu = take(f * B * D);
    L.sdW = take(f * D * D); L.sdb = take(f * D); L.sdu = take(f * B * D);
  }
  if (generic_step(c)) {
    const int Hm = c.h_hop, M = c.n_mix_hop;
    for (int n = 0; n < M; ++n)
      for (int h = 0; h <= H - (long)n * Hm && h < MAX_L; ++h) {
        if (n >= 1) { L.X[n][h] = take(f * L.rows[h] * D); L.GX[n][h] = take(f * L.rows[h] * D); }
        if (h <= H - (long)(n + 1) * Hm) L.DM[n][h] = take(f * (Hm + 1) * L.rows[h] * D);
      }
    for (int j = 1; j <= H; ++j)
      for (int h = 0; h <= H - j; ++h) L.GV[j][h] = take(f * L.rows[h] * D);
    for (int h = 0; h < H; ++h) L.GT[h] = take(f * L.rows[h] * D);
  }
  L.total = off;
  return L;
}

inline Layout handle_layout(const mvin_handle_s* h, long B) {
  const bool table = use_table(h->cfg, B, h->n_shards, h->table_mode, h->entity_leaf_mode);
  // per-entity-group evaluation of the table-gather level: worth its sort and its two extra launches from 65 536 rows on
  // (C4: 524 288 rows, 5.5 -> 4.5 ms per step; C3: 8 192 rows, 1.16 -> 1.28 ms); MVIN_B200_GROUP=2 forces it
  long grows = B;
  for (int hh = 1; hh + 1 < h->cfg.h_hop; ++hh) grows *= h->cfg.neighbor_sample_size;
  const bool group = table && h->group_mode != 0 && (h->group_mode == 2 || grows >= 65536) &&
                     grp_supported(h->cfg.dim, h->cfg.neighbor_sample_size) &&
                     (long)h->cfg.n_entity <= (long)SCAN_PER_BLOCK * SCAN_PER_BLOCK;
  return make_layout(h->cfg, B, use_entity_leaf(h->cfg, B, h->n_shards, h->entity_leaf_mode), table, group);
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
inline int run_gemm(mvin_handle_t h, cudaStream_t st, const GemmArgs& g, const char* name = "gemm") {
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

inline GemmArgs gemm_args() {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.nbatch = 1;
  g.ksplit = 1;
  g.alpha = 1.f;
  return g;
}

inline int pick_ksplit(long K) {
  long s = K / 128;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return (int)s;
}

// cudaFuncSetAttribute is sticky per (kernel, device): remember what was already requested instead of calling the
// runtime on every launch of every step
template <typename KernelT>
int set_smem(KernelT k, size_t bytes) {
  struct Key { const void* f; size_t s; int dev; };
  static thread_local std::vector<Key> done;
  int dev = 0;
  cudaGetDevice(&dev);
  for (const Key& e : done)
    if (e.f == (const void*)k && e.dev == dev && e.s >= bytes) return MVIN_OK;
  done.push_back({(const void*)k, bytes, dev});
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
inline int partition_grid(const long* rows, int nlev, int tile_rows, int cap, int* cta_end) {
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

// tcgen05 backward of the deepest materialised level (level_tcb.cuh): per-entity leaf mode, d in {32, 64}, at least two
// levels, families of K = 2^j <= 64 rows (128-row tiles hold whole families); auto: from 131 072 rows at that level
inline bool use_tc_bwd(mvin_handle_t h, const Layout& L, int D) {
  const mvin_config_t& c = h->cfg;
  const int K = c.neighbor_sample_size, H = c.h_hop;
  if (h->tcb_mode == 0 || !L.entity_leaf || H < 2 || !(D == 32 || D == 64)) return false;
  if (K > 64 || (K & (K - 1)) != 0 || h->n_shards != 1) return false;
  return h->tcb_mode == 2 || L.rows[H - 1] >= 131072;
}

// leaf_entity_kernel: entities per warp visit, sized so that the launch has about 32 warps per SM
inline int leaf_chunk(mvin_handle_t h, long n_entity) {
  const long warps = (long)h->sm_count * 32;
  int chunk = 32;
  while (chunk > 1 && n_entity / chunk < warps) chunk >>= 1;
  return chunk;
}

// activation buffers of a level that dwarf L2 (126 MB) are accessed with streaming hints (common.cuh, ld4a / st4a)
inline int stream_level(mvin_handle_t h, long rows, int D) {
  if (h->stream_mode >= 0) return h->stream_mode;
  return (size_t)rows * D * sizeof(float) >= ((size_t)48 << 20) ? 1 : 0;
}

// Tile list of one aggregator launch (level.cuh, TileList): returns the grid size (every CTA is resident).
inline int make_tile_list(TileList& tl, const long* rows, int nlev, int tile_rows, int cap, int* ctr) {
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

}  // namespace mvin_host
using namespace mvin_host;
