// mvin_capi.cu -- libmvin_b200.so: the C ABI (include/mvin_b200.h).  The per-step orchestration lives in steps.cuh
// (one translation unit per embedding dimension), the shared host state in host.cuh.
#include "host.cuh"

namespace mvin_host {
char* err_buf() {
  thread_local char buf[512] = "";
  return buf;
}
// instantiated in mvin_steps.cu, one object file per dimension
template <int D>
int forward_impl(mvin_handle_t h, const int64_t* item, const int32_t* mem_h, const int32_t* mem_r, const int32_t* mem_t,
                 int B, float* scores, float* scores_norm, void* ws, cudaStream_t st);
template <int D>
int backward_init(mvin_handle_t h, int B, void* ws, cudaStream_t st, cudaEvent_t mid, bool zero_small);
template <int D>
int backward_impl(mvin_handle_t h, const float* labels, int B, float* losses_out, void* ws, cudaStream_t st);
template <int D>
int xchg_expand_impl(mvin_handle_t h, const int64_t* item, int B, int32_t* ids_out, void* ws, cudaStream_t st);
template <int D>
int xchg_owner_impl(mvin_handle_t h, int owner, bool bwd, int B, void* ws, cudaStream_t st);
template <int D>
int xchg_finish_impl(mvin_handle_t h, int B, void* ws, cudaStream_t st);
#define MVIN_EXTERN_D(D)                                                                                              \
  extern template int forward_impl<D>(mvin_handle_t, const int64_t*, const int32_t*, const int32_t*, const int32_t*,  \
                                      int, float*, float*, void*, cudaStream_t);                                      \
  extern template int backward_init<D>(mvin_handle_t, int, void*, cudaStream_t, cudaEvent_t, bool);                   \
  extern template int backward_impl<D>(mvin_handle_t, const float*, int, float*, void*, cudaStream_t);         \
  extern template int xchg_expand_impl<D>(mvin_handle_t, const int64_t*, int, int32_t*, void*, cudaStream_t);         \
  extern template int xchg_owner_impl<D>(mvin_handle_t, int, bool, int, void*, cudaStream_t);                         \
  extern template int xchg_finish_impl<D>(mvin_handle_t, int, void*, cudaStream_t);
MVIN_EXTERN_D(8) MVIN_EXTERN_D(16) MVIN_EXTERN_D(32) MVIN_EXTERN_D(64) MVIN_EXTERN_D(128)
#undef MVIN_EXTERN_D

int check_supported(const mvin_config_t* c) {
  if (!(c->flags & MVIN_FLAG_WIDE_DEEP) || (c->flags & ~0x7f) ||
      ((c->flags & MVIN_FLAG_PS_ONLY) && (c->flags & MVIN_FLAG_HO_ONLY)))
    return fail(MVIN_ERR_UNSUPPORTED,
                "supported: the settings of parameter_ablation.py with wide_deep = 1 (flags bit 4) and at most one of "
                "PS_only / HO_only; got flags 0x%x", c->flags);
  if (!(c->flags & MVIN_FLAG_PS_O_FT) && c->p_hop < 1)
    return fail(MVIN_ERR_UNSUPPORTED, "PS_O_ft = 0 needs p_hop >= 1 (the user MLP would have no input, model.py:232-236)");
  if (c->h_hop < 1 || c->n_mix_hop < 1) return fail(MVIN_ERR_UNSUPPORTED, "h_hop and n_mix_hop must be >= 1");
  if (c->n_mix_hop == 1 && c->h_hop > 3) return fail(MVIN_ERR_UNSUPPORTED, "h_hop must be in 1..3, got %d", c->h_hop);
  if (c->n_mix_hop > 1 && c->h_hop * c->n_mix_hop > MAX_L)
    return fail(MVIN_ERR_UNSUPPORTED, "h_hop * n_mix_hop must be <= %d when n_mix_hop > 1, got %d x %d", MAX_L, c->h_hop,
                c->n_mix_hop);
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
  // the PS_only / n_mix_hop > 1 variants run their step on one stream (steps.cuh: backward_ps_only, backward_mix_impl)
  if ((h->cfg.flags & MVIN_FLAG_PS_ONLY) || generic_step(h->cfg)) return MVIN_OK;
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
int dispatch_xchg_expand(mvin_handle_t h, const int64_t* item, int B, int32_t* ids, void* ws, cudaStream_t st) {
  DISPATCH_D(h->cfg.dim, (xchg_expand_impl<DD>(h, item, B, ids, ws, st)));
}
int dispatch_xchg_owner(mvin_handle_t h, int owner, bool bwd, int B, void* ws, cudaStream_t st) {
  DISPATCH_D(h->cfg.dim, (xchg_owner_impl<DD>(h, owner, bwd, B, ws, st)));
}
int dispatch_xchg_finish(mvin_handle_t h, int B, void* ws, cudaStream_t st) {
  DISPATCH_D(h->cfg.dim, (xchg_finish_impl<DD>(h, B, ws, st)));
}

}  // namespace mvin_host

// ------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------
extern "C" {

int mvin_abi_version(void) { return MVIN_ABI_VERSION; }
const char* mvin_last_error(void) { return err_buf(); }

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
      const cudaError_t e = cudaGetLastError();
      mvin_destroy(h);
      return fail(MVIN_ERR_CUDA, "stream / event creation: %s", cudaGetErrorString(e));
    }
  }
  if (const char* ev = getenv("MVIN_B200_ENTITY_LEAF")) h->entity_leaf_mode = atoi(ev) != 0 ? 1 : 0;
  if (const char* ev = getenv("MVIN_B200_USER_PB_FWD")) { const int n = atoi(ev); if (n == 1 || n == 2 || n == 4) h->user_pb_fwd = n; }
  if (const char* ev = getenv("MVIN_B200_TC")) { const int n = atoi(ev); if (n >= 0 && n <= 2) h->tc_mode = n; }
  if (const char* ev = getenv("MVIN_B200_GROUP_SPLIT")) h->group_split = atoi(ev) == 2 ? 2 : 1;
  if (const char* ev = getenv("MVIN_B200_TCGEMM")) { const int n = atoi(ev); if (n >= 0 && n <= 2) h->tcg_mode = n; }
  if (const char* ev = getenv("MVIN_B200_GROUP")) { const int n = atoi(ev); if (n >= 0 && n <= 2) h->group_mode = n; }
  if (const char* ev = getenv("MVIN_B200_RING")) { const int n = atoi(ev); if (n >= 0 && n <= 2) h->ring_mode = n; }
  if (const char* ev = getenv("MVIN_B200_TABLE")) h->table_mode = atoi(ev) != 0 ? 1 : 0;
  if (const char* ev = getenv("MVIN_B200_TCBWD")) { const int n = atoi(ev); if (n >= 0 && n <= 2) h->tcb_mode = n; }
  if (const char* ev = getenv("MVIN_B200_STREAM")) h->stream_mode = atoi(ev) != 0 ? 1 : 0;
  if (const char* ev = getenv("MVIN_B200_CTAS_PER_SM")) { const int n = atoi(ev); if (n >= 1 && n <= 32) h->max_ctas_per_sm = n; }
  if (cudaMalloc(&h->d_shard_tab, sizeof(void*) * 2 * MAX_SHARDS) != cudaSuccess ||
      cudaMalloc(&h->d_sched, sizeof(int) * 16) != cudaSuccess ||
      cudaMemset(h->d_sched, 0, sizeof(int) * 16) != cudaSuccess) {
    const cudaError_t e = cudaGetLastError();
    mvin_destroy(h);
    return fail(MVIN_ERR_CUDA, "cudaMalloc(shard table / scheduler counters): %s", cudaGetErrorString(e));
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
  if (n_shards > 1 && ((h->cfg.flags != MVIN_FLAGS_ALL && h->cfg.flags != MVIN_FLAGS_NO_KG_EH_UO) || h->cfg.n_mix_hop > 1))
    return fail(MVIN_ERR_UNSUPPORTED, "row-sharded entity table: only --ablation all / no_kg_eh_uo with one mix block");
  int shift = 0;
  while ((1 << shift) < n_shards) ++shift;
  void* host[2 * MAX_SHARDS] = {nullptr};
  for (int i = 0; i < n_shards; ++i) {
    if (!entity_shards[i] || !grad_shards[i]) return fail(MVIN_ERR_INVALID, "null shard pointer %d", i);
    host[i] = const_cast<float*>(entity_shards[i]);
    host[MAX_SHARDS + i] = grad_shards[i];
  }
  CUDA_TRY(cudaMemcpy(h->d_shard_tab, host, sizeof(host), cudaMemcpyHostToDevice));
  memcpy(h->shard_host, host, sizeof(host));
  h->xchg.on = false;
  h->n_shards = n_shards;
  h->n_local_rows = ((long)h->cfg.n_entity + n_shards - 1) / n_shards;
  h->etab = ETab{nullptr, reinterpret_cast<const float* const*>(h->d_shard_tab), shift, n_shards - 1};
  h->gtab = GTab{nullptr, reinterpret_cast<float* const*>(h->d_shard_tab + MAX_SHARDS), shift, n_shards - 1};
  return MVIN_OK;
}

// ---- owner-side partial reduction of the leaf level (exchange.cuh) ------------------------------------------
int mvin_xchg_bind(mvin_handle_t h, int32_t n_src, int32_t src_index, int64_t rows, const int32_t* ids_all,
                   float* const* part, float* const* gsu, float* const* dot) {
  if (!h) return fail(MVIN_ERR_INVALID, "null argument");
  if (n_src == 0) { h->xchg.on = false; return MVIN_OK; }
  if (!ids_all || !part || !gsu || !dot) return fail(MVIN_ERR_INVALID, "null argument");
  if (h->n_shards < 2) return fail(MVIN_ERR_STATE, "mvin_xchg_bind needs a row-sharded entity table (mvin_bind_entity_shards)");
  if (n_src < 1 || n_src > XCHG_MAX_RANKS || src_index < 0 || src_index >= n_src || rows < 1)
    return fail(MVIN_ERR_INVALID, "bad exchange geometry (n_src %d, src_index %d, rows %ld)", n_src, src_index, (long)rows);
  for (int s = 0; s < n_src; ++s) {
    if (!part[s] || !gsu[s] || !dot[s]) return fail(MVIN_ERR_INVALID, "null exchange buffer of source %d", s);
    h->xchg.part[s] = part[s]; h->xchg.gsu[s] = gsu[s]; h->xchg.dot[s] = dot[s];
  }
  h->xchg.n_src = n_src; h->xchg.src_index = src_index; h->xchg.rows = rows; h->xchg.ids_all = ids_all;
  h->xchg.on = true;
  return MVIN_OK;
}

int mvin_xchg_expand(mvin_handle_t h, const int64_t* item_indices, int32_t B, int32_t* ids_out, void* workspace,
                     void* stream) {
  if (!h || !item_indices || !ids_out || !workspace) return fail(MVIN_ERR_INVALID, "null argument");
  if (!h->xchg.on) return fail(MVIN_ERR_STATE, "exchange buffers not bound (mvin_xchg_bind)");
  if (B < 1 || B > h->cfg.max_batch) return fail(MVIN_ERR_INVALID, "B = %d outside 1..max_batch (%d)", B, h->cfg.max_batch);
  if (!h->has_params || !h->adj) return fail(MVIN_ERR_STATE, "parameters / adjacency not bound");
  h->B = B;
  return dispatch_xchg_expand(h, item_indices, B, ids_out, workspace, (cudaStream_t)stream);
}

int mvin_xchg_owner_forward(mvin_handle_t h, int32_t owner, void* workspace, void* stream) {
  if (!h || !workspace) return fail(MVIN_ERR_INVALID, "null argument");
  if (!h->xchg.on || h->B < 1) return fail(MVIN_ERR_STATE, "mvin_xchg_owner_forward must follow mvin_xchg_expand");
  if (owner < 0 || owner >= h->n_shards) return fail(MVIN_ERR_INVALID, "owner %d outside 0..%d", owner, h->n_shards - 1);
  return dispatch_xchg_owner(h, owner, false, h->B, workspace, (cudaStream_t)stream);
}

int mvin_xchg_owner_backward(mvin_handle_t h, int32_t owner, void* workspace, void* stream) {
  if (!h || !workspace) return fail(MVIN_ERR_INVALID, "null argument");
  if (!h->xchg.on || h->fwd_workspace != workspace) return fail(MVIN_ERR_STATE, "mvin_xchg_owner_backward must follow mvin_backward");
  if (owner < 0 || owner >= h->n_shards) return fail(MVIN_ERR_INVALID, "owner %d outside 0..%d", owner, h->n_shards - 1);
  return dispatch_xchg_owner(h, owner, true, h->B, workspace, (cudaStream_t)stream);
}

int mvin_xchg_finish_backward(mvin_handle_t h, void* workspace, void* stream) {
  if (!h || !workspace) return fail(MVIN_ERR_INVALID, "null argument");
  if (!h->xchg.on || h->fwd_workspace != workspace) return fail(MVIN_ERR_STATE, "mvin_xchg_finish_backward must follow mvin_backward");
  return dispatch_xchg_finish(h, h->B, workspace, (cudaStream_t)stream);
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
  return handle_layout(h, B).total;
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
  if (h->cfg.flags & MVIN_FLAG_PS_ONLY) return fail(MVIN_ERR_UNSUPPORTED, "PS_only has no aggregators (model.py:142-144)");
  cudaStream_t st = (cudaStream_t)stream;
  const Layout L = handle_layout(h, h->B);
  const int K = h->cfg.neighbor_sample_size;
  // model.py:294,304: importance_list is reset at iteration 0 of EVERY mix block, so what survives is the attention of the
  // first aggregator of the LAST block (aggregator (M - 1) h_hop) at hops 0 and 1
  const float* s = at<float>(workspace, L.s) + (long)(h->cfg.n_mix_hop - 1) * h->cfg.h_hop * h->cfg.n_relation;
  float* outs[2] = {imp0, imp1};
  for (int lv = 0; lv < 2 && lv < h->cfg.h_hop; ++lv) {
    if (!outs[lv]) continue;
    const long rows = L.rows[lv];
    MVIN_LAUNCH((importance_kernel), (unsigned)((rows * 32 + 255) / 256), 256, 0, st, at<int32_t>(workspace, L.ent[lv]), h->adj,
                                                                           s, rows, K, outs[lv]);
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
  // Hm iterations per mix block, M mix blocks, H = Hm M aggregators / transfer matrices 0 .. H
  const long D = c.dim, Hm = c.h_hop, M = c.n_mix_hop, H = Hm * M, p = c.p_hop, nr = c.n_relation;
  AdamSegments sg;
  memset(&sg, 0, sizeof(sg));
  int n = 0;
  auto add = [&](float* prm, const float* grd, float* mm, float* vv, long cnt) {
    if (cnt <= 0) return;
    sg.param[n] = prm; sg.grad[n] = grd; sg.m[n] = mm; sg.v[n] = vv; sg.n[n] = cnt;
    sg.vec_end[n] = (n ? sg.vec_end[n - 1] : 0) + (cnt + 3) / 4;
    ++n;
  };
#define SEG(field, cnt) add(h->P.field, h->G.field, m->field, v->field, (cnt))
  SEG(user_emb, (long)c.n_user * D);
  SEG(entity_emb, h->n_local_rows * D);
  SEG(relation_emb, nr * D);
  SEG(relation_kge, nr * D * D);
  SEG(mix_w, M * (Hm + 1) * D * D);
  SEG(mix_b, M * D);
  SEG(user_mlp_w, (p + ((c.flags & MVIN_FLAG_PS_O_FT) ? 1 : 0)) * D * D);
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
  {
    const long want = (sg.vec_end[n - 1] + 255) / 256, cap = (long)h->sm_count * 8;
    MVIN_LAUNCH((adam_kernel), (unsigned)(want < cap ? want : cap), 256, 0, st, sg, lr_t, beta1, beta2, eps);
  }
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

int mvin_test_umma_bf16(const float* A, const float* G, const float* W, float* C, float* dump, int64_t M, int32_t D,
                        int32_t n_planes, void* stream) {
  if (!A || !G || !W || !C || !dump || M < 1) return fail(MVIN_ERR_INVALID, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
#define MVIN_BF_TEST(DD, NP)                                                                                  \
  do {                                                                                                        \
    if ((rc = set_smem(umma_bf_test_kernel<DD, NP>, umma_bf_test_smem<DD, NP>()))) return rc;                 \
    MVIN_LAUNCH((umma_bf_test_kernel<DD, NP>), 1, 256, umma_bf_test_smem<DD, NP>(), st, A, G, W, C, dump, M); \
  } while (0)
  if (D == 64 && n_planes == 2) MVIN_BF_TEST(64, 2);
  else if (D == 64 && n_planes == 3) MVIN_BF_TEST(64, 3);
  else if (D == 32 && n_planes == 2) MVIN_BF_TEST(32, 2);
  else if (D == 32 && n_planes == 3) MVIN_BF_TEST(32, 3);
  else return fail(MVIN_ERR_UNSUPPORTED, "tcgen05 bf16 path: dim must be 32 or 64 and n_planes 2 or 3");
#undef MVIN_BF_TEST
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MVIN_ERR_CUDA, "launch umma_bf_test: %s", cudaGetErrorString(e));
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
