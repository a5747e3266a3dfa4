/* mvin_b200.h -- C ABI of libmvin_b200.so: the MVIN hot path (multi-hop KG neighbour gather + attention
 * aggregation + RippleNet-style user o-set propagation, forward and backward) as sm_100a CUDA kernels.
 *
 * The reference (johnnyjana730/MVIN) has no FFI: its only boundary is the Python class `MVIN`
 * (src/model/MVIN/model.py:6-11) driven through `tf.Session.run`.  Every entry point below replaces the TF1
 * sub-graph named beside it; `mvin_b200/model.py` is the Python face that keeps the reference's
 * ctor / feed-dict / train / eval / get_scores API on top of this ABI.
 *
 * Conventions
 *   - plain C, no torch / CUDA-runtime types in signatures: streams are passed as `void*` (a cudaStream_t).
 *   - the CALLER owns every buffer (torch tensors in the Python face); the library allocates nothing on the
 *     device after mvin_create().  All device pointers must be 16-byte aligned, row-major, contiguous.
 *   - all work is enqueued on the caller's stream; no hidden synchronisation except in the *_host entry
 *     points, which are documented as blocking.
 *   - return value 0 = success, negative = error; message via mvin_last_error() (thread-local).
 *   - a handle is not thread-safe; use one per device / per stream.
 *   - supported model configuration: --ablation all (parameter_ablation.py:4-12), n_mix_hop = 1,
 *     1 <= h_hop <= 3, dim in {8,16,32,64,128}, neighbor_sample_size <= 64, p_hop >= 0.
 *     Anything else returns MVIN_ERR_UNSUPPORTED (the product never falls back to a CPU path).
 */
#ifndef MVIN_B200_H
#define MVIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVIN_OK 0
#define MVIN_ERR_INVALID -1
#define MVIN_ERR_UNSUPPORTED -2
#define MVIN_ERR_CUDA -3
#define MVIN_ERR_STATE -4

#define MVIN_ABI_VERSION 4

typedef struct mvin_handle_s* mvin_handle_t;

/* Mirrors the fields MVIN._parse_args reads (model.py:17-47) plus the table sizes of the ctor (model.py:7). */
typedef struct mvin_config {
  int32_t dim;                  /* --dim                    d */
  int32_t neighbor_sample_size; /* --neighbor_sample_size   K */
  int32_t h_hop;                /* --h_hop                  H (= L when n_mix_hop = 1) */
  int32_t n_mix_hop;            /* --n_mix_hop              M mix blocks of h_hop iterations each (model.py:286-315);
                                   h_hop <= 3 when M = 1, h_hop * M <= 4 otherwise */
  int32_t p_hop;                /* --p_hop                  p */
  int32_t n_memory;             /* --n_memory               m */
  int32_t n_user, n_entity, n_relation;
  int32_t max_batch;            /* largest B that will be passed (the reference bakes B into the graph) */
  float l2_weight;              /* --l2_weight     (model.py:412) */
  float l2_agg_weight;          /* --l2_agg_weight (model.py:412) */
  int32_t flags;                /* bit0 User_orient, bit1 User_orient_rela, bit2 User_orient_kg_eh, bit3 PS_O_ft,
                                   bit4 wide_deep, bit5 PS_only, bit6 HO_only; 'all' = 0x1f */
} mvin_config_t;

#define MVIN_FLAGS_ALL 0x1f
#define MVIN_FLAG_USER_ORIENT 0x01      /* bit0: User_orient (model.py:270-283); off: the entity vectors are used untransformed */
#define MVIN_FLAG_USER_ORIENT_RELA 0x02 /* bit1: User_orient_rela (aggregators.py:98-152); off: plain mean over the neighbours */
#define MVIN_FLAG_KG_EH 0x04            /* bit2: User_orient_kg_eh (model.py:152-156); off: the KG side is oriented by U[user] */
#define MVIN_FLAG_PS_O_FT 0x08          /* bit3: PS_O_ft (model.py:204-206); off: no user_h_set, user_mlp_matrix is [p d, d] */
#define MVIN_FLAG_WIDE_DEEP 0x10        /* bit4: wide_deep; must be set (the other branch, model.py:327-376, is broken upstream) */
#define MVIN_FLAG_PS_ONLY 0x20          /* bit5: PS_only (model.py:142-144): score = user_o . E[item], no KG side */
#define MVIN_FLAG_HO_ONLY 0x40          /* bit6: HO_only (model.py:146-150): score = U[user] . item */
#define MVIN_FLAGS_NO_KG_EH_UO 0x1b     /* --ablation no_kg_eh_uo (parameter_ablation.py:22-30) */
#define MVIN_FLAGS_PS_ONLY 0x3f          /* --ablation ps_only */
#define MVIN_FLAGS_HO_ONLY 0x5b          /* --ablation ho_only          (User_orient_kg_eh = 0) */
#define MVIN_FLAGS_HO_ONLY_UO_KG_EH 0x5f /* --ablation ho_only_uo_kg_eh (User_orient_kg_eh = 1) */
/* Supported: every setting of parameter_ablation.py with wide_deep = 1 -- bits 0..3 either way, bit 4 set, at most one of
 * bits 5 / 6; PS_O_ft = 0 needs p_hop >= 1.  Everything else -> MVIN_ERR_UNSUPPORTED from mvin_create.  The tuned kernels
 * (entity tables, entity groups, tcgen05) serve User_orient = User_orient_rela = 1 with one mix block; the other settings
 * run the generic per-level step (steps.cuh). */

/* The parameter set of MVIN._build_model (model.py:72-122) + the aggregators (aggregators.py:83-93), fp32,
 * device pointers.  The same struct describes a gradient set (same shapes) and Adam moment sets.
 * H = h_hop iterations per mix block, M = n_mix_hop mix blocks, L = H M = depth of the neighbourhood = number of
 * aggregators (M = 1: L = H).  Aggregator (i, n) of model.py:290 (iteration i of block n) sits at index n H + i. */
typedef struct mvin_params {
  float* user_emb;      /* [n_user, d]            user_emb_matrix_STWS        model.py:72-74   */
  float* entity_emb;    /* [n_entity, d]          entity_emb_matrix_STWS      model.py:76-78   */
  float* relation_emb;  /* [n_relation, d]        relation_emb_matrix_STWS    model.py:80-82   */
  float* relation_kge;  /* [n_relation, d, d]     relation_emb_KGE_matrix_STWS model.py:84-86  */
  float* mix_w;         /* [M, (H+1) d, d]        enti_transfer_matrix_list[n] model.py:91-98  */
  float* mix_b;         /* [M, d]                 enti_transfer_bias_list[n]                   */
  float* user_mlp_w;    /* [(p+1) d, d]           user_mlp_matrix ([p d, d] when PS_O_ft = 0) model.py:100-104 */
  float* user_mlp_b;    /* [d]                    user_mlp_bias               model.py:105-106 */
  float* transfer_w;    /* [L+1, d, d]            transfer_matrix_list[e]     model.py:107-116 */
  float* transfer_b;    /* [L+1, d]               transfer_matrix_bias[e]                      */
  float* h_item_w;      /* [2 d]                  h_emb_item_mlp_matrix       model.py:118-120 */
  float* h_item_b;      /* [1]                    h_emb_item_mlp_bias         model.py:121-122 */
  float* agg_w;         /* [L, d, d]              aggregator: weights         aggregators.py:83-85 */
  float* agg_b;         /* [L, d]                 aggregator: bias            aggregators.py:86-87 */
  float* agg_urh_w;     /* [L, 3 d]               aggregator: urh_weights     aggregators.py:89-91 */
  float* agg_urh_b;     /* [L]                    aggregator: urh_bias (created, never used) :92-93;
                           under PS_only the reference creates no aggregators: the four fields stay untouched */
} mvin_params_t;

int mvin_abi_version(void);
const char* mvin_last_error(void);

/* Lifetime.  mvin_create validates the configuration (MVIN_ERR_UNSUPPORTED for anything outside the
 * supported set) and records the device that is current on the calling thread. */
int mvin_create(const mvin_config_t* cfg, mvin_handle_t* out);
int mvin_destroy(mvin_handle_t h);

/* Bind the parameter tables (read by forward, updated by mvin_adam_step), the gradient buffers (written by
 * mvin_backward) and the packed adjacency built by mvin_pack_adjacency. */
int mvin_bind_params(mvin_handle_t h, const mvin_params_t* params);
int mvin_bind_grads(mvin_handle_t h, const mvin_params_t* grads);
int mvin_bind_adjacency(mvin_handle_t h, const int32_t* adj_packed /* [n_entity, 2, K] */);

/* adj_entity / adj_relation int64 [n_entity, K] (MVIN.__init__ args, model.py:7,19-20; built by
 * data_loader_user_set.py:375-388) -> packed int32 [n_entity][2][K]: row e = K neighbour ids then K relation
 * ids, one contiguous 8K-byte record per entity.  Device pointers. */
int mvin_pack_adjacency(const int64_t* adj_entity, const int64_t* adj_relation, int32_t n_entity, int32_t K,
                        int32_t* adj_packed, void* stream);

/* Row-sharded entity table (BASELINE.json C5: the 51 GB entity table of the 100 M-entity KG does not share one
 * GPU's HBM with its gradient and Adam state).  Replaces the single `entity_emb_matrix` variable of model.py:76-78
 * and every tf.nn.embedding_lookup on it (model.py:130,134,144,199,267) by a lookup through a shard table:
 * entity e lives in shard (e mod n_shards) at local row (e div n_shards); n_shards is a power of two <= 16.
 * entity_shards[g] / grad_shards[g] are DEVICE pointers valid in this process -- the local shard, and the peers'
 * shards mapped with CUDA IPC, so gathers are NVLink peer loads and gradient scatter-adds are NVLink peer
 * reductions issued by the same kernels (no separate exchange step).  Each shard is fp32
 * [ceil(n_entity / n_shards), d].  After this call params.entity_emb / grads.entity_emb (mvin_bind_params /
 * mvin_bind_grads) must point at THIS rank's shard: mvin_adam_step updates only that shard, and mvin_backward no
 * longer zero-fills the entity gradient -- the caller zeroes its shard and synchronises the ranks before the
 * step (peers scatter into it) and again before reading it.  Both arrays are host arrays of device pointers. */
int mvin_bind_entity_shards(mvin_handle_t h, int32_t n_shards, const float* const* entity_shards,
                            float* const* grad_shards);

/* Owner-side partial reduction of the deepest level for the row-sharded table (BASELINE.json C5 / north_star: "a
 * single all-to-all to route each batch's neighbor indices to the owning shard").  Replaces the leaf gather
 * tf.nn.embedding_lookup(entity_emb_matrix, entities[L]) of model.py:267 as consumed by
 * reduce_mean(probs * neighbor_vectors) (aggregators.py:141-144): that consumer is linear in the rows, so the rank that
 * OWNS a row contributes it to a partial sum and one d-vector per (parent node, owner) crosses NVLink instead of every
 * raw row.  The adjacency is replicated, so the owners re-derive child ids and attention from the parent's entity id:
 * the index routing is an all-gather of 4 bytes per leaf-level parent node (the caller's one NCCL collective).
 *
 * Geometry: n_src source ranks (1 when all shards live in one process), `rows` = B K^(h_hop-1) parent nodes per source.
 * Per source s, three peer-visible buffers: part[s] fp32 [n_shards][rows][d] (written by the owners, read by s),
 * gsu[s] fp32 [rows][d] (written by s, read by the owners), dot[s] fp32 [n_shards][rows] (written by the owners).
 * ids_all int32 [n_src][rows] is local.  Host arrays of device pointers valid in this process (CUDA-IPC mappings for
 * the peers').  n_src = 0 unbinds.  Sequence of one step, every rank, one stream (=> marks the caller's stream-ordered
 * collective that orders the ranks):
 *   mvin_xchg_expand (my ids -> ids_all[src_index]) => all-gather ids_all => mvin_xchg_owner_forward(owner = my shard)
 *   => barrier => mvin_forward, mvin_backward (the leaf level reads part[] and leaves gsu) => barrier =>
 *   mvin_xchg_owner_backward(owner) => barrier => mvin_xchg_finish_backward => all-reduce of the replicated gradients.
 * With all shards in one process the owner calls are made once per shard and no collective is needed. */
int mvin_xchg_bind(mvin_handle_t h, int32_t n_src, int32_t src_index, int64_t rows, const int32_t* ids_all,
                   float* const* part, float* const* gsu, float* const* dot);
int mvin_xchg_expand(mvin_handle_t h, const int64_t* item_indices, int32_t B, int32_t* ids_out, void* workspace,
                     void* stream);
int mvin_xchg_owner_forward(mvin_handle_t h, int32_t owner, void* workspace, void* stream);
int mvin_xchg_owner_backward(mvin_handle_t h, int32_t owner, void* workspace, void* stream);
int mvin_xchg_finish_backward(mvin_handle_t h, void* workspace, void* stream);

/* CUDA IPC helpers for the shard exchange between the one-process-per-GPU ranks of a box: export gives the 64-byte
 * cudaIpcMemHandle_t of the allocation containing dev_ptr and dev_ptr's offset in it; open maps a peer's
 * allocation into this process (once per allocation, peer access enabled lazily) and returns the pointer. */
#define MVIN_IPC_HANDLE_BYTES 64
int mvin_ipc_export(const void* dev_ptr, void* handle_out, int64_t* offset_out);
int mvin_ipc_open(const void* handle, int64_t offset, void** ptr_out);

/* Data-parallel ranks: the base loss and its gradient are divided by `global_batch` (0 = the batch of the call,
 * model.py:379-380 semantics) and the dense L2 terms (model.py:388-410) are multiplied by dense_l2_scale
 * (1 / world size), so that a SUM all-reduce of losses and gradients over the ranks equals the single-device
 * result on the concatenated batch. */
int mvin_set_batch_scale(mvin_handle_t h, int32_t global_batch, float dense_l2_scale);

/* Workspace (activations kept for backward + scratch), in bytes, for batch size B. */
size_t mvin_workspace_bytes(mvin_handle_t h, int32_t B);

/* MVIN.get_neighbors (model.py:243-256): entities[i] int64 [B, K^i] for i = 0..n_levels, relations[i] int64
 * [B, K^(i+1)] for i < n_levels, child k of node j at flat position j*K+k.  `entities` / `relations` are HOST
 * arrays of DEVICE pointers.  Bit-exact integer path. */
int mvin_get_neighbors(mvin_handle_t h, const int64_t* item_indices, int32_t B, int32_t n_levels,
                       int64_t* const* entities, int64_t* const* relations, void* stream);

/* Forward pass = model.py:125-159 (ripple-memory lookups, get_neighbors, _key_addressing,
 * aggregate_delta_whole, score).  Feed contract = model.py:49-64 with the memories stacked per hop:
 *   user_indices, item_indices int64 [B]; mem_h / mem_r / mem_t int32 [max(1,p), B, m].
 * Outputs: scores [B] (model.py:158) and scores_normalized [B] (model.py:159); either may be NULL.
 * Activations needed by mvin_backward stay in `workspace`. */
int mvin_forward(mvin_handle_t h, const int64_t* user_indices, const int64_t* item_indices,
                 const int32_t* mem_h, const int32_t* mem_r, const int32_t* mem_t, int32_t B, float* scores,
                 float* scores_normalized, void* workspace, void* stream);

/* importance_list_0 [B,1,K] and importance_list_1 [B,K,K] (model.py:319-323; aggregators.py:139): attention
 * of aggregator i=0 at hops 0 and 1, for the batch of the last mvin_forward.  imp1 may be NULL (and is unused
 * when h_hop = 1). */
int mvin_importance(mvin_handle_t h, float* imp0, float* imp1, void* workspace, void* stream);

/* Loss (model.py:378-412) and the gradient of every parameter (TF autodiff behind model.py:414) for the batch
 * of the last mvin_forward on the same workspace.  Gradient buffers are overwritten (zero-filled first, so an
 * unused row has gradient 0, matching TF's dense-equivalent sparse Adam).  losses_out (device, 4 floats):
 * loss, base_loss, l2_loss, l2_agg_loss. */
int mvin_backward(mvin_handle_t h, const float* labels, int32_t B, float* losses_out, void* workspace,
                  void* stream);

/* tf.train.AdamOptimizer(lr).minimize (model.py:414) with TF1 defaults and epsilon placement:
 * lr_t = lr sqrt(1-b2^t)/(1-b1^t); var -= lr_t m / (sqrt(v)+eps), dense over every bound parameter. */
int mvin_adam_step(mvin_handle_t h, const mvin_params_t* m, const mvin_params_t* v, float lr, float beta1,
                   float beta2, float eps, int32_t step /* t, 1-based */, void* stream);

/* One training step from HOST buffers (what `MVIN.train(sess, feed_dict)` does, model.py:416-417, including
 * the host->device feed copy TF performs inside Session.run): copies the feed H2D on `stream`, runs forward +
 * backward (+ Adam when adam_m != NULL), copies the 4 loss scalars back to losses_host and synchronises the
 * stream.  `staging` is a device buffer of at least mvin_feed_bytes(h, B) bytes. */
size_t mvin_feed_bytes(mvin_handle_t h, int32_t B);
int mvin_train_step_host(mvin_handle_t h, const int64_t* user_indices, const int64_t* item_indices,
                         const float* labels, const int32_t* mem_h, const int32_t* mem_r, const int32_t* mem_t,
                         int32_t B, void* staging, void* workspace, const mvin_params_t* adam_m,
                         const mvin_params_t* adam_v, float lr, int32_t step, float* losses_host, void* stream);

/* Device-resident feed path -- replaces the host feed assembly get_feed_dict (train.py:112-122, util.py:208-218):
 * the packed ripple sets  user_triplet_set int32 [n_user, max(1,p), 3, n_memory]  (data_loader_user_set.py:402; one
 * [p, 3, m] block per user) are uploaded once and bound; mvin_gather_feed builds memories_{h,r,t} [max(1,p), B, m] of a
 * batch from user_indices on the device; mvin_train_step_users_host is mvin_train_step_host with only
 * user / item / label crossing the bus (20 bytes per pair instead of 20 + 12 p m). */
int mvin_bind_user_triplets(mvin_handle_t h, const int32_t* user_triplet_set);
int mvin_gather_feed(mvin_handle_t h, const int64_t* user_indices, int32_t B, int32_t* mem_h, int32_t* mem_r,
                     int32_t* mem_t, void* stream);
int mvin_train_step_users_host(mvin_handle_t h, const int64_t* user_indices, const int64_t* item_indices,
                               const float* labels, int32_t B, void* staging, void* workspace,
                               const mvin_params_t* adam_m, const mvin_params_t* adam_v, float lr, int32_t step,
                               float* losses_host, void* stream);

/* Sampled fixed-fan-out adjacency on the device -- replaces contruct_random_adj (data_loader_user_set.py:375-388) on
 * the undirected CSR of construct_kg (:324-343): indptr int64 [n_entity + 1], nbr / rel int32 [2 T] (device).  Rows with
 * degree >= K get K distinct edges (uniform subset, uniform order), rows with 0 < degree < K get K draws with
 * replacement, isolated entities stay zero.  Outputs (any may be NULL except one of adj_packed / adj_entity): the packed
 * int32 [n_entity][2][K] record mvin_bind_adjacency takes, the reference's int64 [n_entity, K] pair, and the chosen
 * edge slots (absolute CSR positions, -1 for isolated rows) for tests.  Reproducible for a given seed. */
int mvin_sample_adjacency(const int64_t* indptr, const int32_t* nbr, const int32_t* rel, int32_t n_entity, int32_t K,
                          uint64_t seed, int32_t* adj_packed, int64_t* adj_entity, int64_t* adj_relation,
                          int64_t* picked_edges, void* stream);

/* Ripple sets on the device -- replaces get_user_triplet_set (data_loader_user_set.py:392-441) and its 12-process
 * pool: per user and hop, every source entity (the user's positive items at hop 0: hist_items[hist_ptr[u] ..
 * hist_ptr[u+1]); the previous hop's tails afterwards) offers a random min(degree, n_neighbor)-subset of its edges
 * (the reference passes n_neighbor = 16); n_memory of those candidates are drawn, without replacement when there are
 * enough; an empty candidate list repeats the previous hop.  Output: user_triplet_set int32 [n_user, max(1,p), 3,
 * n_memory] (rows heads, relations, tails) -- the array mvin_bind_user_triplets takes.  `slots` (optional, [n_user,
 * max(1,p), n_memory]) receives the candidate slot of each draw for tests.  n_memory <= 64, n_neighbor <= 16. */
int mvin_build_ripple_sets(const int64_t* indptr, const int32_t* nbr, const int32_t* rel, const int64_t* hist_ptr,
                           const int32_t* hist_items, int32_t n_user, int32_t p_hop, int32_t n_memory, int32_t n_neighbor,
                           uint64_t seed, int32_t* user_triplet_set, int64_t* slots, void* stream);

/* CTR evaluation on the device -- replaces the per-batch sklearn calls of MVIN.eval (model.py:419-426, util.py:44-56):
 * out3 = {roc_auc_score(labels, scores), mean((scores >= 0.5) == labels), f1_score(labels, scores >= 0.5)} from exact
 * pair counts (ties count 1/2, as the trapezoidal ROC area does).  scores / labels / out3 device pointers;
 * scratch40 = 40 bytes of device scratch. */
int mvin_ctr_metrics(mvin_handle_t h, const float* scores_normalized, const float* labels, int32_t B, float* out3,
                     void* scratch40, void* stream);

/* Double-buffered input pipeline for a training loop that knows its next batch (train.py:58-64 iterates a shuffled
 * array): mvin_feed_prefetch copies the NEXT batch's feed (same arguments as mvin_train_step_host, host buffers,
 * pinned for a truly asynchronous copy) into `staging` on an internal copy stream, beside whatever runs on `stream`;
 * mvin_train_step_prefetched then runs forward + backward (+ Adam) from that staging buffer and blocks like
 * mvin_train_step_host.  Two slots (0 / 1), each with its own staging buffer of mvin_feed_bytes(), are used alternately:
 * prefetch(batch i+1 -> slot (i+1)%2) is called before step(batch i from slot i%2). */
int mvin_feed_prefetch(mvin_handle_t h, const int64_t* user_indices, const int64_t* item_indices, const float* labels,
                       const int32_t* mem_h, const int32_t* mem_r, const int32_t* mem_t, int32_t B, void* staging,
                       int32_t slot, void* stream);
int mvin_train_step_prefetched(mvin_handle_t h, int32_t B, void* staging, int32_t slot, void* workspace,
                               const mvin_params_t* adam_m, const mvin_params_t* adam_v, float lr, int32_t step,
                               float* losses_host, void* stream);

/* Top-K evaluation metrics on the device -- replaces the per-user Python sort and metric calls of topk_eval
 * (util.py:183-197; metrics.py:3-31, 34-37, 97-100).  Per user u: scores[u, :n_cand[u]] of its candidate items in
 * candidate order (get_scores output, util.py:159-181), relevant[u, i] = candidate i is in ref_user[user],
 * n_answers[u] = len(ref_user[user]); k_list ascending, nk <= 8.  Outputs [n_users, nk]: precision@k, recall@k and
 * ndcg@k exactly as the reference computes them (stable descending sort; the hit list of ndcg is built over the top
 * k_list[-1] items).  All pointers device; the caller averages over users (util.py:199-201). */
int mvin_topk_metrics(const float* scores, const uint8_t* relevant, const int32_t* n_cand, const int32_t* n_answers,
                      int32_t n_users, int32_t max_cand, const int32_t* k_list, int32_t nk, float* precision, float* recall,
                      float* ndcg, void* stream);

/* Number of kernels the library has launched on behalf of this handle since creation (bench evidence). */
int64_t mvin_launch_count(mvin_handle_t h);

/* Per-kernel device timing for bench.py's roofline figure.  While enabled the library records one CUDA event on
 * the launch stream after each of its kernel launches; mvin_profile_read synchronises on the last one and writes
 * "name:total_ms:launches;..." (accumulated since the previous read) into buf, then clears the records. */
int mvin_profile_enable(mvin_handle_t h, int32_t on);
int mvin_profile_read(mvin_handle_t h, char* buf, size_t buflen);

/* Self-test of the tcgen05 (5th-generation tensor core) building block the row kernels use for their d x d maps:
 * C[M, D] = A[M, D] . W[D, D]^T in 3xTF32 (fp32-level accuracy), device pointers, D in {32, 64}.  Not part of the
 * model path; tests/test_umma.py checks it against an fp64 product. */
int mvin_test_umma_gemm(const float* A, const float* W, float* C, int64_t M, int32_t D, void* stream);
/* Same for the transposed product used by the weight gradients: dump[128, D] receives the raw accumulator lanes of
 * dW = A^T . G (A, G [M, D]), accumulated over all 128-row tiles by one CTA. */
int mvin_test_umma_dw(const float* A, const float* G, float* dump, int64_t M, int32_t D, int32_t variant,
                      void* stream);

/* Self-test of the split-bf16 tcgen05 block of the backward kernels (csrc/umma_bf.cuh): one tile buffer per operand serves
 * both  C[M, D] = G . W^T  (contraction over features)  and  dW = A^T . G  (contraction over rows, accumulated over all
 * 128-row tiles in tensor memory).  n_planes in {2, 3}; dump[128, D] receives the raw accumulator lanes of dW. */
int mvin_test_umma_bf16(const float* A, const float* G, const float* W, float* C, float* dump, int64_t M, int32_t D,
                        int32_t n_planes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVIN_B200_H */
