// steps.cuh -- one forward / backward step of the MVIN hot path as a sequence of kernel launches, templated on the
// embedding dimension.  Forward = model.py:125-159 of the reference (src/model/MVIN/), backward = TF autodiff of
// model.py:378-412.  The factorisation the kernels implement is spelled out in DESIGN.md section 3 and has a CPU twin
// in tests/fused_model.py.  Instantiated once per dimension by mvin_steps.cu (-DMVIN_DIM=...), so the five
// instantiations compile in parallel.
#pragma once
#include "host.cuh"

namespace mvin_host {

// tcgen05 versions of the relation-batched contractions of the user side (gemm_tc.cuh): d in {32, 64}, batches of at least
// 512 pairs (MVIN_B200_TCGEMM=0 / 2: never / always)
template <int D>
bool use_tc_gemm(mvin_handle_t h, int B) {
  if constexpr (D == 32 || D == 64) return h->tcg_mode == 2 || (h->tcg_mode == 1 && B >= 512);
  return false;
}
inline dim3 rel_gemm_grid(mvin_handle_t h, int B, int nr) {
  const int tiles = (B + 127) / 128;
  int ns = h->sm_count / tiles;
  if (ns < 1) ns = 1;
  if (ns > nr) ns = nr;
  return dim3(tiles, ns);
}

// per-entity leaf aggregate: the register-blocked kernel of group.cuh where the children fit the registers (d K <= 2048),
// else leaf_entity_kernel of level.cuh (MVIN_B200_LEAF_REG=0 forces the latter)
template <int D, bool BWD>
int launch_leaf_entity(mvin_handle_t h, cudaStream_t st, LeafEntArgs& a, const char* name) {
  const mvin_config_t& c = h->cfg;
  constexpr int G = 32 / (D / 4);
  const int kpl = (a.K + G - 1) / G;
  a.chunk = leaf_chunk(h, c.n_entity);
  const long want = ((long)c.n_entity + LEAF_NW * a.chunk - 1) / (LEAF_NW * a.chunk);
  const long cap = (long)h->sm_count * 8;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  int rc;
  static const bool reg_on = [] { const char* e = getenv("MVIN_B200_LEAF_REG"); return !(e && e[0] == '0'); }();
  // (the forward -- K independent row loads and one weighted sum -- measured faster on the older kernel: 0.09 vs 0.12 ms)
  if (BWD && reg_on && grp_supported(D, a.K) && h->n_shards == 1) {
    const size_t sm = leaf_entity_reg_smem(a.n_rel, BWD);
#define MVIN_LEAF_GO(KC_)                                                              \
  do {                                                                                 \
    if ((rc = set_smem(leaf_entity_reg_kernel<D, BWD, KC_>, sm))) return rc;           \
    MVIN_LAUNCH((leaf_entity_reg_kernel<D, BWD, KC_>), grid, LEAF_NT, sm, st, a);      \
  } while (0)
    if (kpl <= 4) MVIN_LEAF_GO(4);
    else if (kpl <= 8) MVIN_LEAF_GO(8);
    else MVIN_LEAF_GO(16);
#undef MVIN_LEAF_GO
  } else {
    const size_t sm = leaf_entity_smem(a.n_rel);
    if ((rc = set_smem(leaf_entity_kernel<D, BWD>, sm))) return rc;
    MVIN_LAUNCH((leaf_entity_kernel<D, BWD>), grid, LEAF_NT, sm, st, a);
  }
  LAUNCH_CHECK(h, name);
  return MVIN_OK;
}

// virt_group_kernel (group.cuh) with the compile-time slot count that fits the launch's K
template <int D, bool BWD>
int launch_group(mvin_handle_t h, cudaStream_t st, const GroupArgs& ga, const char* name) {
  constexpr int G = 32 / (D / 4);
  const int kpl = (ga.K + G - 1) / G;
  const size_t smg = grp_smem(ga.n_rel, BWD);
  const long nwin = (ga.rows + GRP_WIN - 1) / GRP_WIN;
  const long cap = (long)h->sm_count * 16;
  const int sp = (BWD && h->group_split == 2 && kpl > 8) ? 2 : 1;
  const long want = (sp * nwin + GRP_NW - 1) / GRP_NW;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  int rc;
#define MVIN_GROUP_GO(SP_, KC_)                                                        \
  do {                                                                                 \
    if ((rc = set_smem(virt_group_kernel<D, BWD, SP_, KC_>, smg))) return rc;          \
    MVIN_LAUNCH((virt_group_kernel<D, BWD, SP_, KC_>), grid, GRP_NT, smg, st, ga);     \
  } while (0)
  if (kpl <= 4) MVIN_GROUP_GO(1, 4);
  else if (kpl <= 8) MVIN_GROUP_GO(1, 8);
  else if (sp == 2) MVIN_GROUP_GO(2, 16);
  else MVIN_GROUP_GO(1, 16);
#undef MVIN_GROUP_GO
  LAUNCH_CHECK(h, name);
  return MVIN_OK;
}

// variants of the step outside --ablation all with one mix block (defined at the end of this file)
template <int D>
int forward_ps_only(mvin_handle_t h, const int64_t* item, const int32_t* mem_h, const int32_t* mem_r, const int32_t* mem_t,
                    int B, float* scores, float* scores_norm, void* ws, cudaStream_t st);
template <int D>
int backward_ps_only(mvin_handle_t h, const float* labels, int B, float* losses_out, void* ws, cudaStream_t st);
template <int D>
int forward_mix_impl(mvin_handle_t h, const int64_t* item, const int32_t* mem_h, const int32_t* mem_r, const int32_t* mem_t,
                     int B, float* scores, float* scores_norm, void* ws, cudaStream_t st);
template <int D>
int backward_mix_impl(mvin_handle_t h, const float* labels, int B, float* losses_out, void* ws, cudaStream_t st);

// the user side of the forward pass (model.py:125-134, :161-240): user_o = key addressing over the ripple memories
template <int D>
int user_side_forward(mvin_handle_t h, const int64_t* item, const int32_t* mem_h, const int32_t* mem_r, const int32_t* mem_t,
                      int B, void* ws, cudaStream_t st) {
  using C = TC<D>;
  const mvin_config_t& c = h->cfg;
  const int p = c.p_hop, m = c.n_memory, nr = c.n_relation;
  const Layout L = handle_layout(h, B);
  const mvin_params_t& P = h->P;
  int rc;
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
    bool done = false;
    if constexpr (D == 32 || D == 64) {
      if (use_tc_gemm<D>(h, B)) {
        RelGemmArgs ra;
        memset(&ra, 0, sizeof(ra));
        ra.V = at<float>(ws, L.Vbuf); ra.RK = P.relation_kge; ra.Q = at<float>(ws, L.Q); ra.B = B; ra.n_rel = nr;
        ra.RKs = at<unsigned char>(ws, L.RKs);
        // relation operands split and staged once per step (also read by the backward's dv kernel)
        MVIN_LAUNCH((rel_stage_kernel<D>), dim3(nr, 2), RG_NT, 0, st, P.relation_kge, at<unsigned char>(ws, L.RKs));
        LAUNCH_CHECK(h, "rel_stage");
        if ((rc = set_smem(relq_tc_kernel<D>, relq_tc_smem<D>()))) return rc;
        MVIN_LAUNCH((relq_tc_kernel<D>), rel_gemm_grid(h, B, nr), RG_NT, relq_tc_smem<D>(), st, ra);
        LAUNCH_CHECK(h, "gemm_q");
        done = true;
      }
    }
    if (!done && (rc = run_gemm(h, st, g, "gemm_q"))) return rc;
  }
  {
    UserArgs a;
    memset(&a, 0, sizeof(a));
    a.E = h->etab; a.item = item; a.RK = P.relation_kge; a.w_hi = P.h_item_w;
    a.W_user = P.user_mlp_w; a.b_user = P.user_mlp_b;
    if (!(c.flags & MVIN_FLAG_PS_O_FT)) {
      // PS_O_ft = 0 (model.py:204-206, :232-236): user_h_set is not part of o_list and user_mlp_matrix is [p D, D].  The
      // kernels keep their p + 1 slots; slot 0 meets a zero block of a padded copy of the weights
      float* Wpad = at<float>(ws, L.Wpad);
      CUDA_TRY(cudaMemsetAsync(Wpad, 0, sizeof(float) * D * D, st));
      CUDA_TRY(cudaMemcpyAsync(Wpad + D * D, P.user_mlp_w, sizeof(float) * (size_t)p * D * D, cudaMemcpyDeviceToDevice, st));
      a.W_user = Wpad;
    }
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
  return MVIN_OK;
}

template <int D>
int forward_impl(mvin_handle_t h, const int64_t* item, const int32_t* mem_h, const int32_t* mem_r,
                 const int32_t* mem_t, int B, float* scores, float* scores_norm, void* ws, cudaStream_t st) {
  using C = TC<D>;
  const mvin_config_t& c = h->cfg;
  const int K = c.neighbor_sample_size, H = c.h_hop, nr = c.n_relation;
  if (c.flags & MVIN_FLAG_PS_ONLY) return forward_ps_only<D>(h, item, mem_h, mem_r, mem_t, B, scores, scores_norm, ws, st);
  if (generic_step(c)) return forward_mix_impl<D>(h, item, mem_h, mem_r, mem_t, B, scores, scores_norm, ws, st);
  const Layout L = handle_layout(h, B);
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
    if (L.table) {
      // parameters only: composed maps M1_h, M2_h, c_h of the entity tables
      MVIN_LAUNCH((compose_kernel), dim3(H, COMPOSE_SPLIT), 256, 0, st, P.transfer_w, P.transfer_b, P.agg_w, P.agg_b, D,
                  1.f / (float)K, at<float>(ws, L.Mc), at<float>(ws, L.cst));
      LAUNCH_CHECK(h, "compose");
    }
    // table mode stamps the entities of EVERY level (each needs its row of A_h); the row kernels only the deepest
    MVIN_LAUNCH((seed_kernel), (unsigned)((B + 255) / 256), 256, 0, st, item, B, at<int32_t>(ws, L.ent[0]),
                (H == 1 || L.table) ? stamp : nullptr);
    LAUNCH_CHECK(h, "seed");
    for (int lv = 0; lv + 1 < H; ++lv) {
      const long n = L.rows[lv] * K;
      // table mode never reads the ids of the deepest level (they are re-read from the adjacency records); levels 0 / 1
      // are still written for mvin_importance
      int32_t* out = (L.table && lv + 1 == H - 1 && lv + 1 >= 2) ? nullptr : at<int32_t>(ws, L.ent[lv + 1]);
      if (!out && L.group) continue;     // stamped per DISTINCT parent entity after the group sort below
      MVIN_LAUNCH((expand_kernel), (unsigned)((n + 255) / 256), 256, 0, st, at<int32_t>(ws, L.ent[lv]), h->adj, L.rows[lv], K,
                                                                 out, (lv + 1 == H - 1 || L.table) ? stamp : nullptr,
                                                                 L.table ? 1 << (lv + 1) : 1);
      LAUNCH_CHECK(h, "expand");
    }
    // entity-group evaluation of the table-gather level (group.cuh): counting sort of the level-(H-2) rows by entity
    if (L.group) {
      const long rows = L.rows[H - 2];
      const long ne = c.n_entity;
      int32_t* cnt = at<int32_t>(ws, L.gcnt);
      int32_t* off = at<int32_t>(ws, L.goff);
      int32_t* tot = at<int32_t>(ws, L.gtot);
      const int32_t* ent = at<int32_t>(ws, L.ent[H - 2]);
      const unsigned nb = (unsigned)((ne + SCAN_PER_BLOCK - 1) / SCAN_PER_BLOCK);
      CUDA_TRY(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)ne, st));
      MVIN_LAUNCH((grp_count_kernel), (unsigned)((rows + 255) / 256), 256, 0, st, ent, rows, cnt);
      LAUNCH_CHECK(h, "grp_count");
      MVIN_LAUNCH((scan_block_kernel), nb, SCAN_NT, 0, st, (const int32_t*)cnt, ne, off, tot);
      LAUNCH_CHECK(h, "grp_scan");
      if (nb > 1) {
        MVIN_LAUNCH((scan_block_kernel), 1, SCAN_NT, 0, st, (const int32_t*)tot, (long)nb, tot + SCAN_PER_BLOCK + 1, (int32_t*)nullptr);
        LAUNCH_CHECK(h, "grp_scan");
        MVIN_LAUNCH((scan_add_kernel), nb, SCAN_NT, 0, st, off, ne, (const int32_t*)(tot + SCAN_PER_BLOCK + 1));
        LAUNCH_CHECK(h, "grp_scan");
      }
      CUDA_TRY(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)ne, st));
      MVIN_LAUNCH((grp_fill_kernel), (unsigned)((rows + 255) / 256), 256, 0, st, ent, rows, (const int32_t*)off, cnt,
                  at<int32_t>(ws, L.gorder), at<int32_t>(ws, L.gesort));
      LAUNCH_CHECK(h, "grp_fill");
      if (H - 1 >= 2) {
        // stamps of the deepest level: the children of every DISTINCT level-(H-2) entity (cnt > 0 after the fill), instead
        // of one store per level-(H-1) row (C4: 2.5 M instead of 16.8 M)
        const long n = ne * K;
        MVIN_LAUNCH((stamp_children_kernel), (unsigned)((n + 255) / 256), 256, 0, st, (const int32_t*)cnt, h->adj, ne, K, stamp, 1 << (H - 1));
        LAUNCH_CHECK(h, "expand");
      }
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
      if ((rc = launch_leaf_entity<D, false>(h, st, a, "leaf_entity_fwd"))) return rc;
    }
    // table mode: A_h = E M1_h + Se M2_h + c_h for every stamped entity
    if (L.table) {
      TableArgs a;
      memset(&a, 0, sizeof(a));
      a.stamp = stamp; a.E = h->etab.base; a.Se = at<float>(ws, L.Se); a.M = at<float>(ws, L.Mc); a.cst = at<float>(ws, L.cst);
      a.A = at<float>(ws, L.Atab); a.n_entity = c.n_entity;
      const size_t sm = table_fwd_smem<D>();
      if ((rc = set_smem(table_fwd_kernel<D>, sm))) return rc;
      const long tiles = ((long)c.n_entity + C::R - 1) / C::R;
      long cap = (long)h->sm_count * resident_ctas(h, table_fwd_kernel<D>, C::NT, sm) / H;   // one resident wave over the H levels
      if (cap < 1) cap = 1;
      MVIN_LAUNCH((table_fwd_kernel<D>), dim3((unsigned)(tiles < cap ? tiles : cap), H), C::NT, sm, st, a);
      LAUNCH_CHECK(h, "table_fwd");
    }
  }
  if ((rc = user_side_forward<D>(h, item, mem_h, mem_r, mem_t, B, ws, st))) return rc;
  // --ablation no_kg_eh_uo (User_orient_kg_eh = 0, model.py:152-156): the KG side is oriented by the raw user embedding
  // U[user] instead of user_o; the score still uses user_o.  HO_only (model.py:146-150): the SCORE uses U[user]; the KG
  // side is oriented by user_o (ho_only_uo_kg_eh) or by U[user] (ho_only)
  const bool kg_eh = (c.flags & MVIN_FLAG_KG_EH) != 0;
  const bool ho_only = (c.flags & MVIN_FLAG_HO_ONLY) != 0;
  const float* u_kg = at<float>(ws, L.u);
  if (!kg_eh || ho_only) {
    if (!h->user) return fail(MVIN_ERR_INVALID, "user_indices are required when User_orient_kg_eh = 0 or HO_only = 1");
    const long n = (long)B * C::LPR;
    MVIN_LAUNCH((gather_user_kernel<D>), (unsigned)((n + 255) / 256), 256, 0, st, h->user, P.user_emb, B, at<float>(ws, L.ukg),
                at<int32_t>(ws, L.user32));
    LAUNCH_CHECK(h, "gather_user");
    if (!kg_eh) u_kg = at<float>(ws, L.ukg);
  }
  const float* u_score = ho_only ? at<float>(ws, L.ukg) : at<float>(ws, L.u);
  par.join(0);
  // user-oriented transform of levels 0..L-1, one launch   (model.py:270-283)
  {
    const size_t sm = transform_fwd_smem<D>();
    if ((rc = set_smem(transform_fwd_kernel<D>, sm))) return rc;
    TransformArgs a;
    memset(&a, 0, sizeof(a));
    long rows[MAX_LV];
    const int ntl = L.table ? 1 : H;                 // table mode: only T[0] is a row buffer (the mix layer reads it)
    for (int lv = 0; lv < ntl; ++lv) {
      TransformLevel& t = a.lv[lv];
      t.ent = at<int32_t>(ws, L.ent[lv]);
      t.W = P.transfer_w + (long)lv * D * D; t.b = P.transfer_b + (long)lv * D;
      t.T = at<float>(ws, L.V[0][lv]);
      t.rows = rows[lv] = L.rows[lv]; t.rpp = (int)(L.rows[lv] / B); t.rpp_magic = div_magic(t.rpp);
        t.stream = stream_level(h, L.rows[lv], D);
    }
    a.nlev = ntl; a.E = h->etab; a.u = u_kg;
    bool done = false;
    if constexpr (D == 32 || D == 64) {
      if (!L.table && use_tc_path(h, L.rows[H - 1])) {
        const size_t smt = transform_fwd_tc_smem<D>();
        if ((rc = set_smem(transform_fwd_tc_kernel<D>, smt))) return rc;
        const int grid = partition_grid(rows, H, TT<D>::R,
                                        h->sm_count * resident_ctas(h, transform_fwd_tc_kernel<D>, TT<D>::NT, smt), a.cta_end);
        MVIN_LAUNCH((transform_fwd_tc_kernel<D>), grid, TT<D>::NT, smt, st, a);
        done = true;
      }
    }
    if (!done) {
      const int grid = partition_grid(rows, ntl, C::R, h->sm_count * resident_ctas(h, transform_fwd_kernel<D>, C::NT, sm), a.cta_end);
      MVIN_LAUNCH((transform_fwd_kernel<D>), grid, C::NT, sm, st, a);
    }
    LAUNCH_CHECK(h, "transform_fwd");
  }
  // table mode: C_h = u (M1_h + M2_h) per pair, then the iteration-0 outputs of every level but the deepest as rows
  // levels 0 .. max(H - 2, 0); the fused group mode (H >= 3) rebuilds level H - 2 on the fly instead (group.cuh)
  const bool gfuse = L.group && H >= 3;
  const int n_virt_rows = H >= 2 ? H - 1 - (gfuse ? 1 : 0) : 1;
  if (L.table) {
    GemmArgs g = gemm_args();
    g.A = u_kg; g.sa_m = D; g.sa_k = 1; g.bsA = 0;
    g.B = at<float>(ws, L.Mc) + (long)TBL_MSUM * D * D; g.sb_k = D; g.sb_n = 1; g.bsB = (long)TBL_NM * D * D;
    g.C = at<float>(ws, L.Cp); g.ldc = D; g.bsC = (long)B * D;
    g.M = B; g.N = D; g.K = D; g.nbatch = H;
    if ((rc = run_gemm(h, st, g, "gemm_cp"))) return rc;
    VirtArgs a;
    memset(&a, 0, sizeof(a));
    long end = 0;
    for (int lv = 0; lv < n_virt_rows; ++lv) {
      VirtLevel& t = a.lv[lv];
      t.ent = at<int32_t>(ws, L.ent[lv]);
      t.A = at<float>(ws, L.Atab) + (long)lv * c.n_entity * D;
      t.Cp = at<float>(ws, L.Cp) + (long)lv * B * D;
      t.V = at<float>(ws, L.V[1][lv]);
      t.rpp_magic = div_magic(L.rows[lv] / B);
      end += L.rows[lv];
      a.end[lv] = end;
    }
    a.nlev = n_virt_rows;
    const long n = end * C::LPR;
    const long cap = (long)h->sm_count * 8;
    const long want = (n + 255) / 256;
    MVIN_LAUNCH((virt_rows_kernel<D, false>), (unsigned)(want < cap ? want : cap), 256, 0, st, a);
    LAUNCH_CHECK(h, "virt_rows_fwd");
  }
  // aggregation iterations (model.py:286-307): one launch per iteration, every level of it
  {
    const size_t sm_leaf = agg_fwd_smem<D, true>(K, nr), sm_in = agg_fwd_smem<D, false>(K, nr);
    if ((rc = set_smem(agg_fwd_kernel<D, true>, sm_leaf))) return rc;
    if ((rc = set_smem(agg_fwd_kernel<D, false>, sm_in))) return rc;
    static const char* names[MAX_L] = {"agg_fwd_0", "agg_fwd_1", "agg_fwd_2"};
    if (L.group) {
      // neighbour phase of (iteration 1, level H-2) per entity group: Y = self + agg
      GroupArgs ga;
      memset(&ga, 0, sizeof(ga));
      ga.order = at<int32_t>(ws, L.gorder); ga.esort = at<int32_t>(ws, L.gesort); ga.adj = h->adj;
      ga.s = at<float>(ws, L.s) + nr;
      ga.tab = at<float>(ws, L.Atab) + (long)(H - 1) * c.n_entity * D;
      ga.Cp = at<float>(ws, L.Cp) + (long)(H - 1) * B * D;
      ga.self = at<float>(ws, L.V[1][H - 2]); ga.Y = at<float>(ws, L.Y[1][H - 2]);
      if (gfuse) {
        ga.fuse = 1;
        ga.tab_self = at<float>(ws, L.Atab) + (long)(H - 2) * c.n_entity * D;
        ga.Cp_self = at<float>(ws, L.Cp) + (long)(H - 2) * B * D;
      }
      ga.rows = L.rows[H - 2]; ga.rpp_magic = div_magic(L.rows[H - 2] / B); ga.K = K; ga.n_rel = nr;
      if ((rc = launch_group<D, false>(h, st, ga, "group_fwd"))) return rc;
    }
    for (int i = L.table ? 1 : 0; i < H; ++i) {
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
        if (t.leaf) {
          t.SU = at<float>(ws, L.SU);
        } else if (L.table && i == 1 && lv == H - 2 && L.group) {
          t.preagg = 1;                                      // Y = self + agg came from virt_group_kernel
        } else if (gfuse && i == 1 && lv == H - 3) {         // children = the fused (never materialised) level H - 2
          t.virt = 1;
          t.tab = at<float>(ws, L.Atab) + (long)(H - 2) * c.n_entity * D;
          t.Cp = at<float>(ws, L.Cp) + (long)(H - 2) * B * D;
        } else if (L.table && i == 1 && lv == H - 2) {       // children = iteration 0 of the deepest level, from its table
          t.virt = 1;
          t.tab = at<float>(ws, L.Atab) + (long)(H - 1) * c.n_entity * D;
          t.Cp = at<float>(ws, L.Cp) + (long)(H - 1) * B * D;
        } else {
          t.child = at<float>(ws, L.V[i][lv + 1]);
        }
      }
      a.adj = h->adj; a.s = at<float>(ws, L.s) + (long)i * nr;
      a.Wa = P.agg_w + (long)i * D * D; a.ba = P.agg_b + (long)i * D;
      a.K = K; a.n_rel = nr;
      if (i == 0) {
        a.E = h->etab; a.u = u_kg;
        a.Se = L.entity_leaf ? at<float>(ws, L.Se) : nullptr;
        a.Wt = P.transfer_w + (long)H * D * D; a.bt = P.transfer_b + (long)H * D;
        if (h->xchg.on) {                                    // leaf rows already reduced by their owners (exchange.cuh)
          if (L.rows[H - 1] != h->xchg.rows)
            return fail(MVIN_ERR_STATE, "exchange buffers are bound for %ld leaf-level nodes, this batch has %ld", h->xchg.rows,
                        L.rows[H - 1]);
          a.Xpart = h->xchg.part[h->xchg.src_index]; a.xG = h->n_shards;
        }
      }
      bool done = false;
      if constexpr (D == 32 || D == 64) {
        if (i == 0 && !h->xchg.on && use_tc_path(h, L.rows[H - 1])) {     // leaf iteration; inner-only iterations are faster on mma.sync
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
        // The forward gather keeps its register loads: with 35 KB of shared memory six CTAs per SM hold 196 KB of rows in
        // flight, more than the two CTAs a ring leaves room for (measured at C4: 0.49 ms against 0.84 ms with the ring;
        // MVIN_B200_RING=2 selects the ring here too).  The backward, whose loads compete with its red.global.add for
        // issue slots, gains from it (2.32 -> 1.99 ms).
        const bool ring = L.table && i == 1 && RowRing<D>::ENABLED && h->ring_mode == 2;
        const size_t sm_i = sm_in + (ring ? RowRing<D>::bytes() : 0);
        if (ring && (rc = set_smem(agg_fwd_kernel<D, false>, sm_i))) return rc;
        a.ring = ring ? 1 : 0;
        const int grid = make_tile_list(a.tl, rows, nlev, C::R,
                                        h->sm_count * resident_ctas(h, agg_fwd_kernel<D, false>, C::NT, sm_i), h->d_sched);
        MVIN_LAUNCH((agg_fwd_kernel<D, false>), grid, C::NT, sm_i, st, a);
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
                                                                        u_score, B, H + 1, at<float>(ws, L.item),
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
  // Hm = iterations per mix block (--h_hop), M = mix blocks, H = Hm M = depth = number of aggregators (n_mix_hop = 1: H = Hm)
  const int Hm = c.h_hop, M = c.n_mix_hop, H = Hm * M, p = c.p_hop, nr = c.n_relation;
  const Layout L = handle_layout(h, B);
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
    const float am = (c.flags & MVIN_FLAG_PS_ONLY) ? 0.f : 1.f;                    // PS_only: no aggregators (model.py:393)
    add(P.user_emb, G.user_emb, (long)c.n_user * D, l2a, 1.f, 1);                 // model.py:392
    add(P.relation_emb, G.relation_emb, (long)nr * D, l2w, 1.f, 0);               // :388
    add(P.relation_kge, G.relation_kge, (long)nr * D * D, 0.f, 0.f, 0);
    add(P.mix_w, G.mix_w, (long)M * (Hm + 1) * D * D, l2a, 1.f, 1);               // :400-401
    add(P.mix_b, G.mix_b, (long)M * D, l2a, 1.f, 1);
    add(P.user_mlp_w, G.user_mlp_w, (long)(p + ((c.flags & MVIN_FLAG_PS_O_FT) ? 1 : 0)) * D * D, l2w, pm, 0);   // :404
    add(P.user_mlp_b, G.user_mlp_b, D, l2w, pm, 0);
    // :405 regularises the LAST transfer matrix created (index H), :407-408 the matrices 0..h_hop: with n_mix_hop = 1 the
    // last one counts twice, with n_mix_hop > 1 the matrices h_hop+1 .. H-1 are not regularised at all
    for (int e = 0; e <= H; ++e) {
      const float mult = ((e <= Hm) ? 1.f : 0.f) + ((e == H) ? 1.f : 0.f);
      int e2 = e;
      while (e2 + 1 <= H && (((e2 + 1 <= Hm) ? 1.f : 0.f) + ((e2 + 1 == H) ? 1.f : 0.f)) == mult) ++e2;   // run of equal weights
      add(P.transfer_w + (long)e * D * D, G.transfer_w + (long)e * D * D, (long)(e2 - e + 1) * D * D, l2w, mult * pm, 0);
      add(P.transfer_b + (long)e * D, G.transfer_b + (long)e * D, (long)(e2 - e + 1) * D, l2w, mult * pm, 0);
      e = e2;
    }
    add(P.h_item_w, G.h_item_w, 2 * D, l2w, 1.f, 0);                              // :410
    add(P.h_item_b, G.h_item_b, 1, l2w, 1.f, 0);
    add(P.agg_w, G.agg_w, (long)H * D * D, l2a, am, 1);                           // :394-396
    add(P.agg_b, G.agg_b, (long)H * D, 0.f, 0.f, 1);
    add(P.agg_urh_w, G.agg_urh_w, (long)H * 3 * D, l2a, am, 1);
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

// the user side of the backward pass: user_o = O . W_user + b, the ripple attention, Q = RK^T v and the un-normalised L2 over
// the gathered memories (model.py:161-240 backward, :383-386).  du_mlp = dL/d user_o [B, D].  Leaves work on both side
// streams of `par`; the caller joins them.
template <int D>
int user_side_backward(mvin_handle_t h, const Par& par, const float* du_mlp, int B, void* ws, cudaStream_t st) {
  using C = TC<D>;
  const mvin_config_t& c = h->cfg;
  const int p = c.p_hop, m = c.n_memory, nr = c.n_relation;
  const Layout L = handle_layout(h, B);
  const mvin_params_t& P = h->P;
  const mvin_params_t& G = h->G;
  const float l2w = c.l2_weight;
  float* acc = at<float>(ws, L.acc);
  int rc;
  // user_o = O . W_user + b  backward
  {
    DwArgs a;
    memset(&a, 0, sizeof(a));
    const bool ft = (c.flags & MVIN_FLAG_PS_O_FT) != 0;   // PS_O_ft = 0: padded weights (user_side_forward), padded gradient
    float* dW_user = ft ? G.user_mlp_w : at<float>(ws, L.dWpad);
    for (int s = 0; s <= p; ++s) {
      a.A[s] = at<float>(ws, L.O) + (long)s * D; a.lda[s] = (long)(p + 1) * D; a.dW[s] = dW_user + (long)s * D * D;
    }
    a.G = du_mlp; a.db = G.user_mlp_b; a.rows = B;
    par.fork(1);                                   // side stream 1: weight gradient of the user MLP
    if (!ft) CUDA_TRY(cudaMemsetAsync(dW_user, 0, sizeof(float) * (size_t)(p + 1) * D * D, par.s(1)));
    if ((rc = launch_dw<D>(h, par.s(1), a, p + 1, "dw_user"))) return rc;
    if (!ft) {                                     // the p real blocks join the L2 term already in the gradient buffer
      cudaStream_t st = par.s(1);
      const long n4 = (long)p * D * D / 4;
      MVIN_LAUNCH((sum_rows_kernel), (unsigned)((n4 + 255) / 256), 256, 0, st, (const float*)G.user_mlp_w,
                  (const float*)(dW_user + D * D), (const float*)nullptr, n4, G.user_mlp_w);
      LAUNCH_CHECK(h, "dw_user_unpad");
    }
    // dO = du . W_user^T as a batched GEMM (a per-warp matvec inside the ripple kernel re-reads W_user per warp and
    // measured slower: +9 us at C2, +78 us at C3)
    GemmArgs g = gemm_args();
    g.A = du_mlp; g.sa_m = D; g.sa_k = 1;
    g.B = ft ? P.user_mlp_w : at<float>(ws, L.Wpad); g.sb_k = 1; g.sb_n = D;
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
    const bool tcg = use_tc_gemm<D>(h, B);
    const bool q_fused_fwd = user_q_fused(D, nr);
    if constexpr (D == 32 || D == 64) {
      if (tcg) {
        cudaStream_t st = par.s(0);
        RelGemmArgs ra;
        memset(&ra, 0, sizeof(ra));
        ra.V = at<float>(ws, L.Vbuf); ra.dQ = at<float>(ws, L.dQ); ra.dRK = G.relation_kge; ra.B = B; ra.n_rel = nr;
        if ((rc = set_smem(reldrk_tc_kernel<D>, reldrk_tc_smem<D>()))) return rc;
        MVIN_LAUNCH((reldrk_tc_kernel<D>), rel_gemm_grid(h, B, nr), RG_NT, reldrk_tc_smem<D>(), st, ra);
        LAUNCH_CHECK(h, "gemm_drk");
      }
    }
    if (!tcg && (rc = run_gemm(h, par.s(0), g, "gemm_drk"))) return rc;
    // dv[b][i] = sum_r sum_j dQ[b][r][j] RK[r][i][j]  (reduced over r inside the CTA);  dE[item_b] += dv[b]
    GemmArgs g2 = gemm_args();
    g2.A = at<float>(ws, L.dQ); g2.sa_m = (long)nr * D; g2.sa_k = 1; g2.bsA = D;
    g2.B = P.relation_kge; g2.sb_k = 1; g2.sb_n = D; g2.bsB = (long)D * D;
    g2.M = B; g2.N = D; g2.K = D; g2.nbatch = nr; g2.reduce = 1;
    g2.ksplit = nr >= 8 ? 4 : 1;
    bool dv_done = false;
    if constexpr (D == 32 || D == 64) {
      if (tcg) {
        RelGemmArgs ra;
        memset(&ra, 0, sizeof(ra));
        ra.dQ = at<float>(ws, L.dQ); ra.RK = P.relation_kge; ra.B = B; ra.n_rel = nr;
        ra.RKs = at<unsigned char>(ws, L.RKs);
        if (q_fused_fwd) {                                   // the forward built Q inside the user kernel: stage the operands now
          MVIN_LAUNCH((rel_stage_kernel<D>), dim3(nr, 2), RG_NT, 0, st, P.relation_kge, at<unsigned char>(ws, L.RKs));
          LAUNCH_CHECK(h, "rel_stage");
        }
        if (h->n_shards == 1) { ra.dE = G.entity_emb; ra.rows = at<int32_t>(ws, L.ent[0]); }
        else { ra.dE = at<float>(ws, L.dv); ra.rows = nullptr; }          // dense dv (zeroed region), scattered below
        if ((rc = set_smem(reldv_tc_kernel<D>, reldv_tc_smem<D>()))) return rc;
        MVIN_LAUNCH((reldv_tc_kernel<D>), rel_gemm_grid(h, B, nr), RG_NT, reldv_tc_smem<D>(), st, ra);
        LAUNCH_CHECK(h, "gemm_dv");
        if (h->n_shards != 1) {
          const long n = (long)B * C::LPR;
          MVIN_LAUNCH((scatter_rows_kernel<D>), (unsigned)((n + 255) / 256), 256, 0, st, at<float>(ws, L.dv), at<int32_t>(ws, L.ent[0]), B,
                      h->gtab);
          LAUNCH_CHECK(h, "scatter_dv");
        }
        dv_done = true;
      }
    }
    if (dv_done) {
    } else if (h->n_shards == 1) {
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
  return MVIN_OK;
}

template <int D>
int backward_impl(mvin_handle_t h, const float* labels, int B, float* losses_out, void* ws, cudaStream_t st) {
  using C = TC<D>;
  const mvin_config_t& c = h->cfg;
  if (c.flags & MVIN_FLAG_PS_ONLY) return backward_ps_only<D>(h, labels, B, losses_out, ws, st);
  if (generic_step(c)) return backward_mix_impl<D>(h, labels, B, losses_out, ws, st);
  const int K = c.neighbor_sample_size, H = c.h_hop, p = c.p_hop, m = c.n_memory, nr = c.n_relation;
  const Layout L = handle_layout(h, B);
  const mvin_params_t& P = h->P;
  const mvin_params_t& G = h->G;
  const float l2w = c.l2_weight, l2a = c.l2_agg_weight;
  int rc;
  float* acc = at<float>(ws, L.acc);
  float* wT = at<float>(ws, L.wT);
  const bool tcb = !L.table && use_tc_bwd(h, L, D);
  const bool gfuse = L.group && H >= 3;                      // level H - 2 of iteration 0 is never materialised (group.cuh)
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

  // Three roles of "the user vector" and where their gradients go (forward_impl):
  //   score side  u_score: user_o, or U[user] under HO_only      -> du  (written by the loss kernel)
  //   KG side     u_kg:    user_o, or U[user] when User_orient_kg_eh = 0 -> du_kg (accumulated by the KG kernels)
  //   user MLP    du_mlp = dL/d user_o: the sum of the roles user_o plays
  // Gradients of the U[user] roles are scattered into the user table (du_user).  `dukg` sits in the zeroed region.
  //   all:                 du_kg = du            du_mlp = du     du_user = none
  //   no_kg_eh_uo:         du_kg = dukg          du_mlp = du     du_user = dukg
  //   ho_only (kg_eh = 0): du_kg = du            du_mlp = dukg (stays zero: user_o is unused)   du_user = du
  //   ho_only_uo_kg_eh:    du_kg = dukg          du_mlp = dukg   du_user = du
  float* du = at<float>(ws, L.du);
  const bool kg_eh = (c.flags & MVIN_FLAG_KG_EH) != 0;
  const bool ho_only = (c.flags & MVIN_FLAG_HO_ONLY) != 0;
  const float* u_kg = kg_eh ? at<float>(ws, L.u) : at<float>(ws, L.ukg);
  const float* u_score = ho_only ? at<float>(ws, L.ukg) : at<float>(ws, L.u);
  float* du_kg = (kg_eh != ho_only) ? du : at<float>(ws, L.dukg);
  const float* du_mlp = ho_only ? at<float>(ws, L.dukg) : du;
  const float* du_user = ho_only ? du : (!kg_eh ? at<float>(ws, L.dukg) : nullptr);
  float* ditem = at<float>(ws, L.ditem);
  const float invB = 1.f / (float)(h->global_batch > 0 ? h->global_batch : B);
  const bool fused_mix_bwd = D <= 64;          // W_mix^T ((H+1) d^2 floats) is staged in shared memory
  if (fused_mix_bwd) {
    // loss gradient + mix backward in one launch: ditem, du, DC[j][0] = ditem . W_mix[j]^T
    const size_t sm = sizeof(float) * ((size_t)D * (H + 1) * D + 16 * D);
    if ((rc = set_smem(loss_mix_bwd_kernel<D>, sm))) return rc;
    const int grid = (B + 15) / 16 < 2 * h->sm_count ? (B + 15) / 16 : 2 * h->sm_count;
    MVIN_LAUNCH((loss_mix_bwd_kernel<D>), grid, 256, sm, st, at<float>(ws, L.scores), labels, u_score, at<float>(ws, L.item),
                                                  P.mix_w, B, H + 1, invB, ditem, du, at<float>(ws, L.DC[0][0]), acc);
    LAUNCH_CHECK(h, "loss_mix_bwd");
  } else {
    const long n = (long)B * C::LPR;
    MVIN_LAUNCH((loss_bwd_kernel<D>), (unsigned)((n + 255) / 256), 256, 0, st, at<float>(ws, L.scores), labels, u_score,
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
    for (int i = H - 1; i >= (L.table ? 1 : 0); --i) {
      AggBwdArgs a;
      memset(&a, 0, sizeof(a));
      long rows[MAX_LV];
      const bool tcb0 = (i == 0 && tcb);                     // the leaf level goes to agg_bwd_leaf_tc_kernel
      const int nlev = H - i - (tcb0 ? 1 : 0);
      const bool leaf_last = (i == 0 && L.entity_leaf && !tcb0);
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
        } else if (L.table && i == 1 && lv == H - 2 && L.group) {
          t.defer = 1;                                       // the children's share is evaluated per entity group below
          t.gp = at<float>(ws, L.GP);
          t.no_dself = gfuse ? 1 : 0;
        } else if (gfuse && i == 1 && lv == H - 3) {
          t.virt = 1;
          t.tab = at<float>(ws, L.Atab) + (long)(H - 2) * c.n_entity * D;
          t.Cp = at<float>(ws, L.Cp) + (long)(H - 2) * B * D;
          t.dtab = at<float>(ws, L.dA) + (long)(H - 2) * c.n_entity * D;
          t.dCs = at<float>(ws, L.dCs) + (long)(H - 2) * B * D;
        } else if (L.table && i == 1 && lv == H - 2) {
          t.virt = 1;
          t.tab = at<float>(ws, L.Atab) + (long)(H - 1) * c.n_entity * D;
          t.Cp = at<float>(ws, L.Cp) + (long)(H - 1) * B * D;
          t.dtab = at<float>(ws, L.dA) + (long)(H - 1) * c.n_entity * D;
          t.dCs = at<float>(ws, L.dCs) + (long)(H - 1) * B * D;
        } else {
          t.child = at<float>(ws, L.V[i][lv + 1]);
          t.dchild = at<float>(ws, L.DC[i][lv + 1]);
          if (tcb0 && lv == H - 2) {                         // children's share deferred to the tcgen05 leaf kernel
            t.defer = 1;
            t.gp = at<float>(ws, L.DC[0][H - 1]);            // grow of this level's rows, in the unused dchild buffer
          }
        }
      }
      a.adj = h->adj; a.s = at<float>(ws, L.s) + (long)i * nr;
      a.WaT = wT + (long)i * D * D;
      a.dWa = G.agg_w + (long)i * D * D; a.dba = G.agg_b + (long)i * D;
      a.ds = at<float>(ws, L.ds) + (long)i * nr;
      a.K = K; a.n_rel = nr;
      if (L.table && i == 1) par.join(0);   // zeroed dA / dCs are first needed here
      if (i == 0 && !tcb0) {
        par.join(0);       // zeroed dE / GSe / dQ / cnt are first needed here
        a.E = h->etab; a.WtT = wT + (long)(H + H) * D * D;
        a.dWt = G.transfer_w + (long)H * D * D; a.dbt = G.transfer_b + (long)H * D;
        a.dE = h->gtab; a.du = du_kg;
        a.GSe = L.entity_leaf ? at<float>(ws, L.GSe) : nullptr;
        a.Se = L.entity_leaf ? at<float>(ws, L.Se) : nullptr; a.u = u_kg;
        if (h->xchg.on) a.Xgsu = h->xchg.gsu[h->xchg.src_index];
        const int grid = make_tile_list(a.tl, rows, nlev, C::R,
                                        h->sm_count * resident_ctas(h, agg_bwd_kernel<D, true>, C::NT, sm_leaf), h->d_sched + 2);
        MVIN_LAUNCH((agg_bwd_kernel<D, true>), grid, C::NT, sm_leaf, st, a);
      } else {
        const bool ring = L.table && i == 1 && !L.group && RowRing<D>::ENABLED && h->ring_mode != 0;
        bool launched = false;
        if constexpr (RowRing<D>::ENABLED) {
          if (ring) {
            const size_t sm_i = sm_in + RowRing<D>::bytes();
            if ((rc = set_smem(agg_bwd_kernel<D, false, true>, sm_i))) return rc;
            a.ring = 1;
            const int grid = make_tile_list(a.tl, rows, nlev, C::R,
                                            h->sm_count * resident_ctas(h, agg_bwd_kernel<D, false, true>, C::NT, sm_i), h->d_sched + 2);
            MVIN_LAUNCH((agg_bwd_kernel<D, false, true>), grid, C::NT, sm_i, st, a);
            launched = true;
          }
        }
        if (!launched) {
          const int grid = make_tile_list(a.tl, rows, nlev, C::R,
                                          h->sm_count * resident_ctas(h, agg_bwd_kernel<D, false>, C::NT, sm_in), h->d_sched + 2);
          MVIN_LAUNCH((agg_bwd_kernel<D, false>), grid, C::NT, sm_in, st, a);
        }
      }
      LAUNCH_CHECK(h, names[i]);
      if (L.group && i == 1) {
        GroupArgs ga;
        memset(&ga, 0, sizeof(ga));
        ga.order = at<int32_t>(ws, L.gorder); ga.esort = at<int32_t>(ws, L.gesort); ga.adj = h->adj;
        ga.s = at<float>(ws, L.s) + nr;
        ga.tab = at<float>(ws, L.Atab) + (long)(H - 1) * c.n_entity * D;
        ga.Cp = at<float>(ws, L.Cp) + (long)(H - 1) * B * D;
        ga.gp = at<float>(ws, L.GP);
        ga.dtab = at<float>(ws, L.dA) + (long)(H - 1) * c.n_entity * D;
        ga.dCs = at<float>(ws, L.dCs) + (long)(H - 1) * B * D;
        ga.ds = at<float>(ws, L.ds) + nr;
        if (gfuse) {
          ga.fuse = 1;
          ga.tab_self = at<float>(ws, L.Atab) + (long)(H - 2) * c.n_entity * D;
          ga.Cp_self = at<float>(ws, L.Cp) + (long)(H - 2) * B * D;
          ga.dtab_self = at<float>(ws, L.dA) + (long)(H - 2) * c.n_entity * D;
          ga.dCs_self = at<float>(ws, L.dCs) + (long)(H - 2) * B * D;
        }
        ga.rows = L.rows[H - 2]; ga.rpp_magic = div_magic(L.rows[H - 2] / B); ga.K = K; ga.n_rel = nr;
        if ((rc = launch_group<D, true>(h, st, ga, "group_bwd"))) return rc;
      }
      if constexpr (D == 32 || D == 64) {
        if (tcb0) {
          // deepest level of iteration 0 on the tensor cores (level_tcb.cuh): dself + the parents' dchild in one buffer
          par.join(0);     // zeroed GSe / ds / du-side accumulators are first needed here
          const int lv = H - 1;
          LeafBwdArgs b;
          memset(&b, 0, sizeof(b));
          b.ent = at<int32_t>(ws, L.ent[lv]); b.ent_par = at<int32_t>(ws, L.ent[lv - 1]);
          b.adj = h->adj; b.s = at<float>(ws, L.s);
          b.g1 = at<float>(ws, L.DC[1][lv]); b.V1 = at<float>(ws, L.V[1][lv]); b.Y = at<float>(ws, L.Y[0][lv]);
          b.T = at<float>(ws, L.V[0][lv]); b.GP = at<float>(ws, L.DC[0][lv]);
          b.Se = at<float>(ws, L.Se); b.u = u_kg;
          b.Wa = P.agg_w; b.Wt = P.transfer_w + (long)H * D * D;
          b.dT = at<float>(ws, L.DS[0][lv]);
          b.dWa = G.agg_w; b.dba = G.agg_b; b.dWt = G.transfer_w + (long)H * D * D; b.dbt = G.transfer_b + (long)H * D;
          b.GSe = at<float>(ws, L.GSe); b.du = du_kg; b.ds = at<float>(ws, L.ds);
          b.rows = L.rows[lv]; b.rpp = (int)(L.rows[lv] / B); b.rpp_magic = div_magic(b.rpp);
          b.K = K; b.n_rel = nr; b.stream = stream_level(h, L.rows[lv], D);
          while ((1 << b.kshift) < K) ++b.kshift;
          const size_t smb = leaf_bwd_tc_smem<D>(nr);
          if ((rc = set_smem(agg_bwd_leaf_tc_kernel<D>, smb))) return rc;
          const long tiles = (b.rows + TB<D>::R - 1) / TB<D>::R;
          const long cap = (long)h->sm_count * resident_ctas(h, agg_bwd_leaf_tc_kernel<D>, TB<D>::NT, smb);
          MVIN_LAUNCH((agg_bwd_leaf_tc_kernel<D>), (unsigned)(tiles < cap ? tiles : cap), TB<D>::NT, smb, st, b);
          LAUNCH_CHECK(h, "agg_bwd_leaf_tc");
        }
      }
    }
  }
  // table mode: iteration 0 backward = per-entity / per-pair sums of the pre-activation gradients, then dense algebra
  if (L.table) {
    if (H == 1) par.join(0);
    const int n_virt_rows = H >= 2 ? H - 1 - (gfuse ? 1 : 0) : 1;
    {
      VirtArgs a;
      memset(&a, 0, sizeof(a));
      long end = 0;
      for (int lv = 0; lv < n_virt_rows; ++lv) {
        VirtLevel& t = a.lv[lv];
        t.ent = at<int32_t>(ws, L.ent[lv]);
        t.V = at<float>(ws, L.V[1][lv]);
        t.g1 = at<float>(ws, L.DC[1][lv]);
        t.g2 = has_agg(H, 1, lv) ? at<float>(ws, L.DS[1][lv]) : nullptr;
        t.dA = at<float>(ws, L.dA) + (long)lv * c.n_entity * D;
        t.dCs = at<float>(ws, L.dCs) + (long)lv * B * D;
        t.rpp_magic = div_magic(L.rows[lv] / B);
        end += L.rows[lv];
        a.end[lv] = end;
      }
      a.nlev = n_virt_rows;
      const long n = end * C::LPR;
      const long cap = (long)h->sm_count * 8;
      const long want = (n + 255) / 256;
      MVIN_LAUNCH((virt_rows_kernel<D, true>), (unsigned)(want < cap ? want : cap), 256, 0, st, a);
      LAUNCH_CHECK(h, "virt_rows_bwd");
    }
    {
      TableArgs a;
      memset(&a, 0, sizeof(a));
      a.stamp = at<int32_t>(ws, L.stamp); a.E = h->etab.base; a.Se = at<float>(ws, L.Se); a.M = at<float>(ws, L.Mc);
      a.dA = at<float>(ws, L.dA); a.dE = h->gtab.base; a.GSe = at<float>(ws, L.GSe); a.dM = at<float>(ws, L.dM);
      a.dc = at<float>(ws, L.dcst); a.n_entity = c.n_entity;
      const size_t sm = table_bwd_smem<D>();
      if ((rc = set_smem(table_bwd_kernel<D>, sm))) return rc;
      const long tiles = ((long)c.n_entity + C::R - 1) / C::R;
      // every CTA ends with a 2 d^2-float reduction of its weight-gradient partials into global memory: at least four
      // tiles per CTA, one resident wave over the H levels
      long cap = (long)h->sm_count * resident_ctas(h, table_bwd_kernel<D>, C::NT, sm) / H;
      if (cap < 1) cap = 1;
      long want = (tiles + 3) / 4;
      if (want < 1) want = 1;
      MVIN_LAUNCH((table_bwd_kernel<D>), dim3((unsigned)(want < cap ? want : cap), H), C::NT, sm, st, a);
      LAUNCH_CHECK(h, "table_bwd");
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
    if ((rc = launch_leaf_entity<D, true>(h, st, a, "leaf_entity_bwd"))) return rc;
  }
  if (!h->xchg.on) {       // exchange mode: ds of aggregator 0 is completed by the owners; mvin_xchg_finish_backward runs this
    MVIN_LAUNCH((rel_scores_bwd_kernel), H, 128, 0, st, P.relation_emb, P.agg_urh_w, at<float>(ws, L.ds), nr, D, G.relation_emb,
                                             G.agg_urh_w);
    LAUNCH_CHECK(h, "rel_scores_bwd");
  }
  }
  if (L.table) {
    // pair side of the tables: dMu_h = u^T dCs_h ; du += sum_h dCs_h (M1_h + M2_h)^T ; then the d x d parameter chain
    DwArgs a;
    memset(&a, 0, sizeof(a));
    for (int lv = 0; lv < H; ++lv) {
      a.A[lv] = u_kg; a.lda[lv] = D;
      a.Gg[lv] = at<float>(ws, L.dCs) + (long)lv * B * D;
      a.dW[lv] = at<float>(ws, L.dM) + (long)(lv * 3 + 2) * D * D;
    }
    a.G = a.Gg[0]; a.db = nullptr; a.rows = B;
    if ((rc = launch_dw<D>(h, st, a, H, "dw_pair"))) return rc;
    GemmArgs g = gemm_args();
    g.A = at<float>(ws, L.dCs); g.sa_m = D; g.sa_k = 1; g.bsA = (long)B * D;
    g.B = at<float>(ws, L.Mc) + (long)TBL_MSUM * D * D; g.sb_k = 1; g.sb_n = D; g.bsB = (long)TBL_NM * D * D;
    g.C = du_kg; g.ldc = D; g.bsC = 0;
    g.M = B; g.N = D; g.K = D; g.nbatch = H; g.reduce = 1; g.accumulate = 1;
    if ((rc = run_gemm(h, st, g, "gemm_du_pair"))) return rc;
    MVIN_LAUNCH((compose_bwd_kernel), dim3(H, COMPOSE_SPLIT), 256, 0, st, P.transfer_w, P.transfer_b, P.agg_w, at<float>(ws, L.dM),
                at<float>(ws, L.dcst), D, 1.f / (float)K, G.transfer_w, G.transfer_b, G.agg_w, G.agg_b);
    LAUNCH_CHECK(h, "compose_bwd");
  }
  // user-oriented transform backward, levels 0..L-1, one launch
  {
    const size_t sm = transform_bwd_smem<D>();
    if ((rc = set_smem(transform_bwd_kernel<D>, sm))) return rc;
    TransformArgs a;
    memset(&a, 0, sizeof(a));
    long rows[MAX_LV];
    const int ntl = L.table ? 1 : H;
    for (int q = 0; q < ntl; ++q) {
      const int lv = ntl - 1 - q;
      TransformLevel& t = a.lv[q];
      t.ent = at<int32_t>(ws, L.ent[lv]);
      t.W = tcb ? P.transfer_w + (long)lv * D * D : wT + (long)(H + lv) * D * D;   // the tcgen05 kernel takes W_t[lv] as stored
      t.g1 = at<float>(ws, L.DC[0][lv]); t.g2 = L.table ? nullptr : at<float>(ws, L.DS[0][lv]);
      if (tcb && lv == H - 1) { t.g1 = at<float>(ws, L.DS[0][lv]); t.g2 = nullptr; }   // one buffer: dself + dchild
      t.dW = G.transfer_w + (long)lv * D * D; t.db = G.transfer_b + (long)lv * D;
      t.rows = rows[q] = L.rows[lv]; t.rpp = (int)(L.rows[lv] / B); t.rpp_magic = div_magic(t.rpp);
        t.stream = stream_level(h, L.rows[lv], D);
    }
    a.nlev = ntl; a.E = h->etab; a.u = u_kg; a.dE = h->gtab; a.du = du_kg;
    bool done = false;
    if constexpr (D == 32 || D == 64) {
      if (tcb) {
        const size_t smt = transform_bwd_tc_smem<D>();
        if ((rc = set_smem(transform_bwd_tc_kernel<D>, smt))) return rc;
        const int grid = partition_grid(rows, H, TB<D>::R,
                                        h->sm_count * resident_ctas(h, transform_bwd_tc_kernel<D>, TB<D>::NT, smt), a.cta_end);
        MVIN_LAUNCH((transform_bwd_tc_kernel<D>), grid, TB<D>::NT, smt, st, a);
        done = true;
      }
    }
    if (!done) {
      const int grid = partition_grid(rows, ntl, C::R, h->sm_count * resident_ctas(h, transform_bwd_kernel<D>, C::NT, sm), a.cta_end);
      MVIN_LAUNCH((transform_bwd_kernel<D>), grid, C::NT, sm, st, a);
    }
    LAUNCH_CHECK(h, "transform_bwd");
  }
  if (du_user) {
    const long n = (long)B * C::LPR;
    MVIN_LAUNCH((scatter_rows_kernel<D>), (unsigned)((n + 255) / 256), 256, 0, st, du_user, at<int32_t>(ws, L.user32), B,
                GTab{G.user_emb, nullptr, 0, 0});
    LAUNCH_CHECK(h, "scatter_du_user");
  }
  if ((rc = user_side_backward<D>(h, par, du_mlp, B, ws, st))) return rc;
  par.join(0);
  par.join(1);
  MVIN_LAUNCH((finalize_loss_kernel), 1, 32, 0, st, acc, l2w, l2a, losses_out);
  LAUNCH_CHECK(h, "finalize_loss");
  return MVIN_OK;
}

// ------------------------------------------------------------------------------------------------------
// --ablation ps_only (PS_only = 1, model.py:142-144): score = user_o . E[item]; the KG side does not exist
// ------------------------------------------------------------------------------------------------------
template <int D>
int forward_ps_only(mvin_handle_t h, const int64_t* item, const int32_t* mem_h, const int32_t* mem_r, const int32_t* mem_t,
                    int B, float* scores, float* scores_norm, void* ws, cudaStream_t st) {
  using C = TC<D>;
  const Layout L = handle_layout(h, B);
  int rc;
  prof_mark(h, st, nullptr);
  MVIN_LAUNCH((seed_kernel), (unsigned)((B + 255) / 256), 256, 0, st, item, B, at<int32_t>(ws, L.ent[0]), (int32_t*)nullptr);
  LAUNCH_CHECK(h, "seed");
  if ((rc = user_side_forward<D>(h, item, mem_h, mem_r, mem_t, B, ws, st))) return rc;   // leaves Vbuf = E[item], u = user_o
  const long n = (long)B * C::LPR;
  MVIN_LAUNCH((score_kernel<D>), (unsigned)((n + 255) / 256), 256, 0, st, (const float*)at<float>(ws, L.u),
              (const float*)at<float>(ws, L.Vbuf), B, at<float>(ws, L.scores), scores_norm);
  LAUNCH_CHECK(h, "score");
  if (scores) CUDA_TRY(cudaMemcpyAsync(scores, at<float>(ws, L.scores), sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
  return MVIN_OK;
}

template <int D>
int backward_ps_only(mvin_handle_t h, const float* labels, int B, float* losses_out, void* ws, cudaStream_t st) {
  using C = TC<D>;
  const mvin_config_t& c = h->cfg;
  const Layout L = handle_layout(h, B);
  float* acc = at<float>(ws, L.acc);
  int rc;
  prof_mark(h, st, nullptr);
  const Par par{h, st, false};                               // one stream: the step is a handful of small kernels
  CUDA_TRY(cudaMemsetAsync(at<char>(ws, L.zero_begin), 0, L.zero_mid - L.zero_begin, st));
  if ((rc = backward_init<D>(h, B, ws, st, nullptr, false))) return rc;
  if (c.p_hop > 0) {
    const long n = (long)c.p_hop * B * c.n_memory;
    MVIN_LAUNCH((hist_r_kernel), h->sm_count * 4, 256, sizeof(float) * c.n_relation, st, h->mem_r, n, c.n_relation,
                at<float>(ws, L.cnt));
    LAUNCH_CHECK(h, "hist_r");
  }
  const float invB = 1.f / (float)(h->global_batch > 0 ? h->global_batch : B);
  const long n = (long)B * C::LPR;
  // ditem = g user_o -> dE[item] ;  du = g E[item] -> the user MLP
  MVIN_LAUNCH((loss_bwd_kernel<D>), (unsigned)((n + 255) / 256), 256, 0, st, (const float*)at<float>(ws, L.scores), labels,
              (const float*)at<float>(ws, L.u), (const float*)at<float>(ws, L.Vbuf), B, invB, at<float>(ws, L.ditem),
              at<float>(ws, L.du), acc);
  LAUNCH_CHECK(h, "loss_bwd");
  MVIN_LAUNCH((scatter_rows_kernel<D>), (unsigned)((n + 255) / 256), 256, 0, st, (const float*)at<float>(ws, L.ditem),
              (const int32_t*)at<int32_t>(ws, L.ent[0]), B, h->gtab);
  LAUNCH_CHECK(h, "scatter_ditem");
  if ((rc = user_side_backward<D>(h, par, at<float>(ws, L.du), B, ws, st))) return rc;
  MVIN_LAUNCH((finalize_loss_kernel), 1, 32, 0, st, acc, c.l2_weight, c.l2_agg_weight, losses_out);
  LAUNCH_CHECK(h, "finalize_loss");
  return MVIN_OK;
}

// ------------------------------------------------------------------------------------------------------
// The generic per-level step (host.cuh: generic_step): n_mix_hop > 1, User_orient = 0, User_orient_rela = 0 -- any M >= 1.
// (model.py:286-315): M mix blocks of Hm = h_hop aggregator iterations over a neighbourhood of depth
// Lt = Hm M.  Aggregator g = n Hm + i (block n, iteration i) maps levels 0 .. Lt-g-1; after the Hm iterations of block n
// the mix layer n is applied to EVERY surviving level 0 .. Lt-(n+1)Hm:
//   X[n+1][h] = [ in_n[h] ; V[n Hm + 1][h] ; ... ; V[n Hm + Hm][h] ] . W_mix[n] + b_mix[n],   in_0 = T (transform), in_n = X[n]
// and the item vector is X[M][0].  Built from the row kernels of the single-block path, one launch per (iteration,
// level), plus small GEMMs for the mix layers: a correctness path for the reference's other variants, not a tuned one.
// tests/mix_twin.py restates these loops launch by launch on the CPU (checked against the oracle's autograd).
// ------------------------------------------------------------------------------------------------------
struct MixGeom {
  int Hm, M, Lt;
  int last_level(int g) const { return Lt - g - 1; }                    // aggregator g maps levels 0 .. Lt-g-1
  int mix_levels(int n) const { return Lt - (n + 1) * Hm; }             // mix block n maps levels 0 .. this
};

// input of aggregator g (block n = g / Hm, iteration i = g % Hm) at level lv: the block's input for i = 0, else the
// previous iteration's output
inline size_t mix_in(const Layout& L, const MixGeom& q, int g, int lv) {
  const int n = g / q.Hm, i = g % q.Hm;
  if (i > 0) return L.V[g][lv];
  return n == 0 ? L.V[0][lv] : L.X[n][lv];
}
// slot k of mix block n at level lv: its input (k = 0) or the output of iteration k-1 of the block
inline size_t mix_slot(const Layout& L, const MixGeom& q, int n, int k, int lv) {
  if (k > 0) return L.V[n * q.Hm + k][lv];
  return n == 0 ? L.V[0][lv] : L.X[n][lv];
}

// a + b + c of the non-null inputs: the pointer itself when there is one, else summed into `out`
inline int sum_grads(mvin_handle_t h, cudaStream_t st, const float* a, const float* b, const float* c, long n_floats, float* out,
                     const float** res) {
  const float* v[3];
  int n = 0;
  if (a) v[n++] = a;
  if (b) v[n++] = b;
  if (c) v[n++] = c;
  if (n == 0) return fail(MVIN_ERR_STATE, "gradient without a producer");
  if (n == 1) { *res = v[0]; return MVIN_OK; }
  const long n4 = n_floats / 4;
  const long want = (n4 + 255) / 256, cap = (long)h->sm_count * 8;
  MVIN_LAUNCH((sum_rows_kernel), (unsigned)(want < cap ? want : cap), 256, 0, st, v[0], v[1], n == 3 ? v[2] : (const float*)nullptr, n4, out);
  LAUNCH_CHECK(h, "sum_grads");
  *res = out;
  return MVIN_OK;
}

template <int D>
int forward_mix_impl(mvin_handle_t h, const int64_t* item, const int32_t* mem_h, const int32_t* mem_r, const int32_t* mem_t,
                     int B, float* scores, float* scores_norm, void* ws, cudaStream_t st) {
  using C = TC<D>;
  const mvin_config_t& c = h->cfg;
  const int K = c.neighbor_sample_size, nr = c.n_relation;
  const MixGeom q{c.h_hop, c.n_mix_hop, c.h_hop * c.n_mix_hop};
  const int Hm = q.Hm, M = q.M, Lt = q.Lt;
  const Layout L = handle_layout(h, B);
  const mvin_params_t& P = h->P;
  int rc;
  prof_mark(h, st, nullptr);
  // integer expansion (model.py:243-256): ids of levels 0 .. Lt-1 (level Lt is read from the adjacency records)
  MVIN_LAUNCH((seed_kernel), (unsigned)((B + 255) / 256), 256, 0, st, item, B, at<int32_t>(ws, L.ent[0]), (int32_t*)nullptr);
  LAUNCH_CHECK(h, "seed");
  for (int lv = 0; lv + 1 < Lt; ++lv) {
    const long n = L.rows[lv] * K;
    MVIN_LAUNCH((expand_kernel), (unsigned)((n + 255) / 256), 256, 0, st, (const int32_t*)at<int32_t>(ws, L.ent[lv]), h->adj,
                L.rows[lv], K, at<int32_t>(ws, L.ent[lv + 1]), (int32_t*)nullptr, 1);
    LAUNCH_CHECK(h, "expand");
  }
  MVIN_LAUNCH((rel_scores_kernel), (Lt * nr * 32 + 255) / 256, 256, 0, st, P.relation_emb, P.agg_urh_w, nr, D, Lt, at<float>(ws, L.s));
  LAUNCH_CHECK(h, "rel_scores");
  if ((rc = user_side_forward<D>(h, item, mem_h, mem_r, mem_t, B, ws, st))) return rc;
  // the roles of the user vector: forward_impl
  const bool kg_eh = (c.flags & MVIN_FLAG_KG_EH) != 0;
  const bool ho_only = (c.flags & MVIN_FLAG_HO_ONLY) != 0;
  const float* u_kg = at<float>(ws, L.u);
  if (!kg_eh || ho_only) {
    if (!h->user) return fail(MVIN_ERR_INVALID, "user_indices are required when User_orient_kg_eh = 0 or HO_only = 1");
    const long n = (long)B * C::LPR;
    MVIN_LAUNCH((gather_user_kernel<D>), (unsigned)((n + 255) / 256), 256, 0, st, h->user, P.user_emb, B, at<float>(ws, L.ukg),
                at<int32_t>(ws, L.user32));
    LAUNCH_CHECK(h, "gather_user");
    if (!kg_eh) u_kg = at<float>(ws, L.ukg);
  }
  const float* u_score = ho_only ? at<float>(ws, L.ukg) : at<float>(ws, L.u);
  // User_orient = 0 (model.py:270): the entity vectors enter the aggregators untransformed -- the transform kernels run
  // with the identity, a zero bias and a zero user vector.  User_orient_rela = 0: the UNIFORM kernels (level.cuh)
  const bool uo = (c.flags & MVIN_FLAG_USER_ORIENT) != 0;
  const bool uniform = !(c.flags & MVIN_FLAG_USER_ORIENT_RELA);
  if (!uo) {
    MVIN_LAUNCH((eye_kernel), (D * D + 255) / 256, 256, 0, st, at<float>(ws, L.eye), D);
    LAUNCH_CHECK(h, "eye");
    CUDA_TRY(cudaMemsetAsync(at<float>(ws, L.zb), 0, sizeof(float) * D, st));
    CUDA_TRY(cudaMemsetAsync(at<float>(ws, L.zu), 0, sizeof(float) * (size_t)B * D, st));
    u_kg = at<float>(ws, L.zu);
  }
  auto Wt = [&](int e) -> const float* { return uo ? P.transfer_w + (long)e * D * D : at<float>(ws, L.eye); };
  auto bt = [&](int e) -> const float* { return uo ? P.transfer_b + (long)e * D : at<float>(ws, L.zb); };
  // user-oriented transform of levels 0 .. Lt-1   (model.py:270-283)
  {
    const size_t sm = transform_fwd_smem<D>();
    if ((rc = set_smem(transform_fwd_kernel<D>, sm))) return rc;
    for (int lv = 0; lv < Lt; ++lv) {
      TransformArgs a;
      memset(&a, 0, sizeof(a));
      long rows[MAX_LV];
      TransformLevel& t = a.lv[0];
      t.ent = at<int32_t>(ws, L.ent[lv]);
      t.W = Wt(lv); t.b = bt(lv);
      t.T = at<float>(ws, L.V[0][lv]);
      t.rows = rows[0] = L.rows[lv]; t.rpp = (int)(L.rows[lv] / B); t.rpp_magic = div_magic(t.rpp);
      t.stream = stream_level(h, L.rows[lv], D);
      a.nlev = 1; a.E = h->etab; a.u = u_kg;
      const int grid = partition_grid(rows, 1, C::R, h->sm_count * resident_ctas(h, transform_fwd_kernel<D>, C::NT, sm), a.cta_end);
      MVIN_LAUNCH((transform_fwd_kernel<D>), grid, C::NT, sm, st, a);
      LAUNCH_CHECK(h, "transform_fwd");
    }
  }
  const size_t sm_leaf = agg_fwd_smem<D, true>(K, nr), sm_in = agg_fwd_smem<D, false>(K, nr);
  for (int n = 0; n < M; ++n) {
    for (int i = 0; i < Hm; ++i) {
      const int g = n * Hm + i;
      for (int lv = q.last_level(g); lv >= 0; --lv) {
        AggArgs a;
        memset(&a, 0, sizeof(a));
        long rows[MAX_LV];
        AggLevel& t = a.lv[0];
        const bool leaf = (g == 0 && lv == Lt - 1);          // children = entity rows, transformed on the fly (W_t[Lt])
        t.ent = at<int32_t>(ws, L.ent[lv]);
        t.self = at<float>(ws, mix_in(L, q, g, lv));
        t.Y = at<float>(ws, L.Y[g][lv]); t.V = at<float>(ws, L.V[g + 1][lv]);
        t.rows = rows[0] = L.rows[lv]; t.rpp = (int)(L.rows[lv] / B); t.rpp_magic = div_magic(t.rpp);
        t.stream = stream_level(h, L.rows[lv], D);
        t.leaf = leaf ? 1 : 0;
        if (leaf) t.SU = at<float>(ws, L.SU); else t.child = at<float>(ws, mix_in(L, q, g, lv + 1));
        a.adj = h->adj; a.s = at<float>(ws, L.s) + (long)g * nr;
        a.Wa = P.agg_w + (long)g * D * D; a.ba = P.agg_b + (long)g * D;
        a.K = K; a.n_rel = nr;
        if (leaf) { a.E = h->etab; a.u = u_kg; a.Wt = Wt(Lt); a.bt = bt(Lt); }
#define MVIN_AGG_GO(LEAF_, UNI_, SM_)                                                                                    \
  do {                                                                                                                   \
    if ((rc = set_smem(agg_fwd_kernel<D, LEAF_, UNI_>, SM_))) return rc;                                                 \
    const int grid = make_tile_list(a.tl, rows, 1, C::R,                                                                 \
                                    h->sm_count * resident_ctas(h, agg_fwd_kernel<D, LEAF_, UNI_>, C::NT, SM_), h->d_sched); \
    MVIN_LAUNCH((agg_fwd_kernel<D, LEAF_, UNI_>), grid, C::NT, SM_, st, a);                                              \
  } while (0)
        if (leaf && uniform) MVIN_AGG_GO(true, true, sm_leaf);
        else if (leaf) MVIN_AGG_GO(true, false, sm_leaf);
        else if (uniform) MVIN_AGG_GO(false, true, sm_in);
        else MVIN_AGG_GO(false, false, sm_in);
#undef MVIN_AGG_GO
        LAUNCH_CHECK(h, "agg_fwd_mix");
      }
    }
    // mix layer n on every surviving level: one GEMM per input slot, accumulated into the output
    for (int lv = 0; lv <= q.mix_levels(n); ++lv) {
      float* out = (n + 1 < M) ? at<float>(ws, L.X[n + 1][lv]) : at<float>(ws, L.item);
      for (int k = 0; k <= Hm; ++k) {
        GemmArgs gm = gemm_args();
        gm.A = at<float>(ws, mix_slot(L, q, n, k, lv)); gm.sa_m = D; gm.sa_k = 1;
        gm.B = P.mix_w + ((long)n * (Hm + 1) + k) * D * D; gm.sb_k = D; gm.sb_n = 1;
        gm.C = out; gm.ldc = D;
        gm.M = (int)L.rows[lv]; gm.N = D; gm.K = D;
        gm.bias = k == 0 ? P.mix_b + (long)n * D : nullptr;
        gm.accumulate = k > 0;
        if ((rc = run_gemm(h, st, gm, "gemm_mix"))) return rc;
      }
    }
  }
  {
    const long n = (long)B * C::LPR;
    MVIN_LAUNCH((score_kernel<D>), (unsigned)((n + 255) / 256), 256, 0, st, u_score, (const float*)at<float>(ws, L.item), B,
                at<float>(ws, L.scores), scores_norm);
    LAUNCH_CHECK(h, "score");
    if (scores) CUDA_TRY(cudaMemcpyAsync(scores, at<float>(ws, L.scores), sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
  }
  return MVIN_OK;
}

template <int D>
int backward_mix_impl(mvin_handle_t h, const float* labels, int B, float* losses_out, void* ws, cudaStream_t st) {
  using C = TC<D>;
  const mvin_config_t& c = h->cfg;
  const int K = c.neighbor_sample_size, p = c.p_hop, m = c.n_memory, nr = c.n_relation;
  const MixGeom q{c.h_hop, c.n_mix_hop, c.h_hop * c.n_mix_hop};
  const int Hm = q.Hm, M = q.M, Lt = q.Lt;
  const Layout L = handle_layout(h, B);
  const mvin_params_t& P = h->P;
  const mvin_params_t& G = h->G;
  float* acc = at<float>(ws, L.acc);
  float* wT = at<float>(ws, L.wT);                           // wT[g] = W_a[g]^T (g < Lt), wT[Lt + e] = W_t[e]^T (e <= Lt)
  int rc;
  prof_mark(h, st, nullptr);
  const Par par{h, st, false};                               // one stream
  CUDA_TRY(cudaMemsetAsync(at<char>(ws, L.zero_begin), 0, L.zero_mid - L.zero_begin, st));
  if ((rc = backward_init<D>(h, B, ws, st, nullptr, false))) return rc;
  if (p > 0) {
    const long n = (long)p * B * m;
    MVIN_LAUNCH((hist_r_kernel), h->sm_count * 4, 256, sizeof(float) * nr, st, h->mem_r, n, nr, at<float>(ws, L.cnt));
    LAUNCH_CHECK(h, "hist_r");
  }
  // the roles of the user vector and their gradients: backward_impl
  float* du = at<float>(ws, L.du);
  const bool kg_eh = (c.flags & MVIN_FLAG_KG_EH) != 0;
  const bool ho_only = (c.flags & MVIN_FLAG_HO_ONLY) != 0;
  const float* u_kg = kg_eh ? at<float>(ws, L.u) : at<float>(ws, L.ukg);
  const float* u_score = ho_only ? at<float>(ws, L.ukg) : at<float>(ws, L.u);
  float* du_kg = (kg_eh != ho_only) ? du : at<float>(ws, L.dukg);
  const float* du_mlp = ho_only ? at<float>(ws, L.dukg) : du;
  const float* du_user = ho_only ? du : (!kg_eh ? at<float>(ws, L.dukg) : nullptr);
  float* ditem = at<float>(ws, L.ditem);
  const float invB = 1.f / (float)(h->global_batch > 0 ? h->global_batch : B);
  {
    const long n = (long)B * C::LPR;
    MVIN_LAUNCH((loss_bwd_kernel<D>), (unsigned)((n + 255) / 256), 256, 0, st, (const float*)at<float>(ws, L.scores), labels, u_score,
                (const float*)at<float>(ws, L.item), B, invB, ditem, du, acc);
    LAUNCH_CHECK(h, "loss_bwd");
  }
  const size_t sm_leaf = agg_bwd_smem<D, true>(K, nr), sm_in = agg_bwd_smem<D, false>(K, nr);
  // User_orient = 0: identity transforms (forward_mix_impl); what the kernels accumulate for W_t, b_t and the user vector
  // lands in scratch.  User_orient_rela = 0: the UNIFORM kernels
  const bool uo = (c.flags & MVIN_FLAG_USER_ORIENT) != 0;
  const bool uniform = !(c.flags & MVIN_FLAG_USER_ORIENT_RELA);
  if (!uo) {
    CUDA_TRY(cudaMemsetAsync(at<float>(ws, L.sdW), 0, sizeof(float) * D * D, st));
    CUDA_TRY(cudaMemsetAsync(at<float>(ws, L.sdb), 0, sizeof(float) * D, st));
    CUDA_TRY(cudaMemsetAsync(at<float>(ws, L.sdu), 0, sizeof(float) * (size_t)B * D, st));
    u_kg = at<float>(ws, L.zu);
    du_kg = at<float>(ws, L.sdu);
  }
  auto WtT = [&](int e) -> const float* { return uo ? wT + (long)(Lt + e) * D * D : at<float>(ws, L.eye); };
  auto dWt = [&](int e) -> float* { return uo ? G.transfer_w + (long)e * D * D : at<float>(ws, L.sdW); };
  auto dbt = [&](int e) -> float* { return uo ? G.transfer_b + (long)e * D : at<float>(ws, L.sdb); };
  const float* gx_next[MAX_L] = {nullptr};                   // gradient of X[n + 1][lv] (n + 1 < M), per level
  for (int n = M - 1; n >= 0; --n) {
    // mix layer n backward on every level it maps
    for (int lv = 0; lv <= q.mix_levels(n); ++lv) {
      const float* dout = (n + 1 < M) ? gx_next[lv] : ditem;
      const long rows = L.rows[lv];
      GemmArgs gm = gemm_args();                             // DM[k] = dout . W_mix[n][k]^T, k = 0 .. Hm
      gm.A = dout; gm.sa_m = D; gm.sa_k = 1; gm.bsA = 0;
      gm.B = P.mix_w + (long)n * (Hm + 1) * D * D; gm.sb_k = 1; gm.sb_n = D; gm.bsB = (long)D * D;
      gm.C = at<float>(ws, L.DM[n][lv]); gm.ldc = D; gm.bsC = rows * D;
      gm.M = (int)rows; gm.N = D; gm.K = D; gm.nbatch = Hm + 1;
      if ((rc = run_gemm(h, st, gm, "gemm_mix_bwd"))) return rc;
      DwArgs a;                                              // dW_mix[n][k] += slot_k^T dout ; db_mix[n] += sum dout
      memset(&a, 0, sizeof(a));
      for (int k = 0; k <= Hm; ++k) {
        a.A[k] = at<float>(ws, mix_slot(L, q, n, k, lv)); a.lda[k] = D;
        a.dW[k] = G.mix_w + ((long)n * (Hm + 1) + k) * D * D;
      }
      a.G = dout; a.db = G.mix_b + (long)n * D; a.rows = rows;
      if ((rc = launch_dw<D>(h, st, a, Hm + 1, "dw_mix"))) return rc;
    }
    // the block's aggregator iterations, reversed
    for (int i = Hm - 1; i >= 0; --i) {
      const int g = n * Hm + i;
      for (int lv = 0; lv <= q.last_level(g); ++lv) {
        const long rows_lv = L.rows[lv];
        // gradient of V[g + 1][lv]: mix slot i + 1, and inside the block its own next step and its parent's children sum
        const float* s_mix = lv <= q.mix_levels(n) ? at<float>(ws, L.DM[n][lv]) + (long)(i + 1) * rows_lv * D : nullptr;
        const float* s_self = (i + 1 < Hm && lv <= q.last_level(g + 1)) ? at<float>(ws, L.DS[g + 1][lv]) : nullptr;
        const float* s_par = (i + 1 < Hm && lv >= 1) ? at<float>(ws, L.DC[g + 1][lv]) : nullptr;
        const float* gsum;
        if ((rc = sum_grads(h, st, s_mix, s_self, s_par, rows_lv * D, at<float>(ws, L.GV[g + 1][lv]), &gsum))) return rc;
        AggBwdArgs a;
        memset(&a, 0, sizeof(a));
        long rows[MAX_LV];
        AggBwdLevel& t = a.lv[0];
        const bool leaf = (g == 0 && lv == Lt - 1);
        t.ent = at<int32_t>(ws, L.ent[lv]);
        t.V = at<float>(ws, L.V[g + 1][lv]); t.Y = at<float>(ws, L.Y[g][lv]);
        t.g1 = gsum; t.g2 = nullptr;
        t.dself = at<float>(ws, L.DS[g][lv]);
        t.rows = rows[0] = rows_lv; t.rpp = (int)(rows_lv / B); t.rpp_magic = div_magic(t.rpp);
        t.stream = stream_level(h, rows_lv, D);
        t.leaf = leaf ? 1 : 0;
        if (leaf) {
          t.SU = at<float>(ws, L.SU);
        } else {
          t.child = at<float>(ws, mix_in(L, q, g, lv + 1));
          t.dchild = at<float>(ws, L.DC[g][lv + 1]);
        }
        a.adj = h->adj; a.s = at<float>(ws, L.s) + (long)g * nr;
        a.WaT = wT + (long)g * D * D;
        a.dWa = G.agg_w + (long)g * D * D; a.dba = G.agg_b + (long)g * D;
        a.ds = at<float>(ws, L.ds) + (long)g * nr;
        a.K = K; a.n_rel = nr;
        if (leaf) {
          a.E = h->etab; a.WtT = WtT(Lt); a.dWt = dWt(Lt); a.dbt = dbt(Lt);
          a.dE = h->gtab; a.du = du_kg; a.u = u_kg;
        }
#define MVIN_AGG_GO(LEAF_, UNI_, SM_)                                                                                    \
  do {                                                                                                                   \
    if ((rc = set_smem(agg_bwd_kernel<D, LEAF_, false, UNI_>, SM_))) return rc;                                          \
    const int grid = make_tile_list(a.tl, rows, 1, C::R,                                                                 \
                                    h->sm_count * resident_ctas(h, agg_bwd_kernel<D, LEAF_, false, UNI_>, C::NT, SM_),   \
                                    h->d_sched + 2);                                                                     \
    MVIN_LAUNCH((agg_bwd_kernel<D, LEAF_, false, UNI_>), grid, C::NT, SM_, st, a);                                       \
  } while (0)
        if (leaf && uniform) MVIN_AGG_GO(true, true, sm_leaf);
        else if (leaf) MVIN_AGG_GO(true, false, sm_leaf);
        else if (uniform) MVIN_AGG_GO(false, true, sm_in);
        else MVIN_AGG_GO(false, false, sm_in);
#undef MVIN_AGG_GO
        LAUNCH_CHECK(h, "agg_bwd_mix");
      }
    }
    // gradient of the block's inputs: mix slot 0, the first iteration's own step and its parents' children sums
    const int g0 = n * Hm;
    const int n_in = n == 0 ? Lt - 1 : Lt - g0;              // deepest input level that is a row buffer
    for (int lv = 0; lv <= n_in; ++lv) {
      const long rows_lv = L.rows[lv];
      const float* s_mix = lv <= q.mix_levels(n) ? at<float>(ws, L.DM[n][lv]) : nullptr;
      const float* s_self = lv <= q.last_level(g0) ? at<float>(ws, L.DS[g0][lv]) : nullptr;
      const float* s_par = lv >= 1 ? at<float>(ws, L.DC[g0][lv]) : nullptr;
      float* out = n == 0 ? at<float>(ws, L.GT[lv]) : at<float>(ws, L.GX[n][lv]);
      if ((rc = sum_grads(h, st, s_mix, s_self, s_par, rows_lv * D, out, &gx_next[lv]))) return rc;
    }
  }
  // user-oriented transform backward, levels 0 .. Lt-1 (gx_next = gradient of T[lv])
  {
    const size_t sm = transform_bwd_smem<D>();
    if ((rc = set_smem(transform_bwd_kernel<D>, sm))) return rc;
    for (int lv = 0; lv < Lt; ++lv) {
      TransformArgs a;
      memset(&a, 0, sizeof(a));
      long rows[MAX_LV];
      TransformLevel& t = a.lv[0];
      t.ent = at<int32_t>(ws, L.ent[lv]);
      t.W = WtT(lv);
      t.g1 = gx_next[lv]; t.g2 = nullptr;
      t.dW = dWt(lv); t.db = dbt(lv);
      t.rows = rows[0] = L.rows[lv]; t.rpp = (int)(L.rows[lv] / B); t.rpp_magic = div_magic(t.rpp);
      t.stream = stream_level(h, L.rows[lv], D);
      a.nlev = 1; a.E = h->etab; a.u = u_kg; a.dE = h->gtab; a.du = du_kg;
      const int grid = partition_grid(rows, 1, C::R, h->sm_count * resident_ctas(h, transform_bwd_kernel<D>, C::NT, sm), a.cta_end);
      MVIN_LAUNCH((transform_bwd_kernel<D>), grid, C::NT, sm, st, a);
      LAUNCH_CHECK(h, "transform_bwd");
    }
  }
  MVIN_LAUNCH((rel_scores_bwd_kernel), Lt, 128, 0, st, P.relation_emb, P.agg_urh_w, (const float*)at<float>(ws, L.ds), nr, D,
              G.relation_emb, G.agg_urh_w);
  LAUNCH_CHECK(h, "rel_scores_bwd");
  if (du_user) {
    const long n = (long)B * C::LPR;
    MVIN_LAUNCH((scatter_rows_kernel<D>), (unsigned)((n + 255) / 256), 256, 0, st, du_user, (const int32_t*)at<int32_t>(ws, L.user32), B,
                GTab{G.user_emb, nullptr, 0, 0});
    LAUNCH_CHECK(h, "scatter_du_user");
  }
  if ((rc = user_side_backward<D>(h, par, du_mlp, B, ws, st))) return rc;
  MVIN_LAUNCH((finalize_loss_kernel), 1, 32, 0, st, acc, c.l2_weight, c.l2_agg_weight, losses_out);
  LAUNCH_CHECK(h, "finalize_loss");
  return MVIN_OK;
}

// ------------------------------------------------------------------------------------------------------
// exchange mode of the row-sharded entity table (exchange.cuh): the phases around forward_impl / backward_impl
// ------------------------------------------------------------------------------------------------------
// ids of the leaf-level parent nodes (level H-1) of this rank's batch + the relation scores the owners need
template <int D>
int xchg_expand_impl(mvin_handle_t h, const int64_t* item, int B, int32_t* ids_out, void* ws, cudaStream_t st) {
  const mvin_config_t& c = h->cfg;
  const int K = c.neighbor_sample_size, H = c.h_hop, nr = c.n_relation;
  const Layout L = handle_layout(h, B);
  MVIN_LAUNCH((seed_kernel), (unsigned)((B + 255) / 256), 256, 0, st, item, B, H == 1 ? ids_out : at<int32_t>(ws, L.ent[0]), nullptr);
  LAUNCH_CHECK(h, "seed");
  for (int lv = 0; lv + 1 < H; ++lv) {
    const long n = L.rows[lv] * K;
    MVIN_LAUNCH((expand_kernel), (unsigned)((n + 255) / 256), 256, 0, st, at<int32_t>(ws, L.ent[lv]), h->adj, L.rows[lv], K,
                lv + 1 == H - 1 ? ids_out : at<int32_t>(ws, L.ent[lv + 1]), nullptr, 1);
    LAUNCH_CHECK(h, "expand");
  }
  MVIN_LAUNCH((rel_scores_kernel), (H * nr * 32 + 255) / 256, 256, 0, st, h->P.relation_emb, h->P.agg_urh_w, nr, D, H, at<float>(ws, L.s));
  LAUNCH_CHECK(h, "rel_scores");
  return MVIN_OK;
}

template <int D>
int xchg_owner_impl(mvin_handle_t h, int owner, bool bwd, int B, void* ws, cudaStream_t st) {
  const mvin_config_t& c = h->cfg;
  const Layout L = handle_layout(h, B);
  XchgArgs a;
  memset(&a, 0, sizeof(a));
  a.ids = h->xchg.ids_all; a.adj = h->adj; a.s = at<float>(ws, L.s);
  a.E = static_cast<const float*>(h->shard_host[owner]);
  a.dE = static_cast<float*>(h->shard_host[MAX_SHARDS + owner]);
  for (int s = 0; s < h->xchg.n_src; ++s) { a.part[s] = h->xchg.part[s]; a.gsu[s] = h->xchg.gsu[s]; a.dot[s] = h->xchg.dot[s]; }
  a.ds = at<float>(ws, L.ds);
  a.rows = h->xchg.rows; a.n_src = h->xchg.n_src; a.owner = owner; a.shift = h->etab.shift; a.mask = h->etab.mask;
  a.K = c.neighbor_sample_size; a.n_rel = c.n_relation;
  const size_t sm = xchg_smem(c.n_relation, bwd);
  const long want = ((long)a.n_src * a.rows + XCHG_NW - 1) / XCHG_NW;
  const long cap = (long)h->sm_count * 8;
  int rc;
  if (bwd) {
    if ((rc = set_smem(xchg_owner_bwd_kernel<D>, sm))) return rc;
    MVIN_LAUNCH((xchg_owner_bwd_kernel<D>), (unsigned)(want < cap ? want : cap), XCHG_NT, sm, st, a);
    LAUNCH_CHECK(h, "xchg_owner_bwd");
  } else {
    if ((rc = set_smem(xchg_owner_fwd_kernel<D>, sm))) return rc;
    MVIN_LAUNCH((xchg_owner_fwd_kernel<D>), (unsigned)(want < cap ? want : cap), XCHG_NT, sm, st, a);
    LAUNCH_CHECK(h, "xchg_owner_fwd");
  }
  return MVIN_OK;
}

// source side: close the softmax gradient of the leaf level with the owners' partial dots, then the relation-score
// backward that backward_impl deferred
template <int D>
int xchg_finish_impl(mvin_handle_t h, int B, void* ws, cudaStream_t st) {
  const mvin_config_t& c = h->cfg;
  const Layout L = handle_layout(h, B);
  XchgFinishArgs a;
  memset(&a, 0, sizeof(a));
  a.ids = h->xchg.ids_all + (long)h->xchg.src_index * h->xchg.rows; a.adj = h->adj; a.s = at<float>(ws, L.s);
  a.dot = h->xchg.dot[h->xchg.src_index]; a.ds = at<float>(ws, L.ds);
  a.rows = h->xchg.rows; a.G = h->n_shards; a.K = c.neighbor_sample_size; a.n_rel = c.n_relation;
  const size_t sm = sizeof(float) * (size_t)c.n_relation * (1 + leaf_ds_copies(c.n_relation));
  const long want = (a.rows + XCHG_NW - 1) / XCHG_NW;
  const long cap = (long)h->sm_count * 8;
  MVIN_LAUNCH((xchg_finish_bwd_kernel), (unsigned)(want < cap ? want : cap), XCHG_NT, sm, st, a);
  LAUNCH_CHECK(h, "xchg_finish_bwd");
  MVIN_LAUNCH((rel_scores_bwd_kernel), c.h_hop, 128, 0, st, h->P.relation_emb, h->P.agg_urh_w, at<float>(ws, L.ds), c.n_relation, D,
              h->G.relation_emb, h->G.agg_urh_w);
  LAUNCH_CHECK(h, "rel_scores_bwd");
  return MVIN_OK;
}

}  // namespace mvin_host
