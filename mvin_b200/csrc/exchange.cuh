// exchange.cuh -- owner-side partial reduction of the deepest level for the ROW-SHARDED entity table (BASELINE.json C5,
// SURVEY.md 8(e)): instead of pulling every raw leaf row over NVLink, each rank reduces the rows it OWNS and ships one
// d-vector per (parent node, owner).
//
// Reference (src/model/MVIN/): the leaf gather model.py:267 is consumed only through
// reduce_mean(probs * neighbor_vectors) (aggregators.py:141-144), which is linear in the rows:
//     S_j = sum_k p_k E[n_jk] = sum_g  S_j^(g),      S_j^(g) = sum_{k : owner(n_jk) = g} p_k E_g[n_jk div G]
// The adjacency and the relation scores are replicated on every rank, so an owner can re-derive the children ids and
// the attention p_k of a parent from the parent's ENTITY id alone: "routing the neighbour indices to the owning shard"
// is an all-gather of 4 bytes per parent node (one NCCL collective), and the return exchange -- one 4d-byte partial per
// (parent, owner) instead of K/G raw rows -- is fused into the owner's kernel as coalesced stores into the source
// rank's receive buffer through its CUDA-IPC peer mapping (NVLink / NVSwitch).  Volume per rank at C5 on 8 GPUs:
// 1.9 GB of partials instead of 15 GB of rows, in each direction.
//
// Backward (TF autodiff of the same ops): the source leaves gsu_j = dL/dS_j (one row per parent) in a buffer its peers
// read over NVLink; the owner scatter-adds p_k gsu_j into its LOCAL gradient shard (local L2 atomics, no peer
// reductions), evaluates dp_k = gsu_j . E[n_jk] for its children, adds p_k dp_k to its relation-score gradient and
// returns one scalar per (parent, owner), dot_j^(g) = sum_{k in g} p_k dp_k; the source closes the softmax gradient
// dlogit_k = p_k (dp_k - sum_g dot_j^(g)) with ds[rel_k] -= p_k sum_g dot_j^(g).
//
// Cross-rank ordering between the phases is the caller's: one stream-ordered NCCL collective between them
// (mvin_b200/model.py).
#pragma once
#include "level.cuh"

namespace mvin {

constexpr int XCHG_NT = 256, XCHG_NW = XCHG_NT / 32, XCHG_MAX_RANKS = 16;

struct XchgArgs {
  const int32_t* ids;      // [n_src][rows] entity id of every parent node of every source rank (all-gathered)
  const int32_t* adj;      // packed adjacency (replicated)
  const float* s;          // [n_rel] relation scores of aggregator 0
  const float* E;          // this owner's entity shard [n_local_rows, D]
  float* dE;               // this owner's gradient shard
  float* part[XCHG_MAX_RANKS];        // fwd: part[s] = source s's receive buffer [G][rows][D]
  const float* gsu[XCHG_MAX_RANKS];   // bwd: gsu[s]  = source s's gsu buffer [rows][D]
  float* dot[XCHG_MAX_RANKS];         // bwd: dot[s]  = source s's buffer [G][rows]
  float* ds;               // bwd: [n_rel] relation-score gradient of aggregator 0 on this rank (+=)
  long rows;               // parent nodes per source rank
  int n_src, owner, shift, mask, K, n_rel;
};

inline size_t xchg_smem(int n_rel, bool bwd) {
  return sizeof(float) * (2 * XCHG_NW * MAX_K + (bwd ? (size_t)n_rel * (1 + leaf_ds_copies(n_rel)) : (size_t)n_rel));
}

// softmax over the K slots of an adjacency record held as (lane, lane + 32)
MVIN_DEV void rec_softmax(const AdjRec& rec, const float* __restrict__ s_s, int K, int lane, float& p0, float& p1) {
  const float l0 = lane < K ? s_s[rec.rel0] : -INFINITY;
  const float l1 = lane + 32 < K ? s_s[rec.rel1] : -INFINITY;
  const float mx = warp_max(fmaxf(l0, l1));
  const float e0 = lane < K ? expf(l0 - mx) : 0.f;
  const float e1 = lane + 32 < K ? expf(l1 - mx) : 0.f;
  const float inv = 1.f / warp_sum(e0 + e1);
  p0 = e0 * inv;
  p1 = e1 * inv;
}

// children of the record owned by `owner`, compacted into the warp's shared-memory list (p, local row[, rel]); returns
// their number
MVIN_DEV int compact_owned(const AdjRec& rec, float p0, float p1, int K, int lane, int owner, int shift, int mask,
                           float* __restrict__ pw, int* __restrict__ rw, int* __restrict__ relw) {
  const bool m0 = lane < K && (rec.id0 & mask) == owner;
  const bool m1 = lane + 32 < K && (rec.id1 & mask) == owner;
  const unsigned b0 = __ballot_sync(FULL_MASK, m0), b1 = __ballot_sync(FULL_MASK, m1);
  const unsigned below = (1u << lane) - 1u;
  const int n0 = __popc(b0);
  if (m0) {
    const int q = __popc(b0 & below);
    pw[q] = p0; rw[q] = rec.id0 >> shift;
    if (relw) relw[q] = rec.rel0;
  }
  if (m1) {
    const int q = n0 + __popc(b1 & below);
    pw[q] = p1; rw[q] = rec.id1 >> shift;
    if (relw) relw[q] = rec.rel1;
  }
  __syncwarp();
  return n0 + __popc(b1);
}

// forward: part[s][owner][j] = sum_{k owned} p_k E[n_jk]  for every parent j of every source s.  One warp per parent.
template <int D>
__global__ void __launch_bounds__(XCHG_NT) xchg_owner_fwd_kernel(XchgArgs a) {
  pdl_enter();
  constexpr int LPR = D / 4, G = 32 / LPR;
  extern __shared__ __align__(16) float smem[];
  float* pw = smem;                                        // [NW][MAX_K]
  int* rw = reinterpret_cast<int*>(pw + XCHG_NW * MAX_K);  // [NW][MAX_K]
  float* s_s = reinterpret_cast<float*>(rw + XCHG_NW * MAX_K);
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, g = lane / LPR, c = lane % LPR;
  for (int i = tid; i < a.n_rel; i += XCHG_NT) s_s[i] = a.s[i];
  __syncthreads();
  float* pw_w = pw + warp * MAX_K;
  int* rw_w = rw + warp * MAX_K;
  const long total = (long)a.n_src * a.rows;
  const int K = a.K;
  long idx = (long)blockIdx.x * XCHG_NW + warp;
  const long stride = (long)gridDim.x * XCHG_NW;
  AdjRec nxt{0, 0, 0, 0};
  if (idx < total) nxt = load_adj(a.adj, __ldg(a.ids + idx), K, lane);
  for (; idx < total; idx += stride) {
    const AdjRec rec = nxt;
    if (idx + stride < total) nxt = load_adj(a.adj, __ldg(a.ids + idx + stride), K, lane);
    float p0, p1;
    rec_softmax(rec, s_s, K, lane, p0, p1);
    const int cnt = compact_owned(rec, p0, p1, K, lane, a.owner, a.shift, a.mask, pw_w, rw_w, nullptr);
    float4 acc = f4zero();
#pragma unroll 4
    for (int k = g; k < cnt; k += G) acc = f4fma(pw_w[k], ldg4(a.E + (long)rw_w[k] * D + c * 4), acc);
    acc = cross_group_sum4<LPR>(acc);
    const int src = (int)(idx / a.rows);
    const long j = idx - (long)src * a.rows;
    if (g == 0) st4(a.part[src] + ((long)a.owner * a.rows + j) * D + c * 4, acc);
    __syncwarp();
  }
  __threadfence_system();                                  // the partials are read by the source rank after the next collective
}

// backward: for every parent j of every source s and every child k this owner holds:
//   dE[n_jk] += p_k gsu_j ;  dp_k = gsu_j . E[n_jk] ;  ds[rel_k] += p_k dp_k ;  dot[s][owner][j] = sum_k p_k dp_k
template <int D>
__global__ void __launch_bounds__(XCHG_NT) xchg_owner_bwd_kernel(XchgArgs a) {
  pdl_enter();
  constexpr int LPR = D / 4, G = 32 / LPR;
  extern __shared__ __align__(16) float smem[];
  float* pw = smem;
  int* rw = reinterpret_cast<int*>(pw + XCHG_NW * MAX_K);
  float* s_s = reinterpret_cast<float*>(rw + XCHG_NW * MAX_K);
  float* ds_s = s_s + a.n_rel;
  const int NWH = leaf_ds_copies(a.n_rel);
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, g = lane / LPR, c = lane % LPR;
  for (int i = tid; i < a.n_rel; i += XCHG_NT) s_s[i] = a.s[i];
  for (int i = tid; i < NWH * a.n_rel; i += XCHG_NT) ds_s[i] = 0.f;
  __syncthreads();
  float* pw_w = pw + warp * MAX_K;
  int* rw_w = rw + warp * MAX_K;
  float* ds_w = ds_s + (NWH > 1 ? warp : 0) * a.n_rel;
  const long total = (long)a.n_src * a.rows;
  const int K = a.K;
  long idx = (long)blockIdx.x * XCHG_NW + warp;
  const long stride = (long)gridDim.x * XCHG_NW;
  AdjRec nxt{0, 0, 0, 0};
  if (idx < total) nxt = load_adj(a.adj, __ldg(a.ids + idx), K, lane);
  for (; idx < total; idx += stride) {
    const AdjRec rec = nxt;
    if (idx + stride < total) nxt = load_adj(a.adj, __ldg(a.ids + idx + stride), K, lane);
    const int src = (int)(idx / a.rows);
    const long j = idx - (long)src * a.rows;
    const float4 gr = ld4(a.gsu[src] + j * D + c * 4);      // NVLink peer load when src is another rank
    float p0, p1;
    rec_softmax(rec, s_s, K, lane, p0, p1);
    const bool m0 = lane < K && (rec.id0 & a.mask) == a.owner;
    const bool m1 = lane + 32 < K && (rec.id1 & a.mask) == a.owner;
    const int cnt = compact_owned(rec, p0, p1, K, lane, a.owner, a.shift, a.mask, pw_w, rw_w, nullptr);
    float dot = 0.f;                                       // sum over the owned children of p_k dp_k (complete on every lane)
    for (int k0 = 0; k0 < cnt; k0 += G) {
      const int k = k0 + g;
      float part = 0.f;
      float pk = 0.f;
      if (k < cnt) {
        const long row = rw_w[k];
        pk = pw_w[k];
        part = f4dot(gr, ldg4(a.E + row * D + c * 4));
        red_add4(a.dE + row * D + c * 4, f4scale(gr, pk));
      }
      part = group_sum<LPR>(part);                          // dp_k on every lane of group g
      // overwrite p with p_k dp_k for the histogram pass below (lane c == 0 of each group)
      if (k < cnt && c == 0) pw_w[k] = pk * part;
      float contrib = (k < cnt && c == 0) ? pk * part : 0.f;
      dot += warp_sum(contrib);
    }
    __syncwarp();
    // ds[rel_k] += p_k dp_k : lane l owns slots l, l + 32; its position in the compacted list is the ballot rank
    {
      const unsigned b0 = __ballot_sync(FULL_MASK, m0), b1 = __ballot_sync(FULL_MASK, m1);
      const unsigned below = (1u << lane) - 1u;
      const int n0 = __popc(b0);
      if (m0) atomicAdd(&ds_w[rec.rel0], pw_w[__popc(b0 & below)]);
      if (m1) atomicAdd(&ds_w[rec.rel1], pw_w[n0 + __popc(b1 & below)]);
    }
    if (lane == 0) a.dot[src][(long)a.owner * a.rows + j] = dot;
    __syncwarp();
  }
  __syncthreads();
  for (int i = tid; i < a.n_rel; i += XCHG_NT) {
    float s = 0.f;
    for (int w = 0; w < NWH; ++w) s += ds_s[w * a.n_rel + i];
    if (s != 0.f) atomicAdd(a.ds + i, s);
  }
  __threadfence_system();
}

// source side, after the owners: dot_j = sum_g dot[g][j];  ds[rel_k] -= p_k dot_j  over the K slots of parent j
struct XchgFinishArgs {
  const int32_t* ids;      // [rows] this rank's parent entities
  const int32_t* adj;
  const float* s;
  const float* dot;        // [G][rows]
  float* ds;               // [n_rel] (+=)
  long rows;
  int G, K, n_rel;
};
static __global__ void __launch_bounds__(XCHG_NT) xchg_finish_bwd_kernel(XchgFinishArgs a) {
  pdl_enter();
  extern __shared__ __align__(16) float smem[];
  float* s_s = smem;
  float* ds_s = s_s + a.n_rel;
  const int NWH = leaf_ds_copies(a.n_rel);
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  for (int i = tid; i < a.n_rel; i += XCHG_NT) s_s[i] = a.s[i];
  for (int i = tid; i < NWH * a.n_rel; i += XCHG_NT) ds_s[i] = 0.f;
  __syncthreads();
  float* ds_w = ds_s + (NWH > 1 ? warp : 0) * a.n_rel;
  const int K = a.K;
  for (long j = (long)blockIdx.x * XCHG_NW + warp; j < a.rows; j += (long)gridDim.x * XCHG_NW) {
    const AdjRec rec = load_adj(a.adj, __ldg(a.ids + j), K, lane);
    float d = lane < a.G ? a.dot[(long)lane * a.rows + j] : 0.f;
    d = warp_sum(d);
    float p0, p1;
    rec_softmax(rec, s_s, K, lane, p0, p1);
    if (lane < K) atomicAdd(&ds_w[rec.rel0], -p0 * d);
    if (lane + 32 < K) atomicAdd(&ds_w[rec.rel1], -p1 * d);
  }
  __syncthreads();
  for (int i = tid; i < a.n_rel; i += XCHG_NT) {
    float s = 0.f;
    for (int w = 0; w < NWH; ++w) s += ds_s[w * a.n_rel + i];
    if (s != 0.f) atomicAdd(a.ds + i, s);
  }
}

}  // namespace mvin
