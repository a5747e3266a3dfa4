// misc.cuh -- integer neighbour expansion (get_neighbors), score / loss, dense L2 terms, Adam and other small
// kernels of the MVIN path.  Reference citations are relative to src/model/MVIN/.
#pragma once
#include "common.cuh"

namespace mvin {

// ---- adjacency packing: int64 [n_entity, K] x 2 (model.py:7,19-20) -> int32 [n_entity][2][K] -----------
static __global__ void pack_adj_kernel(const int64_t* __restrict__ adjE, const int64_t* __restrict__ adjR, long n_entity,
                                int K, int32_t* __restrict__ out) {
  pdl_enter();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_entity * K) return;
  const long e = i / K;
  const int k = (int)(i % K);
  out[e * 2 * K + k] = (int32_t)adjE[i];
  out[e * 2 * K + K + k] = (int32_t)adjR[i];
}

// ---- get_neighbors (model.py:243-256): one level of expansion, child k of node j at j*K+k -------------
// `stamp` (optional): mark the produced ids (entity mode of the leaf level, level.cuh); `out` may be null (table mode:
// the deepest level is only stamped, its ids are re-read from the adjacency records)
// `bit`: the stamp is a mask of the levels the entity occurs at (table mode: bit h = level h); 1 elsewhere
static __global__ void expand_kernel(const int32_t* __restrict__ ent, const int32_t* __restrict__ adj, long rows, int K,
                              int32_t* __restrict__ out, int32_t* __restrict__ stamp, int bit) {
  pdl_enter();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * K) return;
  const long j = i / K;
  const int k = (int)(i % K);
  const int32_t id = __ldg(adj + (long)ent[j] * 2 * K + k);
  if (out) out[i] = id;
  if (stamp) atomicOr(stamp + id, bit);
}

// stamp the neighbours of every entity that is present (count > 0) at the parent level
static __global__ void stamp_children_kernel(const int32_t* __restrict__ present, const int32_t* __restrict__ adj, long n_entity,
                                             int K, int32_t* __restrict__ stamp, int bit) {
  pdl_enter();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_entity * K) return;
  const long e = i / K;
  const int k = (int)(i % K);
  if (__ldg(present + e) > 0) atomicOr(stamp + __ldg(adj + e * 2 * K + k), bit);
}

static __global__ void expand_i64_kernel(const int64_t* __restrict__ ent, const int32_t* __restrict__ adj, long rows,
                                  int K, int64_t* __restrict__ out_e, int64_t* __restrict__ out_r) {
  pdl_enter();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * K) return;
  const long j = i / K;
  const int k = (int)(i % K);
  const int32_t* rec = adj + ent[j] * 2 * K;
  out_e[i] = (int64_t)__ldg(rec + k);
  out_r[i] = (int64_t)__ldg(rec + K + k);
}

static __global__ void copy_i64_kernel(const int64_t* __restrict__ src, long n, int64_t* __restrict__ dst) {
  pdl_enter();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// ---- seeds: ent[0] = item as int32 (model.py:243-256 starts from item_indices); optional stamps -------------
static __global__ void seed_kernel(const int64_t* __restrict__ item, int B, int32_t* __restrict__ ent0,
                            int32_t* __restrict__ stamp) {   // level 0: stamp bit 0
  pdl_enter();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const long e = item[b];
  ent0[b] = (int32_t)e;
  if (stamp) atomicOr(stamp + e, 1);
}

// ---- feed assembly on the device (train.py:112-122, util.py:208-218): the ripple memories of a batch gathered from
// the packed per-user sets  uts int32 [n_user, P, 3, m]  (data_loader_user_set.py:402)  into  mem_x int32 [P, B, m]
static __global__ void gather_feed_kernel(const int32_t* __restrict__ uts, const int64_t* __restrict__ user, int B, int P, int m,
                                   int32_t* __restrict__ mem_h, int32_t* __restrict__ mem_r, int32_t* __restrict__ mem_t) {
  pdl_enter();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;        // over [P][B][m]
  const long n = (long)P * B * m;
  if (i >= n) return;
  const int j = (int)(i % m);
  const long b = (i / m) % B;
  const int hop = (int)(i / ((long)m * B));
  const int32_t* src = uts + (((long)user[b] * P + hop) * 3) * m + j;
  mem_h[i] = __ldg(src);
  mem_r[i] = __ldg(src + m);
  mem_t[i] = __ldg(src + 2 * m);
}

// ---- seeds: ent[0] = item (int32) and Vbuf = E[item]  (model.py:199) -----------------------------------
template <int D>
__global__ void prep_items_kernel(const int64_t* __restrict__ item, ETab E, int B,
                                  int32_t* __restrict__ ent0, float* __restrict__ Vbuf, int32_t* __restrict__ stamp) {
  pdl_enter();
  constexpr int LPR = D / 4;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * LPR) return;
  const long b = i / LPR;
  const int c = (int)(i % LPR);
  const long e = item[b];
  if (c == 0 && ent0) { ent0[b] = (int32_t)e; if (stamp) stamp[e] = 1; }
  st4(Vbuf + b * D + c * 4, ldg4(erow(E, e, D) + c * 4));
}

// ---- User_orient_kg_eh = 0 (model.py:152-156): ukg[b] = U[user[b]], and the user ids as int32 for the backward scatter
template <int D>
__global__ void gather_user_kernel(const int64_t* __restrict__ user, const float* __restrict__ U, int B,
                                   float* __restrict__ ukg, int32_t* __restrict__ user32) {
  pdl_enter();
  constexpr int LPR = D / 4;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * LPR) return;
  const long b = i / LPR;
  const int c = (int)(i % LPR);
  const long u = user[b];
  if (c == 0) user32[b] = (int32_t)u;
  st4(ukg + b * D + c * 4, ldg4(U + u * D + c * 4));
}

// ---- dE[ent[b]] += rows[b]   (row scatter-add of a dense [B, D] buffer) ---------------------------------
template <int D>
__global__ void scatter_rows_kernel(const float* __restrict__ rows, const int32_t* __restrict__ ent, int B, GTab dE) {
  pdl_enter();
  constexpr int LPR = D / 4;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * LPR) return;
  const long b = i / LPR;
  const int c = (int)(i % LPR);
  red_add4(grow_of(dE, ent[b], D) + c * 4, ld4(rows + b * D + c * 4));
}

// ---- out = a + b (+ c) over n4 float4 elements: the gradient of a node vector that has more than two consumers
// (n_mix_hop > 1, model.py:286-315: its own aggregator step, its parent's children sum and the mix layer) --------------
static __global__ void sum_rows_kernel(const float* a, const float* b, const float* c, long n4, float* out) {   // out may alias an input
  pdl_enter();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 v = f4add(ld4(a + i * 4), ld4(b + i * 4));
    if (c) v = f4add(v, ld4(c + i * 4));
    st4(out + i * 4, v);
  }
}

// ---- W = identity [D, D]   (User_orient = 0: the transform kernels run with it, steps.cuh) --------------------------
static __global__ void eye_kernel(float* __restrict__ W, int D) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < D * D) W[i] = (i / D == i % D) ? 1.f : 0.f;
}

// ---- relation scores s[i][r] = Rel[r] . urh_weights_i[D:2D]  (aggregators.py:130-133, relation third) --
static __global__ void rel_scores_kernel(const float* __restrict__ Rel, const float* __restrict__ urh, int n_rel, int D,
                                  int H, float* __restrict__ s) {
  pdl_enter();
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (w >= H * n_rel) return;
  const int i = w / n_rel, r = w % n_rel;
  float acc = 0.f;
  for (int j = lane; j < D; j += 32) acc += Rel[(long)r * D + j] * urh[(long)i * 3 * D + D + j];
  acc = warp_sum(acc);
  if (lane == 0) s[w] = acc;
}

// dRel[r] += sum_i ds[i][r] w_i ;  durh_i[D:2D] += sum_r ds[i][r] Rel[r]     (one CTA per aggregator i)
static __global__ void rel_scores_bwd_kernel(const float* __restrict__ Rel, const float* __restrict__ urh,
                                      const float* __restrict__ ds, int n_rel, int D, float* __restrict__ dRel,
                                      float* __restrict__ durh) {
  pdl_enter();
  const int i = blockIdx.x;
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    const float wj = urh[(long)i * 3 * D + D + j];
    float acc = 0.f;
    for (int r = 0; r < n_rel; ++r) {
      const float d = ds[i * n_rel + r];
      acc += d * Rel[(long)r * D + j];
      atomicAdd(dRel + (long)r * D + j, d * wj);
    }
    atomicAdd(durh + (long)i * 3 * D + D + j, acc);
  }
}

// ---- score (model.py:158-159) -------------------------------------------------------------------------
template <int D>
__global__ void score_kernel(const float* __restrict__ u, const float* __restrict__ item, int B,
                             float* __restrict__ scores, float* __restrict__ scores_norm) {
  pdl_enter();
  constexpr int LPR = D / 4;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long b = i / LPR;
  const int c = (int)(i % LPR);
  float part = 0.f;
  if (b < B) part = f4dot(ld4(u + b * D + c * 4), ld4(item + b * D + c * 4));
  part = group_sum<LPR>(part);
  if (b < B && c == 0) {
    if (scores) scores[b] = part;
    if (scores_norm) scores_norm[b] = 1.f / (1.f + expf(-part));
  }
}

// ---- wide&deep mix + score in one launch (model.py:309-315, :158-159):
//   item[b] = concat_j V[j][0][b] . W_mix + b_mix ;  score[b] = u[b] . item[b]
// RT rows per CTA; thread (r, tx) owns 4 output columns of row r; the concatenated inputs sit in shared memory, the
// weights are read through L1 (every CTA reads the same (H+1) d^2 floats).
template <int D>
__global__ void __launch_bounds__(256) mix_score_kernel(const float* __restrict__ Vtop /* [H+1][B][D] */,
                                                        const float* __restrict__ W, const float* __restrict__ bias,
                                                        const float* __restrict__ u, int B, int H1,
                                                        float* __restrict__ item, float* __restrict__ scores,
                                                        float* __restrict__ scores_norm) {
  pdl_enter();
  constexpr int LPR = D / 4, RT = 256 / LPR;
  extern __shared__ __align__(16) float a_s[];            // [RT][H1 * D] (+ [H1 * D][D] weights when d <= 64)
  const int tid = threadIdx.x, tx = tid % LPR, r = tid / LPR;
  const long row0 = (long)blockIdx.x * RT;
  const int KD = H1 * D;
  constexpr bool STAGE_W = D <= 64;                        // one coalesced pass instead of KD dependent L2 round trips
  float* w_s = a_s + RT * KD;
  if (STAGE_W)
    for (int i = tid * 4; i < KD * D; i += 256 * 4) st4(&w_s[i], ldg4(W + i));
  for (int i = tid; i < RT * H1 * LPR; i += 256) {
    const int rr = i / (H1 * LPR), rem = i % (H1 * LPR), j = rem / LPR, c4 = rem % LPR;
    float4 v = f4zero();
    if (row0 + rr < B) v = ldg4(Vtop + ((long)j * B + row0 + rr) * D + c4 * 4);
    st4(&a_s[rr * KD + j * D + c4 * 4], v);
  }
  __syncthreads();
  const long b = row0 + r;
  float4 acc = ldg4(bias + tx * 4);
  const float* ar = a_s + r * KD;
  if (STAGE_W) {
    const float* Wc = w_s + tx * 4;
#pragma unroll 8
    for (int k = 0; k < KD; ++k) acc = f4fma(ar[k], ld4(Wc + k * D), acc);
  } else {
    const float* Wc = W + tx * 4;
#pragma unroll 8
    for (int k = 0; k < KD; ++k) acc = f4fma(ar[k], ldg4(Wc + (long)k * D), acc);
  }
  float part = 0.f;
  if (b < B) {
    st4(item + b * D + tx * 4, acc);
    part = f4dot(acc, ldg4(u + b * D + tx * 4));
  }
  part = group_sum<LPR>(part);
  if (b < B && tx == 0) {
    scores[b] = part;
    if (scores_norm) scores_norm[b] = 1.f / (1.f + expf(-part));
  }
}

// ---- loss gradient + mix-layer backward in one launch (model.py:379-380, :309-315 backward), d <= 64:
//   g = (sigmoid(score) - label) / B ;  ditem = g u ;  du = g item ;  DC[j][b] = ditem[b] . W_mix[j]^T
// W_mix^T is staged in shared memory once per CTA; RT rows per tile.
template <int D>
__global__ void __launch_bounds__(256) loss_mix_bwd_kernel(const float* __restrict__ scores, const float* __restrict__ labels,
                                                           const float* __restrict__ u, const float* __restrict__ item,
                                                           const float* __restrict__ W, int B, int H1, float invB,
                                                           float* __restrict__ ditem, float* __restrict__ du,
                                                           float* __restrict__ DC /* [H1][B][D] */,
                                                           float* __restrict__ bce_acc) {
  pdl_enter();
  constexpr int LPR = D / 4, RT = 16;
  extern __shared__ __align__(16) float sm[];
  const int KO = H1 * D;                                   // outputs per row
  float* wt_s = sm;                                        // [D (k)][KO]   wt_s[k][j D + n] = W[(j D + n) D + k]... see below
  float* d_s = wt_s + D * KO;                              // [RT][D] ditem tile
  __shared__ float red;
  const int tid = threadIdx.x;
  if (tid == 0) red = 0.f;
  // DC[j][b][n] = sum_k ditem[b][k] W[(j D + n)][k]  with W [(H+1) D, D] row-major  ->  stage transposed: [k][j D + n]
  for (int i = tid; i < KO * D; i += 256) {
    const int o = i / D, k = i % D;                        // coalesced read of W rows
    wt_s[k * KO + o] = __ldg(W + i);
  }
  __syncthreads();
  float bce = 0.f;
  for (long row0 = (long)blockIdx.x * RT; row0 < B; row0 += (long)gridDim.x * RT) {
    for (int i = tid; i < RT * LPR; i += 256) {
      const int rr = i / LPR, c4 = i % LPR;
      const long b = row0 + rr;
      float4 di = f4zero();
      if (b < B) {
        const float x = scores[b], z = labels[b];
        const float gsc = (1.f / (1.f + expf(-x)) - z) * invB;
        di = f4scale(ldg4(u + b * D + c4 * 4), gsc);
        st4(ditem + b * D + c4 * 4, di);
        st4(du + b * D + c4 * 4, f4scale(ldg4(item + b * D + c4 * 4), gsc));
        if (c4 == 0) bce += fmaxf(x, 0.f) - x * z + log1pf(expf(-fabsf(x)));
      }
      st4(&d_s[rr * D + c4 * 4], di);
    }
    __syncthreads();
    for (int i = tid; i < RT * (KO / 4); i += 256) {
      const int rr = i / (KO / 4), oc = i % (KO / 4);
      const long b = row0 + rr;
      if (b >= B) continue;
      float4 acc = f4zero();
      const float* dr = d_s + rr * D;
#pragma unroll 8
      for (int k = 0; k < D; ++k) acc = f4fma(dr[k], ld4(&wt_s[k * KO + oc * 4]), acc);
      const int j = (oc * 4) / D, n = (oc * 4) % D;
      st4(DC + ((long)j * B + b) * D + n, acc);
    }
    __syncthreads();
  }
  bce = warp_sum(bce);
  if (tid % 32 == 0 && bce != 0.f) atomicAdd(&red, bce);
  __syncthreads();
  if (tid == 0 && red != 0.f) atomicAdd(bce_acc, red * invB);
}

// base loss (model.py:379-380) and its gradient (invB = 1 / batch size of the whole job): g = (sigmoid(x) - z) / B ; ditem = g u ; du = g item
template <int D>
__global__ void loss_bwd_kernel(const float* __restrict__ scores, const float* __restrict__ labels,
                                const float* __restrict__ u, const float* __restrict__ item, int B, float invB,
                                float* __restrict__ ditem, float* __restrict__ du, float* __restrict__ bce_acc) {
  pdl_enter();
  constexpr int LPR = D / 4;
  __shared__ float red;
  if (threadIdx.x == 0) red = 0.f;
  __syncthreads();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long b = i / LPR;
  const int c = (int)(i % LPR);
  float bce = 0.f;
  if (b < B) {
    const float x = scores[b], z = labels[b];
    const float gsc = (1.f / (1.f + expf(-x)) - z) * invB;
    st4(ditem + b * D + c * 4, f4scale(ld4(u + b * D + c * 4), gsc));
    st4(du + b * D + c * 4, f4scale(ld4(item + b * D + c * 4), gsc));
    if (c == 0) bce = fmaxf(x, 0.f) - x * z + log1pf(expf(-fabsf(x)));
  }
  bce = warp_sum(bce);
  if (threadIdx.x % 32 == 0 && bce != 0.f) atomicAdd(&red, bce);
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(bce_acc, red * invB);
}

// ---- dense L2 terms (model.py:388-410): grad = coef * mult * param (store: this also zero-fills buffers
// whose coef is 0), loss accumulators += mult * 0.5 * |param|^2 -------------------------------------------
constexpr int MAX_SEG = 24;
struct L2Segments {
  const float* param[MAX_SEG];
  float* grad[MAX_SEG];
  long n[MAX_SEG];
  float coef[MAX_SEG];   // l2_weight or l2_agg_weight (times multiplicity), 0 = no regulariser
  float mult[MAX_SEG];   // multiplicity in the loss value
  int which[MAX_SEG];    // 0 -> l2_loss accumulator, 1 -> l2_agg_loss accumulator
  int count;
};

static __global__ void l2_dense_kernel(L2Segments sg, float* __restrict__ acc /* [1]=l2, [2]=l2_agg */) {
  pdl_enter();
  __shared__ float red[2];
  if (threadIdx.x < 2) red[threadIdx.x] = 0.f;
  __syncthreads();
  for (int sidx = 0; sidx < sg.count; ++sidx) {
    const float* p = sg.param[sidx];
    float* gr = sg.grad[sidx];
    const float coef = sg.coef[sidx];
    float sq = 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < sg.n[sidx]; i += (long)gridDim.x * blockDim.x) {
      const float v = p[i];
      gr[i] = coef * v;
      sq = fmaf(v, v, sq);
    }
    if (sg.mult[sidx] != 0.f) {
      sq = warp_sum(sq);
      if (threadIdx.x % 32 == 0 && sq != 0.f) atomicAdd(&red[sg.which[sidx]], 0.5f * sg.mult[sidx] * sq);
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 && red[threadIdx.x] != 0.f) atomicAdd(acc + 1 + threadIdx.x, red[threadIdx.x]);
}

// ---- un-normalised L2 over the gathered relation-KGE matrices (model.py:386): sum_r cnt[r] |RK[r]|^2 -----
static __global__ void hist_r_kernel(const int32_t* __restrict__ mem_r, long n, int n_rel, float* __restrict__ cnt) {
  pdl_enter();
  extern __shared__ float h[];
  for (int i = threadIdx.x; i < n_rel; i += blockDim.x) h[i] = 0.f;
  __syncthreads();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    atomicAdd(&h[mem_r[i]], 1.f);
  __syncthreads();
  for (int i = threadIdx.x; i < n_rel; i += blockDim.x)
    if (h[i] != 0.f) atomicAdd(cnt + i, h[i]);
}

static __global__ void rk_l2_kernel(const float* __restrict__ RK, const float* __restrict__ cnt, int DD, float two_l2,
                             float* __restrict__ dRK, float* __restrict__ l2_acc) {
  pdl_enter();
  const int r = blockIdx.x;
  const float cr = cnt[r];
  float sq = 0.f;
  for (int i = threadIdx.x; i < DD; i += blockDim.x) {
    const float v = RK[(long)r * DD + i];
    sq = fmaf(v, v, sq);
    if (cr != 0.f) atomicAdd(dRK + (long)r * DD + i, two_l2 * cr * v);
  }
  sq = warp_sum(sq);
  if (threadIdx.x % 32 == 0 && cr != 0.f) atomicAdd(l2_acc, cr * sq);
}

// losses_out = {loss, base_loss, l2_loss, l2_agg_loss}  (model.py:412)
static __global__ void finalize_loss_kernel(const float* __restrict__ acc, float l2w, float l2a, float* __restrict__ out) {
  pdl_enter();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    out[0] = acc[0] + l2w * acc[1] + l2a * acc[2];
    out[1] = acc[0];
    out[2] = acc[1];
    out[3] = acc[2];
  }
}

// ---- transposed copies of the d x d weights: dst[i] = W_a[i]^T (i < H), dst[H + e] = W_t[e]^T (e <= H) ------
static __global__ void transpose_kernel(const float* __restrict__ agg_w, const float* __restrict__ transfer_w, int H, int D,
                                 float* __restrict__ dst) {
  pdl_enter();
  const int b = blockIdx.x;
  const float* s = b < H ? agg_w + (long)b * D * D : transfer_w + (long)(b - H) * D * D;
  float* d = dst + (long)b * D * D;
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
    const int r = i / D, c = i % D;
    d[c * D + r] = s[i];
  }
}

// ---- importance_list (model.py:319-323): p = softmax_k(s_0[rel_k]) for the nodes of one level ------------
static __global__ void importance_kernel(const int32_t* __restrict__ ent, const int32_t* __restrict__ adj,
                                  const float* __restrict__ s, long rows, int K, float* __restrict__ probs) {
  pdl_enter();
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x % 32;
  if (row >= rows) return;
  const int32_t* arow = adj + (long)ent[row] * 2 * K;
  float l0 = -INFINITY, l1 = -INFINITY;
  if (lane < K) l0 = s[__ldg(arow + K + lane)];
  if (lane + 32 < K) l1 = s[__ldg(arow + K + lane + 32)];
  const float mx = warp_max(fmaxf(l0, l1));
  const float e0 = lane < K ? expf(l0 - mx) : 0.f;
  const float e1 = lane + 32 < K ? expf(l1 - mx) : 0.f;
  const float inv = 1.f / warp_sum(e0 + e1);
  if (lane < K) probs[row * K + lane] = e0 * inv;
  if (lane + 32 < K) probs[row * K + lane + 32] = e1 * inv;
}

// ---- sampled fixed-fan-out adjacency on the device (contruct_random_adj, data_loader_user_set.py:375-388) ----------
// Per entity with degree deg in the undirected CSR (construct_kg, :324-343): deg >= K -> K distinct edges, a uniformly
// random K-subset in uniformly random order (np.random.choice(replace=False)); 0 < deg < K -> K independent uniform
// draws (replace=True); deg = 0 -> the row stays zero.  The reference is unseeded; a counter-based generator keyed by
// (seed, entity, draw) makes the result reproducible and independent of the launch geometry.  One thread per entity:
// Floyd's subset algorithm needs only the <= K picks made so far, whatever the degree (hubs reach 4e5 edges).
MVIN_DEV unsigned long long mix64(unsigned long long x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}
// uniform integer in [0, n), n < 2^32 ... 2^40: 64-bit multiply-high of a 64-bit hash (bias < 2^-24)
MVIN_DEV long rand_below(unsigned long long seed, long entity, int draw, long n) {
  const unsigned long long h = mix64(mix64(seed ^ (unsigned long long)entity * 0xd1342543de82ef95ull) + (unsigned long long)draw);
  return (long)__umul64hi(h, (unsigned long long)n);
}
static __global__ void sample_adjacency_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ nbr,
                                        const int32_t* __restrict__ rel, int n_entity, int K, unsigned long long seed,
                                        int32_t* __restrict__ adj_packed, int64_t* __restrict__ adj_entity,
                                        int64_t* __restrict__ adj_relation, int64_t* __restrict__ picked_edges) {
  pdl_enter();
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_entity) return;
  const long beg = indptr[e], deg = indptr[e + 1] - beg;
  long pick[64];                                             // MAX_K
  if (deg >= K) {
    // Floyd: for j = deg-K .. deg-1: t = U[0, j]; take t unless already taken, else take j  -> uniform K-subset
    for (int i = 0; i < K; ++i) {
      const long j = deg - K + i;
      long t = rand_below(seed, e, i, j + 1);
      for (int q = 0; q < i; ++q)
        if (pick[q] == t) { t = j; break; }
      pick[i] = t;
    }
    // Fisher-Yates over the K picks -> uniformly random order
    for (int i = K - 1; i > 0; --i) {
      const int j = (int)rand_below(seed, e, K + i, i + 1);
      const long tmp = pick[i]; pick[i] = pick[j]; pick[j] = tmp;
    }
  } else if (deg > 0) {
    for (int i = 0; i < K; ++i) pick[i] = rand_below(seed, e, i, deg);
  }
  for (int i = 0; i < K; ++i) {
    int32_t n = 0, r = 0;
    long edge = -1;
    if (deg > 0) { edge = beg + pick[i]; n = nbr[edge]; r = rel[edge]; }
    if (adj_packed) { adj_packed[e * 2 * K + i] = n; adj_packed[e * 2 * K + K + i] = r; }
    if (adj_entity) adj_entity[e * K + i] = n;
    if (adj_relation) adj_relation[e * K + i] = r;
    if (picked_edges) picked_edges[e * K + i] = edge;
  }
}

// ---- ripple sets on the device (get_user_triplet_set / _get_user_triplet_set, data_loader_user_set.py:392-441) --------
// Per user and hop: sources = the user's positive items (hop 0) or the previous hop's tails; every source contributes a
// random min(deg, n_neighbor)-subset of its edges as candidates (random.sample, :419); n_memory candidates are then
// drawn, without replacement if there are at least n_memory of them, else with (:431-432); an empty candidate list
// copies the previous hop (:425-426).  One thread per user; a candidate slot is addressed as (source position, q) and
// resolved lazily: the q-th element of the source's subset comes from Floyd's algorithm keyed by (seed, user, hop,
// source position), so the same subset is seen by every draw that lands on that source.
MVIN_DEV long subset_element(unsigned long long key, long deg, int take, int q) {
  if (deg <= take) return q;                              // the whole neighbourhood: slot q is edge q
  long pick[16];
  for (int i = 0; i <= q; ++i) {                          // Floyd, first q + 1 picks (prefix-stable)
    const long j = deg - take + i;
    long t = (long)__umul64hi(mix64(key + (unsigned long long)i), (unsigned long long)(j + 1));
    for (int z = 0; z < i; ++z)
      if (pick[z] == t) { t = j; break; }
    pick[i] = t;
  }
  return pick[q];
}
static __global__ void ripple_sets_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ nbr,
                                   const int32_t* __restrict__ rel, const int64_t* __restrict__ hist_ptr,
                                   const int32_t* __restrict__ hist_items, int n_user, int P, int m, int n_neighbor,
                                   unsigned long long seed, int32_t* __restrict__ uts /* [n_user, P, 3, m] */,
                                   int64_t* __restrict__ slots /* optional [n_user, P, m]: candidate slot ids */) {
  pdl_enter();
  const long u = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_user) return;
  for (int hop = 0; hop < P; ++hop) {
    int32_t* out = uts + ((u * P + hop) * 3) * m;
    const int32_t* src = hop == 0 ? hist_items + hist_ptr[u] : uts + ((u * P + hop - 1) * 3 + 2) * m;
    const long ns = hop == 0 ? hist_ptr[u + 1] - hist_ptr[u] : m;
    long total = 0;
    for (long i = 0; i < ns; ++i) {
      const long d = indptr[src[i] + 1] - indptr[src[i]];
      total += d < n_neighbor ? d : n_neighbor;
    }
    if (total == 0) {
      for (int i = 0; i < 3 * m; ++i) out[i] = hop ? out[i - 3 * m] : 0;
      if (slots) for (int i = 0; i < m; ++i) slots[(u * P + hop) * m + i] = -1;
      continue;
    }
    const unsigned long long key = mix64(seed ^ ((unsigned long long)u * 0x9e3779b97f4a7c15ull + (unsigned long long)hop));
    long pick[64];                                         // n_memory <= 64 on this path
    if (total >= m) {
      for (int i = 0; i < m; ++i) {
        const long j = total - m + i;
        long t = (long)__umul64hi(mix64(key + 0x100 + (unsigned long long)i), (unsigned long long)(j + 1));
        for (int z = 0; z < i; ++z)
          if (pick[z] == t) { t = j; break; }
        pick[i] = t;
      }
      for (int i = m - 1; i > 0; --i) {
        const int j = (int)__umul64hi(mix64(key + 0x200 + (unsigned long long)i), (unsigned long long)(i + 1));
        const long tmp = pick[i]; pick[i] = pick[j]; pick[j] = tmp;
      }
    } else {
      for (int i = 0; i < m; ++i)
        pick[i] = (long)__umul64hi(mix64(key + 0x100 + (unsigned long long)i), (unsigned long long)total);
    }
    for (int i = 0; i < m; ++i) {
      // candidate slot -> (source position, q)
      long c = pick[i], pos = 0;
      for (;; ++pos) {
        const long d = indptr[src[pos] + 1] - indptr[src[pos]];
        const long take = d < n_neighbor ? d : n_neighbor;
        if (c < take) break;
        c -= take;
      }
      const long head = src[pos];
      const long beg = indptr[head], deg = indptr[head + 1] - beg;
      const int take = (int)(deg < n_neighbor ? deg : n_neighbor);
      const long edge = beg + subset_element(mix64(key ^ ((unsigned long long)pos * 0xd1342543de82ef95ull + 0x300)), deg, take, (int)c);
      // the heads / tails of this hop must not overwrite the sources of the SAME hop: hop >= 1 reads the previous
      // hop's block, hop 0 reads the history -- both distinct from `out`
      out[i] = (int32_t)head;
      out[m + i] = rel[edge];
      out[2 * m + i] = nbr[edge];
      if (slots) slots[(u * P + hop) * m + i] = pick[i];
    }
  }
}

// ---- CTR metrics on the device (model.py:419-426, util.py:44-56: per-batch sklearn roc_auc_score / accuracy / f1) ---
// AUC = (#{(i in pos, j in neg): s_i > s_j} + 0.5 #{s_i == s_j}) / (P N): the Mann-Whitney form of the trapezoidal ROC
// area, ties included, as exact integer pair counts (B <= 65536: P N < 2^32 pairs, counted in 64 bits).
// acc / f1 use the reference's threshold (score >= 0.5 -> 1).  out = {auc, acc, f1}; acc64 = 5 zeroed counters.
static __global__ void ctr_count_kernel(const float* __restrict__ scores, const float* __restrict__ labels, int B,
                                 unsigned long long* __restrict__ acc64 /* [0] 2*gt + eq, [1] P, [2] tp, [3] fp, [4] fn */) {
  pdl_enter();
  extern __shared__ float sj[];                       // tile of scores / labels of the "j" side
  float* lj = sj + blockDim.x;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool vi = i < B;
  const float si = vi ? scores[i] : 0.f;
  const bool pi = vi && labels[i] > 0.5f;
  unsigned long long cnt = 0;
  for (int j0 = blockIdx.y * blockDim.x; j0 < B; j0 += gridDim.y * blockDim.x) {
    const int j = j0 + threadIdx.x;
    sj[threadIdx.x] = j < B ? scores[j] : 0.f;
    lj[threadIdx.x] = j < B ? labels[j] : 1.f;        // padding counts as positive: never a negative partner
    __syncthreads();
    if (pi) {
      const int n = min((int)blockDim.x, B - j0);
      unsigned int c = 0;
      for (int t = 0; t < n; ++t)
        if (lj[t] <= 0.5f) c += si > sj[t] ? 2u : (si == sj[t] ? 1u : 0u);
      cnt += c;
    }
    __syncthreads();
  }
  // warp + CTA reduction of the pair count; confusion-matrix counts once (blockIdx.y == 0)
  unsigned long long v[5] = {cnt, 0, 0, 0, 0};
  if (blockIdx.y == 0 && vi) {
    const bool pred = si >= 0.5f;
    v[1] = pi;
    v[2] = pi && pred;
    v[3] = !pi && pred;
    v[4] = pi && !pred;
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(FULL_MASK, v[k], o);
    if (threadIdx.x % 32 == 0 && v[k]) atomicAdd(acc64 + k, v[k]);
  }
}
static __global__ void ctr_finalize_kernel(const unsigned long long* __restrict__ acc64, int B, float* __restrict__ out) {
  pdl_enter();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double P = (double)acc64[1], N = (double)B - P;
  const double tp = (double)acc64[2], fp = (double)acc64[3], fn = (double)acc64[4];
  out[0] = (P > 0 && N > 0) ? (float)(0.5 * (double)acc64[0] / (P * N)) : nanf("");
  out[1] = (float)((tp + (N - fp)) / (double)B);
  out[2] = (2 * tp + fp + fn) > 0 ? (float)(2 * tp / (2 * tp + fp + fn)) : 0.f;
}

// ---- top-K metrics on the device (util.py:137-205 + metrics.py:3-31,34-37,97-100) -----------------------------------
// One CTA per user.  Candidates are ranked by score, descending, ties in candidate order (Python's stable sorted() over
// the insertion-ordered item_score_map, util.py:183-184); rank by counting.  Per k of k_list:
//   precision@k = |top-k & answers| / k,  recall@k = |top-k & answers| / n_answers,
//   ndcg@k = dcg(r_hit[:k]) / dcg(sorted(r_hit, reverse)[:k])  with r_hit = the hit flags of the top k_list[-1] items
//   (the reference builds r_hit once with the loop variable k left over from the precision loop, util.py:190-197) and
//   dcg(r) = sum_i r_i / log2(i + 2)  (metrics.py:15, method 1).
constexpr int TOPK_MAX_K = 8, TOPK_MAX_RANK = 1024;
struct TopkArgs {
  const float* scores;        // [n_users, max_cand]
  const unsigned char* rel;   // [n_users, max_cand]  candidate is one of the user's held-out items
  const int32_t* n_cand;      // [n_users]
  const int32_t* n_answers;   // [n_users]  len(ref_user[user])
  int max_cand, nk;
  int k_list[TOPK_MAX_K];
  float* precision;           // [n_users, nk]
  float* recall;
  float* ndcg;
};
static __global__ void __launch_bounds__(256) topk_metrics_kernel(TopkArgs a) {
  pdl_enter();
  extern __shared__ float sc[];                          // [max_cand] scores, then hit flags by rank
  unsigned char* hit = reinterpret_cast<unsigned char*>(sc + a.max_cand);   // [k_last]
  __shared__ int cnt[TOPK_MAX_K];
  const int u = blockIdx.x, n = a.n_cand[u], k_last = a.k_list[a.nk - 1];
  const float* s = a.scores + (long)u * a.max_cand;
  const unsigned char* rl = a.rel + (long)u * a.max_cand;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sc[i] = s[i];
  for (int i = threadIdx.x; i < k_last; i += blockDim.x) hit[i] = 0;
  if (threadIdx.x < TOPK_MAX_K) cnt[threadIdx.x] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (!rl[i]) continue;                                // only hits matter
    const float si = sc[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (sc[j] > si) || (sc[j] == si && j < i);
    for (int q = 0; q < a.nk; ++q)
      if (rank < a.k_list[q]) atomicAdd(&cnt[q], 1);
    if (rank < k_last) hit[rank] = 1;
  }
  __syncthreads();
  if (threadIdx.x < a.nk) {
    const int q = threadIdx.x, k = a.k_list[q];
    a.precision[(long)u * a.nk + q] = (float)cnt[q] / (float)k;
    a.recall[(long)u * a.nk + q] = a.n_answers[u] > 0 ? (float)cnt[q] / (float)a.n_answers[u] : 0.f;
    const int len = n < k_last ? n : k_last;             // r_hit has one entry per ranked item, at most k_last
    int total_hits = 0;
    for (int i = 0; i < len; ++i) total_hits += hit[i];
    double dcg = 0.0, ideal = 0.0;
    for (int i = 0; i < len && i < k; ++i) {
      const double disc = 1.0 / log2((double)i + 2.0);
      if (hit[i]) dcg += disc;
      if (i < total_hits) ideal += disc;
    }
    a.ndcg[(long)u * a.nk + q] = ideal > 0.0 ? (float)(dcg / ideal) : 0.f;
  }
}

// ---- Adam, TF1 semantics (model.py:414): dense over every segment, ONE pass --------------------------------
// The segments are laid end to end in units of float4 (vec_end[] = running count of 16-byte chunks, the last chunk of
// a segment may be partial); a grid-stride loop over that flat chunk index finds its segment with a short scan of
// the prefix, so small and large segments share one sweep: 16 B read x 4 streams, 16 B written x 3 per chunk, every
// access a full 16-byte vector (segment bases are 16-byte aligned: separate allocations or multiples of d*d floats).
struct AdamSegments {
  float* param[MAX_SEG];
  const float* grad[MAX_SEG];
  float* m[MAX_SEG];
  float* v[MAX_SEG];
  long n[MAX_SEG];
  long vec_end[MAX_SEG];   // cumulative ceil(n / 4)
  int count;
};

MVIN_DEV float adam_one(float& m, float& v, float g, float p, float lr_t, float beta1, float beta2, float eps) {
  m = beta1 * m + (1.f - beta1) * g;
  v = beta2 * v + (1.f - beta2) * g * g;
  return p - lr_t * m / (sqrtf(v) + eps);
}

static __global__ void __launch_bounds__(256) adam_kernel(AdamSegments sg, float lr_t, float beta1, float beta2, float eps) {
  pdl_enter();
  const long total = sg.vec_end[sg.count - 1];
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int s = 0;
    while (i >= sg.vec_end[s]) ++s;
    const long e = (i - (s ? sg.vec_end[s - 1] : 0)) * 4;
    float* p = sg.param[s] + e;
    const float* g = sg.grad[s] + e;
    float* m = sg.m[s] + e;
    float* v = sg.v[s] + e;
    if (e + 4 <= sg.n[s]) {
      const float4 g4 = __ldcs(reinterpret_cast<const float4*>(g));
      float4 m4 = ld4(m), v4 = ld4(v), p4 = ld4(p);
      p4.x = adam_one(m4.x, v4.x, g4.x, p4.x, lr_t, beta1, beta2, eps);
      p4.y = adam_one(m4.y, v4.y, g4.y, p4.y, lr_t, beta1, beta2, eps);
      p4.z = adam_one(m4.z, v4.z, g4.z, p4.z, lr_t, beta1, beta2, eps);
      p4.w = adam_one(m4.w, v4.w, g4.w, p4.w, lr_t, beta1, beta2, eps);
      st4(m, m4);
      st4(v, v4);
      st4(p, p4);
    } else {
      for (long j = 0; e + j < sg.n[s]; ++j) {
        float mj = m[j], vj = v[j];
        p[j] = adam_one(mj, vj, g[j], p[j], lr_t, beta1, beta2, eps);
        m[j] = mj;
        v[j] = vj;
      }
    }
  }
}

}  // namespace mvin
