// user.cuh -- the user side of the MVIN hot path in ONE kernel per direction: RippleNet-style o-set propagation
// (_key_addressing) from the item seed to user_o, and its backward down to the entity-table gradient.
//
// Replaces (reference, src/model/MVIN/model.py): the seed lookup :199, the ripple-memory lookups :125-134,
// soft_attention_h_set :162-197, the hop loop :210-229, the user MLP :232-234, and their TF autodiff.  Never
// materialises r_emb_list [B,m,d,d] (model.py:132): the logit  v^T R_m h_m  is evaluated as  Q[b, r_m] . h_m  with
// Q[b,r] = RK[r]^T v_b, because v = E[item] is not updated between hops (model.py:199).
//
// One CTA owns PB pairs.  It (1) reads the seeds v_b = E[item_b], (2) builds Q[b, r, :] for its pairs in shared
// memory (each RK element is read once per CTA and used PB times), (3) runs one warp per (pair, slot) -- slot 0 is
// the h-set attention (model.py:162-197, whose user half and bias are constant along m and cancel in the softmax),
// slot s >= 1 is hop s-1 -- and (4) applies the user MLP to the concatenated o-vectors, all without leaving the SM.
// The warp stages the slot's m memory ids in shared memory first, so every embedding-row load depends on a
// shared-memory read only and the row loops unroll for memory-level parallelism: LPR = D/4 lanes x 16 B cover one
// row, a warp load instruction covers G = 32/LPR rows, and UNR of them are in flight per lane.
#pragma once
#include "common.cuh"

namespace mvin {

constexpr int USER_MAX_NT = 512, USER_UNR = 4;   // CTA = one warp per (pair, slot), 4..16 warps

struct UserArgs {
  ETab E;                  // entity table
  const int64_t* item;     // [B]
  const float* RK;         // relation_emb_KGE [n_rel, D, D]
  const float* w_hi;       // h_emb_item_mlp_matrix [2D] (first D used)
  const float* W_user;     // user_mlp_matrix [(p+1) D, D]
  const float* b_user;     // [D]
  const int32_t* mem_h;    // [max(1,p), B, m]
  const int32_t* mem_r;
  const int32_t* mem_t;
  float* Vbuf;             // [B, D]  v_b
  float* Q;                // [B, n_rel, D]  Q[b, r] = RK[r]^T v_b (re-used by the backward)
  int q_ready;             // 1: Q was produced by a batched GEMM (large RK: n_rel d^2 floats do not stay in L1)
  float* probs;            // [p+1, B, m]
  float* O;                // [B, (p+1) D]   concat(user_h_set, o_0 .. o_{p-1})  (model.py:232)
  float* u;                // [B, D]  user_o
  int B, m, p, n_rel;
};

// shared memory (floats): v[PB][D] | Q[PB][n_rel][D] | O[PB][S*D] | per warp: lg[m] + ids 3m
inline int user_warps(int PB, int p) {
  const int w = PB * (p + 1);
  return w < 4 ? 4 : (w > USER_MAX_NT / 32 ? USER_MAX_NT / 32 : w);
}
// RK is re-streamed by every CTA of the fused kernel: only worth it while it stays L1 / L2-cheap
inline bool user_q_fused(int D, int n_rel) { return (size_t)n_rel * D * D * sizeof(float) <= 256u * 1024; }
inline size_t user_fwd_smem(int D, int PB, int n_rel, int p, int m) {
  const size_t q = user_q_fused(D, n_rel) ? (size_t)PB * n_rel * D : 0;
  return sizeof(float) * ((size_t)PB * D + q + (size_t)PB * (p + 1) * D + (size_t)user_warps(PB, p) * 4 * m);
}
// pairs per CTA: as many as keep the shared memory under ~64 KB (three CTAs per SM); 0 = does not fit at all
inline int user_pairs_per_cta(int D, int n_rel, int p, int m, int pb_max = 4) {
  for (int pb = pb_max; pb >= 1; pb >>= 1)
    if (user_fwd_smem(D, pb, n_rel, p, m) <= (pb == 1 ? 200u : 64u) * 1024) return pb;
  return 0;
}

// Q_s[pb][r][j] = sum_i RK[r][i][j] v_s[pb][i]      (model.py:211-220 refactored: Q[b,r] = RK[r]^T v_b)
template <int D, int PB>
MVIN_DEV void build_q(const float* __restrict__ RK, const float* __restrict__ v_s, float* __restrict__ Q_s, int n_rel,
                      int tid, int nt) {
  for (int o = tid; o < n_rel * D; o += nt) {
    const int r = o / D, j = o % D;
    const float* col = RK + (long)r * D * D + j;
    float acc[PB];
#pragma unroll
    for (int q = 0; q < PB; ++q) acc[q] = 0.f;
#pragma unroll 8
    for (int i = 0; i < D; ++i) {
      const float w = __ldg(col + (long)i * D);
#pragma unroll
      for (int q = 0; q < PB; ++q) acc[q] = fmaf(w, v_s[q * D + i], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < PB; ++q) Q_s[(q * n_rel + r) * D + j] = acc[q];
  }
}

template <int D, int PB>
__global__ void __launch_bounds__(USER_MAX_NT) user_fwd_kernel(UserArgs a) {
  pdl_enter();
  constexpr int LPR = D / 4, G = 32 / LPR;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, g = lane / LPR, c = lane % LPR;
  const int NT = blockDim.x, NW = NT / 32;
  const int m = a.m, S = a.p + 1, n_rel = a.n_rel;
  float* v_s = smem;                                   // [PB][D]
  float* Q_s = v_s + PB * D;                           // [PB][n_rel][D]
  float* O_s = Q_s + (a.q_ready ? 0 : PB * n_rel * D); // [PB][S*D]
  float* lg = O_s + PB * S * D + warp * m;             // [NW][m]
  int32_t* ids = reinterpret_cast<int32_t*>(O_s + PB * S * D + NW * m) + warp * 3 * m;
  int32_t *sh = ids, *sr = ids + m, *stt = ids + 2 * m;
  const long b0 = (long)blockIdx.x * PB;

  // (1) seeds
  for (int i = tid; i < PB * LPR; i += NT) {
    const int q = i / LPR, cc = i % LPR;
    const long b = b0 + q;
    float4 v = f4zero();
    if (b < a.B) {
      const long e = a.item[b];
      v = ldg4(erow(a.E, e, D) + cc * 4);
      st4(a.Vbuf + b * D + cc * 4, v);
    }
    st4(&v_s[q * D + cc * 4], v);
  }
  __syncthreads();
  // (2) Q
  if (a.p > 0 && !a.q_ready) build_q<D, PB>(a.RK, v_s, Q_s, n_rel, tid, NT);
  __syncthreads();
  if (a.p > 0 && !a.q_ready) {
    for (int i = tid * 4; i < PB * n_rel * D; i += NT * 4) {
      const int q = i / (n_rel * D);
      if (b0 + q < a.B) st4(a.Q + (b0 + q) * n_rel * D + (i - q * n_rel * D), ld4(&Q_s[i]));
    }
  }
  // (3) one warp per (pair, slot)
  const float4 wk = ldg4(a.w_hi + c * 4);
  for (int w = warp; w < PB * S; w += NW) {
    const int q = w / S, s = w % S, hop = s ? s - 1 : 0;
    const long b = b0 + q;
    if (b >= a.B) continue;                              // warp-uniform
    const long off = ((long)hop * a.B + b) * m;
    for (int i = lane; i < m; i += 32) {
      sh[i] = __ldg(a.mem_h + off + i);
      if (s) { sr[i] = __ldg(a.mem_r + off + i); stt[i] = __ldg(a.mem_t + off + i); }
    }
    __syncwarp();
    const float* Qb = (a.q_ready ? a.Q + b * n_rel * D : Q_s + (long)q * n_rel * D) + c * 4;   // generic pointer
#pragma unroll USER_UNR
    for (int m0 = 0; m0 < m; m0 += G) {
      const int mm = m0 + g;
      const bool valid = mm < m;
      float part = 0.f;
      if (valid) {
        const float4 hrow = ldg4(erow(a.E, sh[mm], D) + c * 4);
        const float4 key = s == 0 ? wk : ld4(Qb + sr[mm] * D);
        part = f4dot(hrow, key);
      }
      part = group_sum<LPR>(part);
      if (valid && c == 0) lg[mm] = part;
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int i = lane; i < m; i += 32) mx = fmaxf(mx, lg[i]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int i = lane; i < m; i += 32) {
      const float e = expf(lg[i] - mx);
      lg[i] = e;
      sum += e;
    }
    const float inv = 1.f / warp_sum(sum);
    float* pout = a.probs + ((long)s * a.B + b) * m;
    for (int i = lane; i < m; i += 32) {
      const float pr = lg[i] * inv;
      lg[i] = pr;
      pout[i] = pr;
    }
    __syncwarp();
    const int32_t* val = s == 0 ? sh : stt;
    float4 acc = f4zero();
#pragma unroll USER_UNR
    for (int mm = g; mm < m; mm += G) acc = f4fma(lg[mm], ldg4(erow(a.E, val[mm], D) + c * 4), acc);
    acc = cross_group_sum4<LPR>(acc);
    if (g == 0) {
      st4(&O_s[(q * S + s) * D + c * 4], acc);
      st4(a.O + b * S * D + s * D + c * 4, acc);
    }
    __syncwarp();
  }
  __syncthreads();
  // (4) user_o = O . W_user + b      (model.py:232-234); thread (q, j), all threads of a warp share q when D >= 32
  for (int o = tid; o < PB * D; o += NT) {
    const int q = o / D, j = o % D;
    const long b = b0 + q;
    if (b >= a.B) continue;
    float acc0 = __ldg(a.b_user + j), acc1 = 0.f;
    const float* Ob = O_s + q * S * D;
    const float* Wj = a.W_user + j;
#pragma unroll 4
    for (int k = 0; k < S * D; k += 2) {
      acc0 = fmaf(Ob[k], __ldg(Wj + (long)k * D), acc0);
      acc1 = fmaf(Ob[k + 1], __ldg(Wj + (long)(k + 1) * D), acc1);
    }
    a.u[b * D + j] = acc0 + acc1;
  }
}

// ---- backward of the ripple attention: one warp per (pair, slot), Q read back from the forward, dQ accumulated
// with vector reds (the d x d side-maps around it -- dO = du . W_user^T, dv = dQ . RK^T, dRK = v (x) dQ -- are
// batched GEMMs in mvin_capi.cu: fused per-pair matvecs re-stream RK per CTA and measured slower for d >= 64).
constexpr int RIPPLE_NT = 256, RIPPLE_NW = RIPPLE_NT / 32;

struct RippleBwdArgs {
  ETab E;
  const float* Q;
  const float* w_hi;
  const int32_t* mem_h;
  const int32_t* mem_r;
  const int32_t* mem_t;
  const float* probs;      // [p+1, B, m]
  const float* dO;         // [B, (p+1) D]
  GTab dE;                 // entity-table gradient (scatter-add)
  float* dQ;               // [B, n_rel, D] (zeroed by the caller)
  float* dw_hi;            // gradient of h_emb_item_mlp_matrix [2D] (first D touched)
  float* l2_acc;           // += sum over gathered h / t rows of |row|^2   (model.py:383-385)
  float l2_weight;
  int B, m, p, n_rel;
  int* ctr;                // [2] work counter pair (zero between launches, see sched_exit in level.cuh)
};

// per-warp: dl[m], pr[m] floats and ids h[m], r[m], t[m]; per CTA: dwh[D] + l2[1]
inline size_t ripple_bwd_smem(int m, int D) {
  return (size_t)RIPPLE_NW * m * (2 * sizeof(float) + 3 * sizeof(int32_t)) + sizeof(float) * (D + 1);
}

template <int D>
__global__ void __launch_bounds__(RIPPLE_NT) ripple_bwd_kernel(RippleBwdArgs a) {
  pdl_enter();
  constexpr int LPR = D / 4, G = 32 / LPR;
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, g = lane / LPR, c = lane % LPR;
  const int m = a.m, S = a.p + 1;
  float* dl = smem + warp * 2 * m;
  float* pr = dl + m;
  int32_t* ids = reinterpret_cast<int32_t*>(smem + RIPPLE_NW * 2 * m) + warp * 3 * m;
  int32_t *sh = ids, *sr = ids + m, *stt = ids + 2 * m;
  float* dwh_s = smem + RIPPLE_NW * 5 * m;   // [D]
  float* l2_s = dwh_s + D;                   // [1]
  for (int i = threadIdx.x; i < D + 1; i += RIPPLE_NT) dwh_s[i] = 0.f;
  __syncthreads();
  // persistent CTAs; every warp pulls (pair, slot) items from a global counter.  Items of the hop slots (three row
  // passes each) come first, the cheaper h-set items last, so the tail of the launch is made of short items.
  const long n_items = (long)a.B * S, n_heavy = (long)a.B * a.p;
  for (;;) {
    long w = 0;
    if (lane == 0) w = atomicAdd(a.ctr, 1);
    w = __shfl_sync(FULL_MASK, w, 0);
    if (w >= n_items) break;
    const long b = w < n_heavy ? w % a.B : w - n_heavy;
    const int s = w < n_heavy ? (int)(w / a.B) + 1 : 0, hop = s ? s - 1 : 0;
    const long off = ((long)hop * a.B + b) * m;
    const float* prg = a.probs + ((long)s * a.B + b) * m;
    for (int i = lane; i < m; i += 32) {
      sh[i] = __ldg(a.mem_h + off + i);
      pr[i] = prg[i];
      if (s) { sr[i] = __ldg(a.mem_r + off + i); stt[i] = __ldg(a.mem_t + off + i); }
    }
    __syncwarp();
    const float4 wk = ldg4(a.w_hi + c * 4);
    const float* Qb = a.Q + b * a.n_rel * D + c * 4;
    float* dQb = a.dQ + b * a.n_rel * D + c * 4;
    const float4 go = ldg4(a.dO + b * S * D + s * D + c * 4);
    const float two_l2 = 2.f * a.l2_weight;
    const int32_t* val = s == 0 ? sh : stt;
    float l2 = 0.f;

    // pass A: dprob_m = go . value_m ; value-side row gradients
#pragma unroll USER_UNR
    for (int m0 = 0; m0 < m; m0 += G) {
      const int mm = m0 + g;
      const bool valid = mm < m;
      float part = 0.f;
      if (valid) {
        const long id = val[mm];
        const float4 row = ldg4(erow(a.E, id, D) + c * 4);
        part = f4dot(go, row);
        if (s > 0) {
          red_add4(grow_of(a.dE, id, D) + c * 4, f4fma(pr[mm], go, f4scale(row, two_l2)));
          l2 += f4dot(row, row);
        }
      }
      part = group_sum<LPR>(part);
      if (valid && c == 0) dl[mm] = part;
    }
    __syncwarp();
    float dot = 0.f;
    for (int i = lane; i < m; i += 32) dot += pr[i] * dl[i];
    dot = warp_sum(dot);
    for (int i = lane; i < m; i += 32) dl[i] = pr[i] * (dl[i] - dot);
    __syncwarp();
    // pass B: key-side gradients
    if (s == 0) {
      float4 dw = f4zero();
#pragma unroll USER_UNR
      for (int mm = g; mm < m; mm += G) {
        const long hid = sh[mm];
        const float4 hrow = ldg4(erow(a.E, hid, D) + c * 4);
        const float dlm = dl[mm];
        red_add4(grow_of(a.dE, hid, D) + c * 4, f4fma(pr[mm], go, f4scale(wk, dlm)));
        dw = f4fma(dlm, hrow, dw);
      }
      dw = cross_group_sum4<LPR>(dw);
      if (g == 0) {
        atomicAdd(&dwh_s[c * 4 + 0], dw.x);
        atomicAdd(&dwh_s[c * 4 + 1], dw.y);
        atomicAdd(&dwh_s[c * 4 + 2], dw.z);
        atomicAdd(&dwh_s[c * 4 + 3], dw.w);
      }
    } else {
#pragma unroll USER_UNR
      for (int mm = g; mm < m; mm += G) {
        const long hid = sh[mm];
        const long r = sr[mm];
        const float4 hrow = ldg4(erow(a.E, hid, D) + c * 4);
        const float4 key = ldg4(Qb + r * D);
        const float dlm = dl[mm];
        red_add4(grow_of(a.dE, hid, D) + c * 4, f4fma(dlm, key, f4scale(hrow, two_l2)));
        red_add4(dQb + r * D, f4scale(hrow, dlm));
        l2 += f4dot(hrow, hrow);
      }
      l2 = warp_sum(l2);
      if (lane == 0) atomicAdd(l2_s, l2);
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x < D) {
    const float v = dwh_s[threadIdx.x];
    if (v != 0.f) atomicAdd(a.dw_hi + threadIdx.x, v);
  }
  if (threadIdx.x == 0 && l2_s[0] != 0.f) atomicAdd(a.l2_acc, l2_s[0]);
  if (threadIdx.x == 0) {                                  // last CTA out resets the counter pair
    __threadfence();
    if (atomicAdd(a.ctr + 1, 1) == (int)gridDim.x - 1) {
      a.ctr[0] = 0;
      a.ctr[1] = 0;
      __threadfence();
    }
  }
}

}  // namespace mvin
