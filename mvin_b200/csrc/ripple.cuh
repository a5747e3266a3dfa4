// ripple.cuh -- the user side of the MVIN hot path: RippleNet-style o-set propagation (_key_addressing).
//
// Replaces (reference, src/model/MVIN/model.py): the ripple-memory lookups :125-134, soft_attention_h_set
// :162-197 and the hop loop :210-229, plus their TF autodiff.  Never materialises r_emb_list [B,m,d,d]
// (model.py:132): the logit  v^T R_m h_m  is evaluated as  Q[b, r_m] . h_m  with Q[b,r] = RK[r]^T v_b computed
// once per pair by a small GEMM (mvin_capi.cu), because v = E[item] is not updated between hops (model.py:199).
//
// One warp per (pair, slot); slot 0 is the h-set attention (model.py:162-197, whose user half and bias are
// constant along m and cancel in the softmax), slot s >= 1 is hop s-1.  The warp first stages the slot's m memory
// ids (coalesced) in shared memory, so every embedding-row load below depends on a shared-memory read only and the
// row loops can be unrolled for memory-level parallelism: LPR = D/4 lanes x 16 B cover one row, a warp load
// instruction covers G = 32/LPR rows, and UNR of them are in flight per lane.
#pragma once
#include "common.cuh"

namespace mvin {

constexpr int RIPPLE_NT = 256, RIPPLE_NW = RIPPLE_NT / 32, RIPPLE_UNR = 4;

struct RippleArgs {
  ETab E;                  // entity table
  const float* Q;          // [B, n_rel, D]
  const float* w_hi;       // h_emb_item_mlp_matrix [2D] (first D used)
  const int32_t* mem_h;    // [max(1,p), B, m]
  const int32_t* mem_r;
  const int32_t* mem_t;
  float* probs;            // [p+1, B, m]
  float* O;                // [B, (p+1) D]   concat(user_h_set, o_0 .. o_{p-1})  (model.py:232)
  int B, m, p, n_rel;
};

// per-warp shared memory: lg[m] floats, then ids h[m], r[m], t[m]
inline size_t ripple_fwd_smem(int m) { return (size_t)RIPPLE_NW * m * (sizeof(float) + 3 * sizeof(int32_t)); }

template <int D>
__global__ void __launch_bounds__(RIPPLE_NT) ripple_fwd_kernel(RippleArgs a) {
  constexpr int LPR = D / 4, G = 32 / LPR;
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, g = lane / LPR, c = lane % LPR;
  const int m = a.m, S = a.p + 1;
  float* lg = smem + warp * m;
  int32_t* ids = reinterpret_cast<int32_t*>(smem + RIPPLE_NW * m) + warp * 3 * m;
  int32_t *sh = ids, *sr = ids + m, *stt = ids + 2 * m;
  const long w = (long)blockIdx.x * RIPPLE_NW + warp;
  if (w >= (long)a.B * S) return;
  const long b = w / S;
  const int s = (int)(w % S), hop = s ? s - 1 : 0;
  const long off = ((long)hop * a.B + b) * m;
  for (int i = lane; i < m; i += 32) {
    sh[i] = __ldg(a.mem_h + off + i);
    if (s) { sr[i] = __ldg(a.mem_r + off + i); stt[i] = __ldg(a.mem_t + off + i); }
  }
  __syncwarp();
  const float4 wk = ldg4(a.w_hi + c * 4);
  const float* Qb = a.Q + b * a.n_rel * D + c * 4;

#pragma unroll RIPPLE_UNR
  for (int m0 = 0; m0 < m; m0 += G) {
    const int mm = m0 + g;
    const bool valid = mm < m;
    float part = 0.f;
    if (valid) {
      const float4 hrow = ldg4(erow(a.E, sh[mm], D) + c * 4);
      const float4 key = s == 0 ? wk : ldg4(Qb + (long)sr[mm] * D);
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
#pragma unroll RIPPLE_UNR
  for (int mm = g; mm < m; mm += G) acc = f4fma(lg[mm], ldg4(erow(a.E, val[mm], D) + c * 4), acc);
  acc = cross_group_sum4<LPR>(acc);
  if (g == 0) st4(a.O + b * S * D + s * D + c * 4, acc);
}

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
};

// per-warp: dl[m], pr[m] floats and ids h[m], r[m], t[m]; per CTA: dwh[D] + l2[1]
inline size_t ripple_bwd_smem(int m, int D) {
  return (size_t)RIPPLE_NW * m * (2 * sizeof(float) + 3 * sizeof(int32_t)) + sizeof(float) * (D + 1);
}

template <int D>
__global__ void __launch_bounds__(RIPPLE_NT) ripple_bwd_kernel(RippleBwdArgs a) {
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
  const long w = (long)blockIdx.x * RIPPLE_NW + warp;
  if (w < (long)a.B * S) {
    const long b = w / S;
    const int s = (int)(w % S), hop = s ? s - 1 : 0;
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
#pragma unroll RIPPLE_UNR
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
#pragma unroll RIPPLE_UNR
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
#pragma unroll RIPPLE_UNR
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
  }
  __syncthreads();
  if (threadIdx.x < D) {
    const float v = dwh_s[threadIdx.x];
    if (v != 0.f) atomicAdd(a.dw_hi + threadIdx.x, v);
  }
  if (threadIdx.x == 0 && l2_s[0] != 0.f) atomicAdd(a.l2_acc, l2_s[0]);
}

}  // namespace mvin
