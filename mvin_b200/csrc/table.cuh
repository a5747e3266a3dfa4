// table.cuh -- the ENTITY-TABLE form of aggregator iteration 0 (CPU twin:
// tests/fused_model.py, table=True).
//
// Reference (src/model/MVIN/): model.py:270-283 (user-oriented transform of every level), aggregators.py:98-146 as
// driven by the first pass of the loop at model.py:286-307, and the TF autodiff of both.
//
// With the attention relation-only (fact 1) and sum_k p_k = 1 (fact 2) the hoist of the leaf level holds at EVERY level
// of iteration 0: the children of a level-h node with entity e are exactly adj[e], transformed by W_t[h+1].  The
// pre-activation of iteration 0 at such a node of pair b is therefore
//     pre = (E[e] + u_b) M1_h + (S_e + u_b) M2_h + c_h  =  A_h[e] + C_h[b]
//     M1_h = W_t[h] W_a0,   M2_h = W_t[h+1] W_a0 / K,   c_h = (b_t[h] + b_t[h+1] / K) W_a0 + b_a0,
//     S_e  = sum_k p_k(e) E[adj[e][k]]                  (leaf_entity_kernel, level.cuh)
// a per-ENTITY table A_h = E M1_h + Se M2_h + c_h plus a per-PAIR vector C_h = u (M1_h + M2_h).  Iteration 0 has no
// per-row dense map and no parent-child traffic; V[1][h][row] = relu(A_h[ent] + C_h[pair]) is a table gather, and the
// deepest level (B K^(H-1) rows) is never materialised: iteration 1 gathers it from A_{H-1} through the packed
// adjacency record (agg_fwd_kernel / agg_bwd_kernel, `virt` levels).  Backward: the pre-activation gradients are
// summed per entity (dA_h, red.global.add.v4 into an L2-resident table) and per pair (dCs_h); below them everything is
// dense algebra on [n_entity, d] / [B, d] matrices and a d x d parameter chain.
#pragma once
#include "level.cuh"

namespace mvin {

// composed maps of one level (compose_kernel): TBL_NM matrices of D x D floats
constexpr int TBL_M1 = 0, TBL_M2 = 1, TBL_MSUM = 2, TBL_M1T = 3, TBL_M2T = 4, TBL_NM = 5;
constexpr int COMPOSE_SPLIT = 16;     // CTAs per level of the two d x d chain kernels

// ---- M1_h, M2_h, their sum and transposes, c_h   (parameters only) -------------------------------------------
// grid (H, COMPOSE_SPLIT) x 256
static __global__ void compose_kernel(const float* __restrict__ Wt, const float* __restrict__ bt,
                                      const float* __restrict__ Wa0, const float* __restrict__ ba0, int D, float invK,
                                      float* __restrict__ M, float* __restrict__ cst) {
  pdl_enter();
  const int h = blockIdx.x;
  const int rows_per = (D + COMPOSE_SPLIT - 1) / COMPOSE_SPLIT;
  const float* W0 = Wt + (long)h * D * D;
  const float* W1 = W0 + (long)D * D;
  float* Mh = M + (long)h * TBL_NM * D * D;
  for (int idx = threadIdx.x; idx < rows_per * D; idx += blockDim.x) {
    const int i = blockIdx.y * rows_per + idx / D, j = idx % D;
    if (i >= D) break;
    float m1 = 0.f, m2 = 0.f;
    for (int k = 0; k < D; ++k) {
      const float w = __ldg(Wa0 + (long)k * D + j);
      m1 = fmaf(__ldg(W0 + (long)i * D + k), w, m1);
      m2 = fmaf(__ldg(W1 + (long)i * D + k), w, m2);
    }
    m2 *= invK;
    Mh[(long)TBL_M1 * D * D + i * D + j] = m1;
    Mh[(long)TBL_M2 * D * D + i * D + j] = m2;
    Mh[(long)TBL_MSUM * D * D + i * D + j] = m1 + m2;
    Mh[(long)TBL_M1T * D * D + j * D + i] = m1;
    Mh[(long)TBL_M2T * D * D + j * D + i] = m2;
  }
  if (blockIdx.y == 0) {
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
      float c = __ldg(ba0 + j);
      for (int k = 0; k < D; ++k)
        c = fmaf(__ldg(bt + (long)h * D + k) + invK * __ldg(bt + (long)(h + 1) * D + k), __ldg(Wa0 + (long)k * D + j), c);
      cst[(long)h * D + j] = c;
    }
  }
}

// ---- parameter chain, backward: (dM1_h + dMu_h, dM2_h + dMu_h, dc_h) -> dW_t, db_t, dW_a0, db_a0 ---------------
// dM [H][3][D][D] = {sum_e E^T dA, sum_e Se^T dA, u^T dCs};  grid (H, COMPOSE_SPLIT) x 256; all outputs accumulated
static __global__ void compose_bwd_kernel(const float* __restrict__ Wt, const float* __restrict__ bt,
                                          const float* __restrict__ Wa0, const float* __restrict__ dM,
                                          const float* __restrict__ dc, int D, float invK, float* __restrict__ dWt,
                                          float* __restrict__ dbt, float* __restrict__ dWa0, float* __restrict__ dba0) {
  pdl_enter();
  const int h = blockIdx.x;
  const int rows_per = (D + COMPOSE_SPLIT - 1) / COMPOSE_SPLIT;
  const float* W0 = Wt + (long)h * D * D;
  const float* W1 = W0 + (long)D * D;
  const float* G1 = dM + (long)h * 3 * D * D;
  const float* G2 = G1 + (long)D * D;
  const float* Gu = G2 + (long)D * D;
  const float* dch = dc + (long)h * D;
  for (int idx = threadIdx.x; idx < rows_per * D; idx += blockDim.x) {
    const int a = blockIdx.y * rows_per + idx / D, b = idx % D;
    if (a >= D) break;
    // dW_t[h][a][b] += sum_j (dM1 + dMu)[a][j] W_a0[b][j];   dW_t[h+1][a][b] += (1/K) sum_j (dM2 + dMu)[a][j] W_a0[b][j]
    float t0 = 0.f, t1 = 0.f;
    // dW_a0[a][b] += sum_i W_t[h][i][a] (dM1 + dMu)[i][b] + (1/K) sum_i W_t[h+1][i][a] (dM2 + dMu)[i][b]
    float wa = 0.f, wb = 0.f;
    for (int j = 0; j < D; ++j) {
      const float gu = __ldg(Gu + (long)a * D + j), w = __ldg(Wa0 + (long)b * D + j);
      t0 = fmaf(__ldg(G1 + (long)a * D + j) + gu, w, t0);
      t1 = fmaf(__ldg(G2 + (long)a * D + j) + gu, w, t1);
      const float gub = __ldg(Gu + (long)j * D + b);
      wa = fmaf(__ldg(W0 + (long)j * D + a), __ldg(G1 + (long)j * D + b) + gub, wa);
      wb = fmaf(__ldg(W1 + (long)j * D + a), __ldg(G2 + (long)j * D + b) + gub, wb);
    }
    atomicAdd(dWt + (long)h * D * D + a * D + b, t0);
    atomicAdd(dWt + (long)(h + 1) * D * D + a * D + b, invK * t1);
    const float bsum = __ldg(bt + (long)h * D + a) + invK * __ldg(bt + (long)(h + 1) * D + a);
    atomicAdd(dWa0 + (long)a * D + b, wa + invK * wb + bsum * __ldg(dch + b));
  }
  if (blockIdx.y == 0) {
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
      float t = 0.f;
      for (int j = 0; j < D; ++j) t = fmaf(__ldg(dch + j), __ldg(Wa0 + (long)k * D + j), t);
      atomicAdd(dbt + (long)h * D + k, t);
      atomicAdd(dbt + (long)(h + 1) * D + k, invK * t);
      atomicAdd(dba0 + k, __ldg(dch + k));
    }
  }
}

// ---- per-entity tables ------------------------------------------------------------------------------------------
struct TableArgs {
  const int32_t* stamp;   // [n_entity] bit h set: the entity occurs at level h of this batch (A_h is needed for it)
  const float* E;         // [n_entity, D]
  const float* Se;        // [n_entity, D]   (valid where stamped)
  const float* M;         // [H][TBL_NM][D][D]
  const float* cst;       // [H][D]
  float* A;               // fwd out [H][n_entity][D]
  const float* dA;        // bwd in  [H][n_entity][D]
  float* dE;              // bwd out [n_entity, D]  (+=)
  float* GSe;             // bwd out [n_entity, D]  (+=)  gradient of Se, consumed by leaf_entity_kernel<BWD>
  float* dM;              // bwd out [H][3][D][D]   (+=)  slots 0, 1
  float* dc;              // bwd out [H][D]         (+=)
  long n_entity;
};

// A_h[e] = E[e] M1_h + Se[e] M2_h + c_h for the stamped entities; grid (CTAs, H), tiles of R consecutive entities
template <int D>
__global__ void __launch_bounds__(TC<D>::NT) table_fwd_kernel(TableArgs a) {
  pdl_enter();
  using C = TC<D>;
  extern __shared__ __align__(16) float smem[];
  float* W1s = smem;
  float* W2s = W1s + C::WSZ;
  float* Es = W2s + C::WSZ;
  float* Ss = Es + C::R * C::LD;
  const int tid = threadIdx.x, tx = tid % C::LPR, ty = tid / C::LPR;
  const int h = blockIdx.y;
  const float* Mh = a.M + (long)h * TBL_NM * D * D;
  load_weight<D>(W1s, Mh + (long)TBL_M1 * D * D, tid);
  load_weight<D>(W2s, Mh + (long)TBL_M2 * D * D, tid);
  const float4 bias = ldg4(a.cst + (long)h * D + tx * 4);
  float* Ah = a.A + (long)h * a.n_entity * D;
  const long ntiles = (a.n_entity + C::R - 1) / C::R;
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long row0 = t * C::R;
    const int mine = (tid < C::R && row0 + tid < a.n_entity && ((__ldg(a.stamp + row0 + tid) >> h) & 1) != 0) ? 1 : 0;
    if (!__syncthreads_or(mine)) continue;                 // also fences the previous tile's reads of Es / Ss
    bool on[C::TM];
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const int r = ty * C::TM + i;
      const long row = row0 + r;
      on[i] = row < a.n_entity && ((__ldg(a.stamp + row) >> h) & 1) != 0;
      float4 e = f4zero(), s = f4zero();
      if (on[i]) { e = ldg4(a.E + row * D + tx * 4); s = ld4(a.Se + row * D + tx * 4); }
      st4(&Es[r * C::LD + tx * 4], e);
      st4(&Ss[r * C::LD + tx * 4], s);
    }
    __syncthreads();
    float acc[C::TM][4];
#pragma unroll
    for (int i = 0; i < C::TM; ++i) { acc[i][0] = bias.x; acc[i][1] = bias.y; acc[i][2] = bias.z; acc[i][3] = bias.w; }
    tile_mm<D>(Es, W1s, ty, tx, acc);
    tile_mm<D>(Ss, W2s, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const long row = row0 + ty * C::TM + i;
      if (on[i]) st4(Ah + row * D + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    }
  }
}

// backward of the tables:  dE[e] += dA_h[e] M1_h^T ;  GSe[e] += dA_h[e] M2_h^T ;  dM1_h += E^T dA_h ;  dM2_h += Se^T dA_h ;
// dc_h += sum_e dA_h[e].   grid (CTAs, H)
template <int D>
__global__ void __launch_bounds__(TC<D>::NT) table_bwd_kernel(TableArgs a) {
  pdl_enter();
  using C = TC<D>;
  extern __shared__ __align__(16) float smem[];
  float* W1s = smem;                       // M1_h^T
  float* W2s = W1s + C::WSZ;               // M2_h^T
  float* Es = W2s + C::WSZ;
  float* Ss = Es + C::R * C::LD;
  float* Gs = Ss + C::R * C::LD;
  float* Gs2 = Gs + C::R * C::LD;
  const int tid = threadIdx.x, tx = tid % C::LPR, ty = tid / C::LPR;
  const int h = blockIdx.y;
  const float* Mh = a.M + (long)h * TBL_NM * D * D;
  load_weight<D>(W1s, Mh + (long)TBL_M1T * D * D, tid);
  load_weight<D>(W2s, Mh + (long)TBL_M2T * D * D, tid);
  const float* dAh = a.dA + (long)h * a.n_entity * D;
  float dw1[C::DWN][4], dw2[C::DWN][4];
#pragma unroll
  for (int i = 0; i < C::DWN; ++i) {
    dw1[i][0] = dw1[i][1] = dw1[i][2] = dw1[i][3] = 0.f;
    dw2[i][0] = dw2[i][1] = dw2[i][2] = dw2[i][3] = 0.f;
  }
  float4 bpart = f4zero();
  const long ntiles = (a.n_entity + C::R - 1) / C::R;
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long row0 = t * C::R;
    const int mine = (tid < C::R && row0 + tid < a.n_entity && ((__ldg(a.stamp + row0 + tid) >> h) & 1) != 0) ? 1 : 0;
    if (!__syncthreads_or(mine)) continue;
    bool on[C::TM];
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const int r = ty * C::TM + i;
      const long row = row0 + r;
      on[i] = row < a.n_entity && ((__ldg(a.stamp + row) >> h) & 1) != 0;
      float4 e = f4zero(), s = f4zero(), g = f4zero();
      if (on[i]) {
        e = ldg4(a.E + row * D + tx * 4);
        s = ld4(a.Se + row * D + tx * 4);
        g = ld4(dAh + row * D + tx * 4);
      }
      bpart = f4add(bpart, g);
      st4(&Es[r * C::LD + tx * 4], e);
      st4(&Ss[r * C::LD + tx * 4], s);
      st4(&Gs[r * C::LD + tx * 4], g);
      st4(&Gs2[r * C::LD + tx * 4], g);
    }
    __syncthreads();
    dw_tile<D>(Es, Gs, ty, tx, dw1);
    dw_tile<D>(Ss, Gs, ty, tx, dw2);
    float acc[C::TM][4];
#pragma unroll
    for (int i = 0; i < C::TM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    tile_mm<D>(Gs, W1s, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const long row = row0 + ty * C::TM + i;
      if (on[i]) red_add4(a.dE + row * D + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
      acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    }
    tile_mm<D>(Gs2, W2s, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const long row = row0 + ty * C::TM + i;
      if (on[i]) red_add4(a.GSe + row * D + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    }
  }
  __syncthreads();
  dw_flush<D>(dw1, a.dM + (long)(h * 3 + 0) * D * D, ty, tx);
  dw_flush<D>(dw2, a.dM + (long)(h * 3 + 1) * D * D, ty, tx);
  bias_flush<D>(bpart, Es, a.dc + (long)h * D, tid);
}

template <int D>
constexpr size_t table_fwd_smem() { return sizeof(float) * (2 * TC<D>::WSZ + 2 * TC<D>::R * TC<D>::LD); }
template <int D>
constexpr size_t table_bwd_smem() { return sizeof(float) * (2 * TC<D>::WSZ + 4 * TC<D>::R * TC<D>::LD); }

// ---- materialised rows of iteration 0: V[1][h][row] = relu(A_h[ent[row]] + C_h[pair]) for the (small) levels that
// later iterations read as rows; backward: dpre = dV * [V > 0] summed per entity (dA_h) and per pair (dCs_h) ----------
struct VirtLevel {
  const int32_t* ent;     // [rows]
  const float* A;         // [n_entity, D]  A_h
  const float* Cp;        // [B, D]         C_h
  float* V;               // [rows, D]      fwd out / bwd mask
  const float* g1;        // bwd in [rows, D]
  const float* g2;        // bwd in, optional
  float* dA;              // bwd out [n_entity, D] (+=)
  float* dCs;             // bwd out [B, D] (+=)
  unsigned long long rpp_magic;
};
struct VirtArgs {
  VirtLevel lv[MAX_LV];
  long end[MAX_LV];       // cumulative row counts
  int nlev;
};

template <int D, bool BWD>
__global__ void __launch_bounds__(256) virt_rows_kernel(VirtArgs a) {
  pdl_enter();
  constexpr int LPR = D / 4;
  const long total = a.end[a.nlev - 1] * LPR;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long gr = i / LPR;
    const int c = (int)(i % LPR);
    int l = 0;
    while (gr >= a.end[l]) ++l;
    const VirtLevel& L = a.lv[l];
    const long row = gr - (l ? a.end[l - 1] : 0);
    const long pair = fastdiv(row, L.rpp_magic);
    const long e = __ldg(L.ent + row);
    if (!BWD) {
      const float4 x = f4add(ld4(L.A + e * D + c * 4), ld4(L.Cp + pair * D + c * 4));
      st4(L.V + row * D + c * 4, make_float4(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f), fmaxf(x.z, 0.f), fmaxf(x.w, 0.f)));
    } else {
      float4 g = ld4(L.g1 + row * D + c * 4);
      if (L.g2) g = f4add(g, ld4(L.g2 + row * D + c * 4));
      const float4 v = ld4(L.V + row * D + c * 4);
      g = make_float4(v.x > 0.f ? g.x : 0.f, v.y > 0.f ? g.y : 0.f, v.z > 0.f ? g.z : 0.f, v.w > 0.f ? g.w : 0.f);
      red_add4(L.dA + e * D + c * 4, g);
      red_add4(L.dCs + pair * D + c * 4, g);
    }
  }
}

}  // namespace mvin
