// level.cuh -- the KG side of the MVIN hot path: per-level fused gather-attend-aggregate kernels.
//
// Replaces (reference, src/model/MVIN/): model.py:267-283 (entity gather + user-oriented transform),
// aggregators.py:98-146 (attention over K, mean of weighted neighbours, (self+agg).W+b, ReLU) as driven by the
// loop at model.py:286-307, and the TF autodiff of the same ops.
//
// Layout: every activation buffer is row-major [rows, D] fp32 with rows = B * K^h (pair-major, child k of node
// j at row j*K+k, model.py:251).  One CTA owns a tile of R = 64 consecutive rows and walks tiles persistently;
// the d x d weights live in shared memory for the CTA's lifetime.  Two thread mappings are used on the tile:
//   * "warp per row" for the neighbour phase: a warp reads the node's packed adjacency record (K ids + K
//     relation ids, contiguous), does the K-softmax with shuffles, then streams the K child rows with G = 32/LPR
//     rows in flight per load instruction (LPR = D/4 lanes x 16 B cover one row).
//   * "register tile" for the dense maps: thread (ty, tx) owns TM rows x 4 columns, FP32 FFMA (see gemm.cuh for
//     why not TF32).
#pragma once
#include "common.cuh"

namespace mvin {

template <int D>
struct TC {
  static constexpr int LPR = D / 4;                  // float4 lanes per row
  static constexpr int NT = (D >= 16) ? 256 : 128;   // threads per CTA
  static constexpr int NTY = NT / LPR;               // thread rows of the register tile
  static constexpr int R = 64;                       // rows per tile
  static constexpr int TM = R / NTY;                 // rows per thread
  static constexpr int LD = D + 4;                   // padded leading dimension of a row tile in smem
  static constexpr int NW = NT / 32;                 // warps per CTA
  static constexpr int G = 32 / LPR;                 // rows one warp load instruction covers
  static constexpr int TMW = (D / NTY) > 0 ? (D / NTY) : 1;   // dW rows per thread in dw_kernel
};

constexpr int MAX_K = 64;

// acc[i][0..3] += sum_k As[(ty*TM+i)][k] * Ws[k][tx*4 .. tx*4+3]
template <int D>
MVIN_DEV void mm_tile(const float* __restrict__ As, const float* __restrict__ Ws, int ty, int tx,
                      float (&acc)[TC<D>::TM][4]) {
  constexpr int TM = TC<D>::TM, LD = TC<D>::LD;
#pragma unroll 2
  for (int k = 0; k < D; k += 4) {
    float4 a[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) a[i] = ld4(&As[(ty * TM + i) * LD + k]);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 w = ld4(&Ws[(k + kk) * D + tx * 4]);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
        acc[i][0] = fmaf(av, w.x, acc[i][0]);
        acc[i][1] = fmaf(av, w.y, acc[i][1]);
        acc[i][2] = fmaf(av, w.z, acc[i][2]);
        acc[i][3] = fmaf(av, w.w, acc[i][3]);
      }
    }
  }
}

template <int D>
MVIN_DEV void load_weight(float* __restrict__ Ws, const float* __restrict__ Wg, int tid) {
  for (int i = tid * 4; i < D * D; i += TC<D>::NT * 4) st4(&Ws[i], ldg4(&Wg[i]));
}

// attention of one node over its K sampled neighbours: p_k = softmax_k(s[rel_k])  (aggregators.py:121-139 with
// the user / self thirds of the logit cancelled).  Lane l handles k = l and k = l + 32.
struct Att {
  float p0, p1;
  int rel0, rel1, id0, id1;
};
MVIN_DEV Att attend(const int32_t* __restrict__ arow, int K, const float* __restrict__ s_s, int lane) {
  Att a;
  a.rel0 = a.rel1 = 0;
  a.id0 = a.id1 = 0;
  float l0 = -INFINITY, l1 = -INFINITY;
  if (lane < K) {
    a.id0 = __ldg(arow + lane);
    a.rel0 = __ldg(arow + K + lane);
    l0 = s_s[a.rel0];
  }
  if (lane + 32 < K) {
    a.id1 = __ldg(arow + lane + 32);
    a.rel1 = __ldg(arow + K + lane + 32);
    l1 = s_s[a.rel1];
  }
  const float mx = warp_max(fmaxf(l0, l1));
  const float e0 = lane < K ? expf(l0 - mx) : 0.f;
  const float e1 = lane + 32 < K ? expf(l1 - mx) : 0.f;
  const float inv = 1.f / warp_sum(e0 + e1);
  a.p0 = e0 * inv;
  a.p1 = e1 * inv;
  return a;
}

// ---------------------------------------------------------------------------------------------------------
// user-oriented transform  T = (E[ent] + u) . W_t[h] + b_t[h]      (model.py:270-283), levels h < L
// ---------------------------------------------------------------------------------------------------------
struct TransformArgs {
  const int32_t* ent;   // [rows]
  const float* E;       // entity table
  const float* u;       // [B, D]  user_o
  const float* W;       // [D, D]
  const float* b;       // [D]
  float* XU;            // [rows, D]  E[ent] + u   (kept for dW_t)
  float* T;             // [rows, D]
  long rows;
  int rpp;              // rows per pair = K^h
};

template <int D>
__global__ void __launch_bounds__(TC<D>::NT) transform_fwd_kernel(TransformArgs a) {
  using C = TC<D>;
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;
  float* As = Ws + D * D;
  const int tid = threadIdx.x, tx = tid % C::LPR, ty = tid / C::LPR;
  load_weight<D>(Ws, a.W, tid);
  const float4 bias = ldg4(a.b + tx * 4);
  const long ntiles = (a.rows + C::R - 1) / C::R;
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long row0 = t * C::R;
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const int r = ty * C::TM + i;
      const long row = row0 + r;
      float4 x = f4zero();
      if (row < a.rows) {
        const long e = a.ent[row];
        x = f4add(ldg4(a.E + e * D + tx * 4), ldg4(a.u + (row / a.rpp) * D + tx * 4));
        st4(a.XU + row * D + tx * 4, x);
      }
      st4(&As[r * C::LD + tx * 4], x);
    }
    __syncthreads();
    float acc[C::TM][4];
#pragma unroll
    for (int i = 0; i < C::TM; ++i) { acc[i][0] = bias.x; acc[i][1] = bias.y; acc[i][2] = bias.z; acc[i][3] = bias.w; }
    mm_tile<D>(As, Ws, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const long row = row0 + ty * C::TM + i;
      if (row < a.rows) st4(a.T + row * D + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    }
    __syncthreads();
  }
}

// accumulate per-pair sums of the rows of a smem tile into du[B, D]; rows of one pair are contiguous, so each
// of D threads walks the tile and flushes one atomic per (pair, column) run.
template <int D>
MVIN_DEV void tile_rows_to_pairs(const float* __restrict__ As, float* __restrict__ du, long row0, long rows,
                                 int rpp, int tid) {
  using C = TC<D>;
  if (tid < D) {
    long cur = -1;
    float acc = 0.f;
    for (int r = 0; r < C::R; ++r) {
      const long row = row0 + r;
      if (row >= rows) break;
      const long b = row / rpp;
      if (b != cur) {
        if (cur >= 0) atomicAdd(du + cur * D + tid, acc);
        cur = b;
        acc = 0.f;
      }
      acc += As[r * C::LD + tid];
    }
    if (cur >= 0) atomicAdd(du + cur * D + tid, acc);
  }
}

struct TransformBwdArgs {
  const int32_t* ent;   // [rows]
  const float* dT;      // [rows, D]
  const float* WT;      // [D, D]  W_t[h] transposed
  float* dE;            // entity-table gradient (scatter-add)
  float* du;            // [B, D]  (accumulated)
  long rows;
  int rpp;
};

template <int D>
__global__ void __launch_bounds__(TC<D>::NT) transform_bwd_kernel(TransformBwdArgs a) {
  using C = TC<D>;
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;
  float* As = Ws + D * D;
  const int tid = threadIdx.x, tx = tid % C::LPR, ty = tid / C::LPR;
  load_weight<D>(Ws, a.WT, tid);
  const long ntiles = (a.rows + C::R - 1) / C::R;
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long row0 = t * C::R;
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const int r = ty * C::TM + i;
      const long row = row0 + r;
      st4(&As[r * C::LD + tx * 4], row < a.rows ? ld4(a.dT + row * D + tx * 4) : f4zero());
    }
    __syncthreads();
    float acc[C::TM][4];
#pragma unroll
    for (int i = 0; i < C::TM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    mm_tile<D>(As, Ws, ty, tx, acc);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const int r = ty * C::TM + i;
      const long row = row0 + r;
      const float4 gx = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      if (row < a.rows) red_add4(a.dE + (long)a.ent[row] * D + tx * 4, gx);
      st4(&As[r * C::LD + tx * 4], gx);
    }
    __syncthreads();
    tile_rows_to_pairs<D>(As, a.du, row0, a.rows, a.rpp, tid);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// one aggregation step, forward  (aggregators.py:98-146; model.py:295-306)
//   LEAF = false:  agg = (1/K) sum_k p_k child[row*K+k]          (children are the rows of the next level)
//   LEAF = true :  S = sum_k p_k E[adj[e][k]];  agg = ((S + u) . W_t[L] + b_t[L]) / K     (hoisted transform)
//   Y = self + agg;   V = relu(Y . W_a + b_a)
// ---------------------------------------------------------------------------------------------------------
struct AggArgs {
  const int32_t* ent;   // [rows] entity id of each node of this level
  const int32_t* adj;   // packed adjacency [n_entity][2][K]
  const float* s;       // [n_rel] relation scores of this aggregator
  const float* child;   // !LEAF: [rows*K, D]
  const float* E;       //  LEAF: entity table
  const float* u;       //  LEAF: [B, D]
  const float* Wt;      //  LEAF: W_t[L] [D, D]
  const float* bt;      //  LEAF: b_t[L]
  const float* self;    // [rows, D]
  const float* Wa;      // [D, D]
  const float* ba;      // [D]
  float* SU;            //  LEAF: [rows, D]  S + u
  float* Y;             // [rows, D]
  float* V;             // [rows, D]
  float* probs;         // optional [rows, K]
  long rows;
  int rpp, K, n_rel;
};

template <int D, bool LEAF>
__global__ void __launch_bounds__(TC<D>::NT) agg_fwd_kernel(AggArgs a) {
  using C = TC<D>;
  extern __shared__ __align__(16) float smem[];
  float* Wa_s = smem;
  float* Wt_s = Wa_s + D * D;
  float* As = Wt_s + (LEAF ? D * D : 0);
  float* pw = As + C::R * C::LD;                      // [NW][MAX_K]
  int* idw = reinterpret_cast<int*>(pw + C::NW * MAX_K);   // [NW][MAX_K]
  float* s_s = reinterpret_cast<float*>(idw + C::NW * MAX_K);
  const int tid = threadIdx.x, tx = tid % C::LPR, ty = tid / C::LPR;
  const int warp = tid / 32, lane = tid % 32, g = lane / C::LPR, c = lane % C::LPR;
  load_weight<D>(Wa_s, a.Wa, tid);
  if (LEAF) load_weight<D>(Wt_s, a.Wt, tid);
  for (int i = tid; i < a.n_rel; i += C::NT) s_s[i] = a.s[i];
  const float4 ba = ldg4(a.ba + tx * 4);
  float4 bt = f4zero();
  if (LEAF) bt = ldg4(a.bt + tx * 4);
  const int K = a.K;
  const float invK = 1.f / (float)K;
  float* pw_w = pw + warp * MAX_K;
  int* idw_w = idw + warp * MAX_K;
  __syncthreads();

  const long ntiles = (a.rows + C::R - 1) / C::R;
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long row0 = t * C::R;
    // ---- neighbour phase: warp per row ----
    for (int r = warp; r < C::R; r += C::NW) {
      const long row = row0 + r;
      if (row < a.rows) {
        const long e = a.ent[row];
        const Att at = attend(a.adj + e * 2 * K, K, s_s, lane);
        pw_w[lane] = at.p0;
        pw_w[lane + 32] = at.p1;
        if (LEAF) { idw_w[lane] = at.id0; idw_w[lane + 32] = at.id1; }
        if (a.probs) {
          if (lane < K) a.probs[row * K + lane] = at.p0;
          if (lane + 32 < K) a.probs[row * K + lane + 32] = at.p1;
        }
        __syncwarp();
        float4 acc = f4zero();
#pragma unroll 4
        for (int k = g; k < K; k += C::G) {
          const float* src = LEAF ? a.E + (long)idw_w[k] * D : a.child + (row * K + k) * D;
          acc = f4fma(pw_w[k], ldg4(src + c * 4), acc);
        }
        acc = cross_group_sum4<C::LPR>(acc);
        if (g == 0) {
          float4 o;
          if (LEAF) {
            o = f4add(acc, ldg4(a.u + (row / a.rpp) * D + c * 4));
            st4(a.SU + row * D + c * 4, o);
          } else {
            o = f4fma(invK, acc, ld4(a.self + row * D + c * 4));
            st4(a.Y + row * D + c * 4, o);
          }
          st4(&As[r * C::LD + c * 4], o);
        }
        __syncwarp();
      } else if (g == 0) {
        st4(&As[r * C::LD + c * 4], f4zero());
      }
    }
    __syncthreads();
    // ---- dense phase: register tile ----
    float acc[C::TM][4];
    if (LEAF) {
#pragma unroll
      for (int i = 0; i < C::TM; ++i) { acc[i][0] = bt.x; acc[i][1] = bt.y; acc[i][2] = bt.z; acc[i][3] = bt.w; }
      mm_tile<D>(As, Wt_s, ty, tx, acc);
      __syncthreads();
#pragma unroll
      for (int i = 0; i < C::TM; ++i) {
        const int r = ty * C::TM + i;
        const long row = row0 + r;
        float4 y = f4zero();
        if (row < a.rows) {
          y = f4fma(invK, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]), ld4(a.self + row * D + tx * 4));
          st4(a.Y + row * D + tx * 4, y);
        }
        st4(&As[r * C::LD + tx * 4], y);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < C::TM; ++i) { acc[i][0] = ba.x; acc[i][1] = ba.y; acc[i][2] = ba.z; acc[i][3] = ba.w; }
    mm_tile<D>(As, Wa_s, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const long row = row0 + ty * C::TM + i;
      if (row < a.rows)
        st4(a.V + row * D + tx * 4, make_float4(fmaxf(acc[i][0], 0.f), fmaxf(acc[i][1], 0.f), fmaxf(acc[i][2], 0.f),
                                                fmaxf(acc[i][3], 0.f)));
    }
    __syncthreads();
  }
}

template <int D, bool LEAF>
constexpr size_t agg_fwd_smem(int n_rel) {
  return sizeof(float) * ((LEAF ? 2 : 1) * D * D + TC<D>::R * TC<D>::LD + 2 * TC<D>::NW * MAX_K + n_rel);
}

// ---------------------------------------------------------------------------------------------------------
// one aggregation step, backward (Appendix B of SURVEY.md; tests/fused_model.py is the CPU twin)
//   gz = gout * [V > 0]            (written in place over gout: it is the G operand of dW_a = Y^T gz)
//   gs = gz . W_a^T ;  dself (+)= gs ;  grow = gs / K
//   !LEAF: dchild[row*K+k] = p_k * grow ; dp_k = grow . child_k
//    LEAF: GROW[row] = grow (G operand of dW_t[L] = SU^T grow);  gsu = grow . W_t[L]^T ;  du[b] += gsu ;
//          dE[n_k] += p_k * gsu ;  dp_k = gsu . E[n_k]
//   dlogit_k = p_k (dp_k - sum_j p_j dp_j) ;  ds[rel_k] += dlogit_k
// ---------------------------------------------------------------------------------------------------------
struct AggBwdArgs {
  const int32_t* ent;
  const int32_t* adj;
  const float* s;
  const float* child;   // !LEAF
  const float* E;       //  LEAF
  const float* WaT;     // W_a transposed
  const float* WtT;     //  LEAF: W_t[L] transposed
  const float* V;       // [rows, D] forward output (ReLU mask)
  float* gout;          // [rows, D] in: dL/dV; out: gz
  float* dself;         // [rows, D]
  float* dchild;        // !LEAF: [rows*K, D]
  float* GROW;          //  LEAF: [rows, D]
  float* dE;            //  LEAF
  float* du;            //  LEAF: [B, D]
  float* ds;            // [n_rel]
  long rows;
  int rpp, K, n_rel;
  int self_accumulate;  // 1: dself += gs, 0: dself = gs
};

template <int D, bool LEAF>
__global__ void __launch_bounds__(TC<D>::NT) agg_bwd_kernel(AggBwdArgs a) {
  using C = TC<D>;
  extern __shared__ __align__(16) float smem[];
  float* Wa_s = smem;
  float* Wt_s = Wa_s + D * D;
  float* Gs = Wt_s + (LEAF ? D * D : 0);
  float* pw = Gs + C::R * C::LD;                       // [NW][MAX_K]
  float* dpw = pw + C::NW * MAX_K;                     // [NW][MAX_K]
  int* idw = reinterpret_cast<int*>(dpw + C::NW * MAX_K);
  float* s_s = reinterpret_cast<float*>(idw + C::NW * MAX_K);
  float* ds_s = s_s + a.n_rel;
  const int tid = threadIdx.x, tx = tid % C::LPR, ty = tid / C::LPR;
  const int warp = tid / 32, lane = tid % 32, g = lane / C::LPR, c = lane % C::LPR;
  load_weight<D>(Wa_s, a.WaT, tid);
  if (LEAF) load_weight<D>(Wt_s, a.WtT, tid);
  for (int i = tid; i < a.n_rel; i += C::NT) { s_s[i] = a.s[i]; ds_s[i] = 0.f; }
  const int K = a.K;
  const float invK = 1.f / (float)K;
  float* pw_w = pw + warp * MAX_K;
  float* dpw_w = dpw + warp * MAX_K;
  int* idw_w = idw + warp * MAX_K;
  __syncthreads();

  const long ntiles = (a.rows + C::R - 1) / C::R;
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long row0 = t * C::R;
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const int r = ty * C::TM + i;
      const long row = row0 + r;
      float4 gz = f4zero();
      if (row < a.rows) {
        const float4 go = ld4(a.gout + row * D + tx * 4);
        const float4 v = ld4(a.V + row * D + tx * 4);
        gz = make_float4(v.x > 0.f ? go.x : 0.f, v.y > 0.f ? go.y : 0.f, v.z > 0.f ? go.z : 0.f,
                         v.w > 0.f ? go.w : 0.f);
        st4(a.gout + row * D + tx * 4, gz);
      }
      st4(&Gs[r * C::LD + tx * 4], gz);
    }
    __syncthreads();
    float acc[C::TM][4];
#pragma unroll
    for (int i = 0; i < C::TM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    mm_tile<D>(Gs, Wa_s, ty, tx, acc);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const int r = ty * C::TM + i;
      const long row = row0 + r;
      const float4 gs = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      const float4 grow = f4scale(gs, invK);
      if (row < a.rows) {
        float* dst = a.dself + row * D + tx * 4;
        st4(dst, a.self_accumulate ? f4add(ld4(dst), gs) : gs);
        if (LEAF) st4(a.GROW + row * D + tx * 4, grow);
      }
      st4(&Gs[r * C::LD + tx * 4], grow);
    }
    __syncthreads();
    if (LEAF) {
#pragma unroll
      for (int i = 0; i < C::TM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      mm_tile<D>(Gs, Wt_s, ty, tx, acc);
      __syncthreads();
#pragma unroll
      for (int i = 0; i < C::TM; ++i)
        st4(&Gs[(ty * C::TM + i) * C::LD + tx * 4], make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
      __syncthreads();
      tile_rows_to_pairs<D>(Gs, a.du, row0, a.rows, a.rpp, tid);
    }
    // ---- neighbour phase: warp per row ----
    for (int r = warp; r < C::R; r += C::NW) {
      const long row = row0 + r;
      if (row >= a.rows) break;
      const long e = a.ent[row];
      const Att at = attend(a.adj + e * 2 * K, K, s_s, lane);
      pw_w[lane] = at.p0;
      pw_w[lane + 32] = at.p1;
      if (LEAF) { idw_w[lane] = at.id0; idw_w[lane + 32] = at.id1; }
      __syncwarp();
      const float4 gr = ld4(&Gs[r * C::LD + c * 4]);
      // uniform trip count: the shuffles inside need every lane of the warp
#pragma unroll 4
      for (int k0 = 0; k0 < K; k0 += C::G) {
        const int k = k0 + g;
        const bool valid = k < K;
        float part = 0.f;
        if (valid) {
          const float pk = pw_w[k];
          if (LEAF) {
            const long n = idw_w[k];
            part = f4dot(gr, ldg4(a.E + n * D + c * 4));
            red_add4(a.dE + n * D + c * 4, f4scale(gr, pk));
          } else {
            const long cr = (row * K + k) * D + c * 4;
            part = f4dot(gr, ld4(a.child + cr));
            st4(a.dchild + cr, f4scale(gr, pk));
          }
        }
        part = group_sum<C::LPR>(part);
        if (valid && c == 0) dpw_w[k] = part;
      }
      __syncwarp();
      const float dp0 = lane < K ? dpw_w[lane] : 0.f;
      const float dp1 = lane + 32 < K ? dpw_w[lane + 32] : 0.f;
      const float dot = warp_sum(at.p0 * dp0 + at.p1 * dp1);
      if (lane < K) atomicAdd(&ds_s[at.rel0], at.p0 * (dp0 - dot));
      if (lane + 32 < K) atomicAdd(&ds_s[at.rel1], at.p1 * (dp1 - dot));
      __syncwarp();
    }
    __syncthreads();
  }
  for (int i = tid; i < a.n_rel; i += C::NT) atomicAdd(a.ds + i, ds_s[i]);
}

template <int D, bool LEAF>
constexpr size_t agg_bwd_smem(int n_rel) {
  return sizeof(float) * ((LEAF ? 2 : 1) * D * D + TC<D>::R * TC<D>::LD + 3 * TC<D>::NW * MAX_K + 2 * n_rel);
}

// ---------------------------------------------------------------------------------------------------------
// dW[i][j] += sum_r A[r][i] G[r][j];  db[j] += sum_r G[r][j]   over rows of A [rows, D] (row stride lda) and
// G [rows, D]
// (the weight gradients of every d x d map on the path)
// ---------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(TC<D>::NT) dw_kernel(const float* __restrict__ A, long lda,
                                                       const float* __restrict__ Gm, long rows,
                                                       float* __restrict__ dW, float* __restrict__ db) {
  using C = TC<D>;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Gs = As + C::R * C::LD;
  const int tid = threadIdx.x, tx = tid % C::LPR, ty = tid / C::LPR;
  const bool active = ty * C::TMW < D;
  float acc[C::TMW][4];
#pragma unroll
  for (int i = 0; i < C::TMW; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  float bsum = 0.f;
  const long ntiles = (rows + C::R - 1) / C::R;
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long row0 = t * C::R;
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const int r = ty * C::TM + i;
      const long row = row0 + r;
      const bool ok = row < rows;
      st4(&As[r * C::LD + tx * 4], ok ? ld4(A + row * lda + tx * 4) : f4zero());
      st4(&Gs[r * C::LD + tx * 4], ok ? ld4(Gm + row * D + tx * 4) : f4zero());
    }
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int r = 0; r < C::R; ++r) {
        const float4 gv = ld4(&Gs[r * C::LD + tx * 4]);
#pragma unroll
        for (int i = 0; i < C::TMW; ++i) {
          const float av = As[r * C::LD + ty * C::TMW + i];
          acc[i][0] = fmaf(av, gv.x, acc[i][0]);
          acc[i][1] = fmaf(av, gv.y, acc[i][1]);
          acc[i][2] = fmaf(av, gv.z, acc[i][2]);
          acc[i][3] = fmaf(av, gv.w, acc[i][3]);
        }
      }
    }
    if (tid < D) {
#pragma unroll 8
      for (int r = 0; r < C::R; ++r) bsum += Gs[r * C::LD + tid];
    }
    __syncthreads();
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < C::TMW; ++i) {
      float* dst = dW + (long)(ty * C::TMW + i) * D + tx * 4;
      atomicAdd(dst + 0, acc[i][0]);
      atomicAdd(dst + 1, acc[i][1]);
      atomicAdd(dst + 2, acc[i][2]);
      atomicAdd(dst + 3, acc[i][3]);
    }
  }
  if (db != nullptr && tid < D) atomicAdd(db + tid, bsum);
}

}  // namespace mvin
