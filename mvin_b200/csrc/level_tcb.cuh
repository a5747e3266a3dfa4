// level_tcb.cuh -- tcgen05 BACKWARD row kernels (d in {32, 64}) for the level that dominates the step: the deepest
// materialised level lv = L - 1, whose neighbours are the hoisted leaf entities (per-entity leaf mode, level.cuh).
//
// Reference semantics: TF autodiff of model.py:270-283 (user-oriented transform) and aggregators.py:98-146 as driven by
// model.py:286-307, for aggregator iteration 0 (tests/fused_model.py is the CPU twin).
//
// What changes against the mma.sync kernels of level.cuh (which stay the general path):
//   * every d x d product of the backward runs on the 5th-generation tensor cores from split-bf16 operands
//     (umma_bf.cuh): the dX products (gz . W_a^T, grow . W_t[L]^T, dT . W_t[lv]^T) into a 128-lane tensor-memory
//     accumulator, and the weight gradients (Y^T gz, SU^T grow, XU^T dT) into 64-lane accumulators that stay in tensor
//     memory for the CTA's lifetime and are flushed once -- on the legacy path these six products (x 3 for the TF32
//     split) were the bound of the step at d = 64 (16.8 + 7.1 ms of 39.7 ms at C4, profiles/README.md).
//   * the parent's share of the row gradient is formed in place: the level-(lv - 1) tiles no longer write
//     dchild[row K + k] = p_k grow_parent (one level-sized buffer written, then read by the transform backward); they
//     leave grow_parent [rows / K, d] and this kernel adds p_k grow_parent to its own dself, writes ONE buffer dT for the
//     transform backward, and evaluates the parent's softmax gradient dp_k = grow_parent . T_row on the rows it already
//     streams.  Level-sized streams of the backward at this level: 10 -> 6 (S + u is recomputed from the per-entity
//     table instead of stored).
//
// Tile = 128 consecutive rows = 128 / K whole families (K children of one level-(lv-1) node; requires 128 % K == 0).
// 256 threads.  Two thread mappings are used: the ROW mapping (ty, tx) of level.cuh for coalesced global loads (LPR = d/4
// lanes x 16 bytes per row), and the ACCUMULATOR mapping of tcgen05.ld (warp w: tensor-memory lanes 32 (w % 4) .. + 31 =
// tile rows, warps 0-3 columns [0, d/2), warps 4-7 columns [d/2, d)): each thread owns d/2 consecutive floats of one row.
#pragma once
#include "level.cuh"
#include "level_tc.cuh"
#include "umma_bf.cuh"

namespace mvin {

template <int D>
struct TB {
  static constexpr int NT = 256, NW = 8, R = 128, NP = 2;
  static constexpr int LPR = D / 4, RP = NT / LPR, NPASS = R / RP;
  static constexpr int HC = D / 2;                                 // accumulator columns per thread
  using LA = umma::BfLayout<D, 144>;                               // activation tiles
  using LW = umma::BfLayout<D, 128>;                               // weights
  static constexpr int APL = LA::bytes(R), WPL = LW::bytes(D);     // bytes per plane
  static constexpr int ATILE = NP * APL + (D < 64 ? 1024 : 0);     // (M = 64 view of a d = 32 tile reads past the planes)
  static constexpr int WTILE = NP * WPL;
};

// column sums over the 32 rows a warp holds in the accumulator mapping: v[j] (j < 16) of lane l = row -> the sum of
// column (lane % 16) over all lanes, valid in every lane; butterfly reduce-scatter (15 shuffles) + one more fold
MVIN_DEV float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < o; ++j) {
      const float send = up ? v[j] : v[j + o];
      const float keep = up ? v[j + o] : v[j];
      v[j] = keep + __shfl_xor_sync(FULL_MASK, send, o);
    }
  }
  return v[0] + __shfl_xor_sync(FULL_MASK, v[0], 16);              // column = lane % 16
}

// flush a 64-lane weight-gradient accumulator (row i of dW in lane 32 (i / 16) + i % 16) with vector reductions
template <int D>
MVIN_DEV void flush_dw_tmem(uint32_t tmem_w, float* __restrict__ dW, int warp, int lane) {
  const int q = warp % 4, c0 = (warp / 4) * (D / 2);
  const int i = 16 * q + lane;                                     // valid for lane < 16
#pragma unroll
  for (int cc = 0; cc < D / 2; cc += 16) {
    float v[16];
    umma::tmem_ld16(tmem_w + ((uint32_t)(32 * q) << 16) + (uint32_t)(c0 + cc), v);
    if (lane < 16 && i < D) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) red_add4(dW + (long)i * D + c0 + cc + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// leaf level of aggregator iteration 0, backward (entity mode)
// ---------------------------------------------------------------------------------------------------------
struct LeafBwdArgs {
  const int32_t* ent;      // [rows]      entity of each level-lv node
  const int32_t* ent_par;  // [rows / K]  entity of each level-(lv-1) node (its adjacency record gives p_k, rel_k)
  const int32_t* adj;      // packed adjacency
  const float* s;          // [n_rel] relation scores of aggregator 0
  const float* g1;         // [rows, D] gradient of V[1][lv] (from the parent's dchild of iteration 1)
  const float* V1;         // [rows, D] forward output of iteration 0 (ReLU mask)
  const float* Y;          // [rows, D] forward GEMM input of iteration 0
  const float* T;          // [rows, D] transformed rows V[0][lv] (children of the level-(lv-1) nodes)
  const float* GP;         // [rows / K, D] grow of the parents (gs / K of their own aggregator step)
  const float* Se;         // [n_entity, D]
  const float* u;          // [B, D]
  const float* Wa;         // [D, D] aggregator 0 weights
  const float* Wt;         // [D, D] W_t[L]
  float* dT;               // out [rows, D]  dself + p_k grow_parent
  float* dWa; float* dba; float* dWt; float* dbt;
  float* GSe;              // [n_entity, D] (+=)
  float* du;               // [B, D] (+=)
  float* ds;               // [n_rel] (+=) softmax gradient of the PARENTS' attention
  long rows;
  int rpp;                 // rows per pair
  unsigned long long rpp_magic;
  int K, kshift, n_rel, stream;   // K = 1 << kshift
};

template <int D>
inline size_t leaf_bwd_tc_smem(int n_rel) {
  return 2 * TB<D>::ATILE + 2 * TB<D>::WTILE + sizeof(float) * (3 * TB<D>::R + 2 * D + (size_t)n_rel * (1 + TB<D>::NW)) +
         sizeof(int) * TB<D>::R + 64;
}

template <int D>
__global__ void __launch_bounds__(TB<D>::NT, 2) agg_bwd_leaf_tc_kernel(LeafBwdArgs a) {
  pdl_enter();
  using T_ = TB<D>;
  using LA = typename T_::LA;
  constexpr int NP = T_::NP, APL = T_::APL, WPL = T_::WPL, HC = T_::HC;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* a1 = smem_raw;                           // gz -> grow
  unsigned char* a2 = a1 + T_::ATILE;                     // Y -> SU
  unsigned char* wa_t = a2 + T_::ATILE;
  unsigned char* wt_t = wa_t + T_::WTILE;
  float* pk_s = reinterpret_cast<float*>(wt_t + T_::WTILE);   // [R] p_k of each row within its family
  float* dp_s = pk_s + T_::R;                             // [2][R] partial dp (column halves)
  float* bias_s = dp_s + 2 * T_::R;                       // [2][D] dba partial (second half unused)
  float* s_s = bias_s + 2 * D;                            // [n_rel]
  float* ds_s = s_s + a.n_rel;                            // [NW][n_rel]
  int* rel_s = reinterpret_cast<int*>(ds_s + T_::NW * a.n_rel);   // [R]
  uint64_t* bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(rel_s + T_::R) + 7) & ~(uintptr_t)7);   // the float / int
                                                          // arrays before it have an n_rel-dependent length
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, tx = tid % T_::LPR, ty = tid / T_::LPR;
  const int K = a.K;
  if (warp == 0) umma::tmem_alloc(tmem_slot, 256);        // [0,64) dX accumulator, [64,128) dW_a, [128,192) dW_t
  if (tid == 32) {
    umma::mbar_init(bar, 1);
    umma::fence_barrier_init();
  }
  umma::stage_weight_bf<D, NP, 128>(wa_t, a.Wa, D, tid, T_::NT);
  umma::stage_weight_bf<D, NP, 128>(wt_t, a.Wt, D, tid, T_::NT);
  for (int i = tid; i < a.n_rel; i += T_::NT) s_s[i] = a.s[i];
  for (int i = tid; i < T_::NW * a.n_rel; i += T_::NT) ds_s[i] = 0.f;
  for (int i = tid; i < 2 * D; i += T_::NT) bias_s[i] = 0.f;
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_x = tmem, tm_wa = tmem + 64, tm_wt = tmem + 128;
  uint32_t phase = 0;
  const int er = 32 * (warp % 4) + lane, c0 = (warp / 4) * HC;     // accumulator mapping: row er, columns c0 .. c0 + HC
  const uint32_t tlane = (uint32_t)(32 * (warp % 4)) << 16;
  const float invK = 1.f / (float)K;
  float* ds_w = ds_s + warp * a.n_rel;
  float4 bpa = f4zero();                                  // dba partial, row mapping (columns 4 tx ..)
  const long ntiles = (a.rows + T_::R - 1) / T_::R;
  bool first = true;

  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long row0 = t * T_::R;
    // next tile's four streams -> L2 through the bulk-copy engine (full tiles only: the range stays inside the buffers)
    if (tid < 4) {
      const long nrow0 = (t + gridDim.x) * T_::R;
      if (nrow0 + T_::R <= a.rows) {
        const float* base = tid == 0 ? a.g1 : tid == 1 ? a.V1 : tid == 2 ? a.Y : a.T;
        bulk_prefetch_l2(base + nrow0 * D, T_::R * D * sizeof(float));
      }
    }
    // ---- family records: p_k and rel_k of every row of the tile (lanes of a warp = children of one parent) ----
    for (int f = warp; f < T_::R / K; f += T_::NW) {
      const long par = (row0 >> a.kshift) + f;
      if ((par << a.kshift) < a.rows) {                     // warp-uniform
        const int32_t* arow = a.adj + (long)__ldg(a.ent_par + par) * 2 * K + K;
        const int r0 = lane < K ? __ldg(arow + lane) : 0, r1 = lane + 32 < K ? __ldg(arow + lane + 32) : 0;
        const float l0 = lane < K ? s_s[r0] : -INFINITY, l1 = lane + 32 < K ? s_s[r1] : -INFINITY;
        const float mx = warp_max(fmaxf(l0, l1));
        const float e0 = lane < K ? expf(l0 - mx) : 0.f, e1 = lane + 32 < K ? expf(l1 - mx) : 0.f;
        const float inv = 1.f / warp_sum(e0 + e1);
        if (lane < K) { pk_s[f * K + lane] = e0 * inv; rel_s[f * K + lane] = r0; }
        if (lane + 32 < K) { pk_s[f * K + lane + 32] = e1 * inv; rel_s[f * K + lane + 32] = r1; }
      }
    }
    // ---- phase 0 (row mapping): gz = g1 * [V1 > 0] and Y as split-bf16 planes ----
#pragma unroll
    for (int h = 0; h < T_::NPASS; h += 4) {
      float4 g[4], v[4], y[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const long row = row0 + (h + q) * T_::RP + ty;
        g[q] = v[q] = y[q] = f4zero();
        if (row < a.rows) {
          g[q] = ld4a(a.g1 + row * D + tx * 4, a.stream);
          v[q] = ld4a(a.V1 + row * D + tx * 4, a.stream);
          y[q] = ld4a(a.Y + row * D + tx * 4, a.stream);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = (h + q) * T_::RP + ty;
        const float4 gz = make_float4(v[q].x > 0.f ? g[q].x : 0.f, v[q].y > 0.f ? g[q].y : 0.f, v[q].z > 0.f ? g[q].z : 0.f,
                                      v[q].w > 0.f ? g[q].w : 0.f);
        bpa = f4add(bpa, gz);
        umma::store_planes4<NP>(a1, APL, LA::off4(r, tx), gz);
        umma::store_planes4<NP>(a2, APL, LA::off4(r, tx), y[q]);
      }
    }
    umma::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_sync();
      umma::issue_dx_bf<D, NP, 144, 128>(tm_x, umma::smem_u32(a1), APL, umma::smem_u32(wa_t), WPL, D, true);       // gs
      umma::issue_dw_bf<D, D, NP, 144>(tm_wa, umma::smem_u32(a2), APL, umma::smem_u32(a1), APL, T_::R, first);     // dW_a
      umma::commit(bar);
    }
    // ---- while the tensor core works: S + u of the tile's rows (row mapping), kept in registers ----
    float4 su[T_::NPASS];
    {
      long e[T_::NPASS];
#pragma unroll
      for (int ps = 0; ps < T_::NPASS; ++ps) {
        const long row = row0 + ps * T_::RP + ty;
        e[ps] = row < a.rows ? (long)__ldg(a.ent + row) : -1;
      }
#pragma unroll
      for (int ps = 0; ps < T_::NPASS; ++ps) {
        const long row = row0 + ps * T_::RP + ty;
        su[ps] = f4zero();
        if (e[ps] >= 0)
          su[ps] = f4add(ldg4(a.Se + e[ps] * D + tx * 4), ldg4(a.u + fastdiv(row, a.rpp_magic) * D + tx * 4));
      }
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
    // ---- phase 1 (accumulator mapping): gs -> grow planes, dT = gs + p_k grow_parent, dp_k = grow_parent . T ----
    {
      const long row = row0 + er;
      const bool valid = row < a.rows;
      const float pk = valid ? pk_s[er] : 0.f;
      const float* gp = a.GP + (valid ? row >> a.kshift : 0) * D + c0;
      const float* tr = a.T + (valid ? row : 0) * D + c0;
      float* dt = a.dT + (valid ? row : 0) * D + c0;
      float dpart = 0.f;
#pragma unroll
      for (int cc = 0; cc < HC; cc += 16) {
        float gs[16];
        umma::tmem_ld16(tm_x + tlane + (uint32_t)(c0 + cc), gs);
        float4 gpv[4], tv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          gpv[j] = valid ? ldg4(gp + cc + 4 * j) : f4zero();
          tv[j] = valid ? ld4a(tr + cc + 4 * j, a.stream) : f4zero();
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 g4 = make_float4(gs[4 * j], gs[4 * j + 1], gs[4 * j + 2], gs[4 * j + 3]);
          const float4 grow = f4scale(g4, invK);
          if (valid) st4a(dt + cc + 4 * j, f4fma(pk, gpv[j], g4), a.stream);
          dpart += f4dot(gpv[j], tv[j]);
          // grow planes (K-major + MN-major operand of the second round): features c0 + cc + 4 j .. + 3 of row er
          umma::store_planes4<NP>(a1, APL, LA::off4(er, (c0 + cc) / 4 + j), grow);
        }
      }
      dp_s[(warp / 4) * T_::R + er] = dpart;
    }
    // S + u planes (row mapping)
#pragma unroll
    for (int ps = 0; ps < T_::NPASS; ++ps) umma::store_planes4<NP>(a2, APL, LA::off4(ps * T_::RP + ty, tx), su[ps]);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_sync();
      umma::issue_dx_bf<D, NP, 144, 128>(tm_x, umma::smem_u32(a1), APL, umma::smem_u32(wt_t), WPL, D, true);       // gsu
      umma::issue_dw_bf<D, D, NP, 144>(tm_wt, umma::smem_u32(a2), APL, umma::smem_u32(a1), APL, T_::R, first);     // dW_t[L]
      umma::commit(bar);
    }
    first = false;
    // ---- while the tensor core works: the parents' softmax gradient (warps 0-3, lane = row) ----
    if (warp < 4) {
      const long row = row0 + er;
      const bool valid = row < a.rows;
      const float pk = valid ? pk_s[er] : 0.f;
      const float dp = (dp_s[er] + dp_s[T_::R + er]) * 1.f;      // dL/dp_k = grow_parent . T_k   (grow already holds 1/K)
      float dot = pk * dp;
      for (int o = 1; o < K && o < 32; o <<= 1) dot += __shfl_xor_sync(FULL_MASK, dot, o);
      if (K > 32) {                                         // a family spans two warps: exchange through shared memory
        __syncwarp();
        dp_s[er] = dot;                                     // (dp_s[0..R) is no longer needed by this warp's rows)
      }
      // named barrier among warps 0-3 only when K > 32
      if (K > 32) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        dot = dp_s[er & ~63] + dp_s[(er & ~63) + 32];
      }
      if (valid) atomicAdd(&ds_w[rel_s[er]], pk * (dp - dot));
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
    // ---- phase 2 (accumulator mapping): gsu -> GSe[entity] (+=), du[pair] (+=) ----
    {
      const long row = row0 + er;
      const bool valid = row < a.rows;
      float* gse = a.GSe + (valid ? (long)__ldg(a.ent + row) : 0) * D + c0;
      const bool whole = (a.rpp % 32) == 0;                 // a warp's 32 rows belong to one pair
      const long pair = fastdiv(valid ? row : row0, a.rpp_magic);
#pragma unroll
      for (int cc = 0; cc < HC; cc += 16) {
        float v[16];
        umma::tmem_ld16(tm_x + tlane + (uint32_t)(c0 + cc), v);
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) red_add4(gse + cc + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
        }
        if (whole) {
          if (!valid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.f;
          }
          const float sum = warp_colsum16(v, lane);
          const bool any = row0 + 32 * (warp % 4) < a.rows;
          if (lane < 16 && any) atomicAdd(a.du + fastdiv(row0 + 32 * (warp % 4), a.rpp_magic) * D + c0 + cc + lane, sum);
        } else if (valid) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) red_add4(a.du + pair * D + c0 + cc + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
        }
      }
    }
    umma::fence_before_sync();
    __syncthreads();                                        // planes and family records are rewritten by the next tile
    umma::fence_after_sync();
  }
  // ---- epilogue: bias partials, relation-score gradients, weight gradients ----
  {
    // dba: row mapping -> column sums over the ty of a warp, then shared memory
    float4 p = cross_group_sum4<T_::LPR>(bpa);
    if (lane / T_::LPR == 0) {
      atomicAdd(&bias_s[tx * 4 + 0], p.x); atomicAdd(&bias_s[tx * 4 + 1], p.y);
      atomicAdd(&bias_s[tx * 4 + 2], p.z); atomicAdd(&bias_s[tx * 4 + 3], p.w);
    }
  }
  __syncthreads();
  // db_t[L] = sum_rows grow = (sum_rows gz) . W_a^T / K: linear in the column sums this CTA already holds
  if (tid < D) {
    const float4* wrow = reinterpret_cast<const float4*>(a.Wa + (long)tid * D);
    float acc = 0.f;
#pragma unroll 4
    for (int n = 0; n < D / 4; ++n) acc += f4dot(__ldg(wrow + n), ld4(&bias_s[4 * n]));
    if (bias_s[tid] != 0.f) atomicAdd(a.dba + tid, bias_s[tid]);
    if (acc != 0.f) atomicAdd(a.dbt + tid, acc * invK);
  }
  for (int i = tid; i < a.n_rel; i += T_::NT) {
    float s = 0.f;
    for (int w = 0; w < T_::NW; ++w) s += ds_s[w * a.n_rel + i];
    if (s != 0.f) atomicAdd(a.ds + i, s);
  }
  if (!first) {                                             // this CTA processed at least one tile
    flush_dw_tmem<D>(tm_wa, a.dWa, warp, lane);
    flush_dw_tmem<D>(tm_wt, a.dWt, warp, lane);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------------------
// user-oriented transform, backward, on the tensor cores:  dT = g1 (+ g2);  dW_t[h] += XU^T dT, XU = E[ent] + u;
// db += sum dT;  gx = dT . W_t[h]^T;  dE[ent] += gx;  du[b] += sum_rows gx        (see transform_bwd_kernel)
// TransformLevel::W holds W_t[h] as stored ([d, d] row-major), not its transpose.
// ---------------------------------------------------------------------------------------------------------
template <int D>
inline size_t transform_bwd_tc_smem() { return 2 * TB<D>::ATILE + TB<D>::WTILE + sizeof(float) * D + 64; }

template <int D>
__global__ void __launch_bounds__(TB<D>::NT, 2) transform_bwd_tc_kernel(TransformArgs a) {
  pdl_enter();
  using T_ = TB<D>;
  using LA = typename T_::LA;
  constexpr int NP = T_::NP, APL = T_::APL, WPL = T_::WPL, HC = T_::HC;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* a1 = smem_raw;                           // dT
  unsigned char* a2 = a1 + T_::ATILE;                     // XU
  unsigned char* w_t = a2 + T_::ATILE;
  float* bias_s = reinterpret_cast<float*>(w_t + T_::WTILE);   // [D]
  uint64_t* bar = reinterpret_cast<uint64_t*>(bias_s + D);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, tx = tid % T_::LPR, ty = tid / T_::LPR;
  const CtaSlice cs = cta_slice(a.cta_end, a.nlev);
  const TransformLevel& L = a.lv[cs.level];
  if (warp == 0) umma::tmem_alloc(tmem_slot, 128);        // [0,64) gx accumulator, [64,128) dW_t[h]
  if (tid == 32) {
    umma::mbar_init(bar, 1);
    umma::fence_barrier_init();
  }
  umma::stage_weight_bf<D, NP, 128>(w_t, L.W, D, tid, T_::NT);
  for (int i = tid; i < D; i += T_::NT) bias_s[i] = 0.f;
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_x = tmem, tm_w = tmem + 64;
  uint32_t phase = 0;
  const int er = 32 * (warp % 4) + lane, c0 = (warp / 4) * HC;
  const uint32_t tlane = (uint32_t)(32 * (warp % 4)) << 16;
  float4 bpart = f4zero();
  const long ntiles = (L.rows + T_::R - 1) / T_::R;
  const bool whole = (L.rpp % 32) == 0;
  bool first = true;
  for (long t = cs.local; t < ntiles; t += cs.count) {
    const long row0 = t * T_::R;
    if (tid < 2) {                                          // next tile's gradient rows -> L2 (bulk-copy engine)
      const long nrow0 = (t + cs.count) * T_::R;
      const float* base = tid == 0 ? L.g1 : L.g2;
      if (base && nrow0 + T_::R <= L.rows) bulk_prefetch_l2(base + nrow0 * D, T_::R * D * sizeof(float));
    }
#pragma unroll
    for (int h = 0; h < T_::NPASS; h += 4) {
      float4 g[4], x[4], uu[4];
      long e[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const long row = row0 + (h + q) * T_::RP + ty;
        e[q] = row < L.rows ? (long)__ldg(L.ent + row) : -1;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const long row = row0 + (h + q) * T_::RP + ty;
        g[q] = x[q] = uu[q] = f4zero();
        if (e[q] >= 0) {
          g[q] = ld4a(L.g1 + row * D + tx * 4, L.stream);
          if (L.g2) g[q] = f4add(g[q], ld4a(L.g2 + row * D + tx * 4, L.stream));
          x[q] = ldg4(erow(a.E, e[q], D) + tx * 4);
          uu[q] = ldg4(a.u + fastdiv(row, L.rpp_magic) * D + tx * 4);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = (h + q) * T_::RP + ty;
        bpart = f4add(bpart, g[q]);
        umma::store_planes4<NP>(a1, APL, LA::off4(r, tx), g[q]);
        umma::store_planes4<NP>(a2, APL, LA::off4(r, tx), f4add(x[q], uu[q]));
      }
    }
    umma::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_sync();
      umma::issue_dx_bf<D, NP, 144, 128>(tm_x, umma::smem_u32(a1), APL, umma::smem_u32(w_t), WPL, D, true);        // gx
      umma::issue_dw_bf<D, D, NP, 144>(tm_w, umma::smem_u32(a2), APL, umma::smem_u32(a1), APL, T_::R, first);      // dW_t[h]
      umma::commit(bar);
    }
    first = false;
    const long row = row0 + er;
    const bool valid = row < L.rows;
    float* ge = valid ? grow_of(a.dE, __ldg(L.ent + row), D) + c0 : nullptr;
    const long pair = fastdiv(valid ? row : row0, L.rpp_magic);
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
#pragma unroll
    for (int cc = 0; cc < HC; cc += 16) {
      float v[16];
      umma::tmem_ld16(tm_x + tlane + (uint32_t)(c0 + cc), v);
      if (valid) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) red_add4(ge + cc + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
      }
      if (whole) {
        if (!valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
        const float sum = warp_colsum16(v, lane);
        const bool any = row0 + 32 * (warp % 4) < L.rows;
        if (lane < 16 && any) atomicAdd(a.du + fastdiv(row0 + 32 * (warp % 4), L.rpp_magic) * D + c0 + cc + lane, sum);
      } else if (valid) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) red_add4(a.du + pair * D + c0 + cc + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
      }
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }
  {
    float4 p = cross_group_sum4<T_::LPR>(bpart);
    if (lane / T_::LPR == 0) {
      atomicAdd(&bias_s[tx * 4 + 0], p.x); atomicAdd(&bias_s[tx * 4 + 1], p.y);
      atomicAdd(&bias_s[tx * 4 + 2], p.z); atomicAdd(&bias_s[tx * 4 + 3], p.w);
    }
  }
  __syncthreads();
  if (tid < D && bias_s[tid] != 0.f) atomicAdd(L.db + tid, bias_s[tid]);
  if (!first) flush_dw_tmem<D>(tm_w, L.dW, warp, lane);
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 128);
}

}  // namespace mvin
