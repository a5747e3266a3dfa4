// level_tc.cuh -- tcgen05 (5th-generation tensor core) versions of the forward row kernels of level.cuh for
// d in {32, 64}: same reference semantics (model.py:267-283 user-oriented transform; aggregators.py:98-146 as driven
// by model.py:286-307), same argument structs, same thread-mapped neighbour phase -- but tiles of 128 rows whose
// d x d maps run as 3xTF32 tcgen05.mma from shared-memory operands into a tensor-memory accumulator (umma.cuh)
// instead of mma.sync fragments.  On the legacy path the three-product split made the dense phase, not the row
// streaming, the bound of these kernels at d = 64 (profiles/README.md); here it is a few dozen instructions issued by
// one thread per tile.
//
// Per tile: all 256 threads gather / aggregate the tile's rows (thread-mapped, LPR = D/4 lanes per row) and store
// them, split into TF32 hi / lo parts, as the K-major A operand; one thread issues the MMAs and commits to an
// mbarrier; everybody waits, then reads the accumulator back with tcgen05.ld -- warp w owns tensor-memory lanes
// 32 (w % 4) .. +31 = tile rows, warps 0-3 take the first D/2 columns and warps 4-7 the rest -- and runs the
// epilogue (bias, residual, ReLU, global stores).  Overlap between the phases comes from the other resident CTAs.
#pragma once
#include "level.cuh"
#include "umma.cuh"

namespace mvin {

template <int D>
struct TT {
  static constexpr int NT = 256, NW = NT / 32, R = 128;
  static constexpr int LPR = D / 4;                  // float4 lanes per row
  static constexpr int RP = NT / LPR;                // rows per thread-mapped pass
  static constexpr int NP = R / RP;                  // passes per tile
  static constexpr int HC = D / 2;                   // accumulator columns per epilogue thread
  static constexpr int TCOLS = D < 32 ? 32 : D;      // tensor-memory columns of the accumulator
  static constexpr int TILE = umma::OpLayout<D>::bytes(R), WTILE = umma::OpLayout<D>::bytes(D);
  static constexpr int SLD = D + 4;                  // row pitch (floats) of the epilogue staging tile
  static_assert(R * SLD * 4 <= TILE, "staging tile must fit the A_hi buffer");
};

// smem operands are complete -> MMAs -> accumulator complete.  Every thread of the CTA calls it.
template <int D>
MVIN_DEV void mma_round(uint32_t tmem, unsigned char* a_hi, unsigned char* a_lo, unsigned char* w_hi, unsigned char* w_lo,
                        uint64_t* bar, uint32_t& phase, int tid) {
  umma::fence_async_smem();
  __syncthreads();
  if (tid == 0) {
    umma::fence_after_sync();
    umma::issue_3xtf32<D>(tmem, umma::smem_u32(a_hi), umma::smem_u32(a_lo), umma::smem_u32(w_hi), umma::smem_u32(w_lo), D,
                          true);
    umma::commit(bar);
  }
  umma::mbar_wait(bar, phase);
  phase ^= 1;
  umma::fence_after_sync();
}

// stage phase for a 128-row tile: warp w handles rows w, w + 8, ... four at a time (see stage_tile in level.cuh)
template <bool WITH_REL>
MVIN_DEV void stage_rows128(const int32_t* __restrict__ ent, const int32_t* __restrict__ adj,
                            const float* __restrict__ s_s, long row0, long rows, int K, int KP, int2* __restrict__ nb_s,
                            uint16_t* __restrict__ rel_s, int warp, int lane) {
  constexpr int CH = 4, RPW = 16;
#pragma unroll 1
  for (int j0 = 0; j0 < RPW; j0 += CH) {
    int id0[CH], id1[CH], rl0[CH], rl1[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const long row = row0 + warp * RPW + j0 + j;
      id0[j] = id1[j] = rl0[j] = rl1[j] = 0;
      if (row < rows) {
        const int32_t* arow = adj + (long)__ldg(ent + row) * 2 * K;
        if (lane < K) { id0[j] = __ldg(arow + lane); rl0[j] = __ldg(arow + K + lane); }
        if (lane + 32 < K) { id1[j] = __ldg(arow + lane + 32); rl1[j] = __ldg(arow + K + lane + 32); }
      }
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int r = warp * RPW + j0 + j;
      if (row0 + r >= rows) continue;                       // warp-uniform
      const float l0 = lane < K ? s_s[rl0[j]] : -INFINITY;
      const float l1 = lane + 32 < K ? s_s[rl1[j]] : -INFINITY;
      const float mx = warp_max(fmaxf(l0, l1));
      const float e0 = lane < K ? expf(l0 - mx) : 0.f;
      const float e1 = lane + 32 < K ? expf(l1 - mx) : 0.f;
      const float inv = 1.f / warp_sum(e0 + e1);
      if (lane < K) {
        nb_s[r * KP + lane] = make_int2(__float_as_int(e0 * inv), id0[j]);
        if (WITH_REL) rel_s[r * KP + lane] = (uint16_t)rl0[j];
      }
      if (lane + 32 < K) {
        nb_s[r * KP + lane + 32] = make_int2(__float_as_int(e1 * inv), id1[j]);
        if (WITH_REL) rel_s[r * KP + lane + 32] = (uint16_t)rl1[j];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// user-oriented transform  T = (E[ent] + u) . W_t[h] + b_t[h]   (model.py:270-283), levels h < L, one launch
// ---------------------------------------------------------------------------------------------------------
template <int D>
inline size_t transform_fwd_tc_smem() { return 2 * TT<D>::TILE + 2 * TT<D>::WTILE + 32; }

template <int D>
__global__ void __launch_bounds__(TT<D>::NT) transform_fwd_tc_kernel(TransformArgs a) {
  pdl_enter();
  using T = TT<D>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* a_hi = smem_raw;
  unsigned char* a_lo = a_hi + T::TILE;
  unsigned char* w_hi = a_lo + T::TILE;
  unsigned char* w_lo = w_hi + T::WTILE;
  uint64_t* bar = reinterpret_cast<uint64_t*>(w_lo + T::WTILE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, tx = tid % T::LPR, ty = tid / T::LPR;
  const CtaSlice cs = cta_slice(a.cta_end, a.nlev);
  const TransformLevel& L = a.lv[cs.level];
  if (warp == 0) umma::tmem_alloc(tmem_slot, T::TCOLS);
  if (tid == 32) {
    umma::mbar_init(bar, 1);
    umma::fence_barrier_init();
  }
  umma::stage_weight_t<D>(w_hi, w_lo, L.W, D, tid, T::NT);   // operand (n, k) = W[k][n]
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  uint32_t phase = 0;
  const int er = 32 * (warp % 4) + lane, c0 = (warp / 4) * T::HC;
  const uint32_t tbase = tmem + ((uint32_t)(32 * (warp % 4)) << 16) + (uint32_t)c0;

  // software pipeline: the gather of tile t + 1 is issued right after the MMAs of tile t, so its latency overlaps the
  // tensor-core round trip and the epilogue
  const long ntiles = (L.rows + T::R - 1) / T::R;
  float4 x[T::NP];
  auto gather = [&](long tile) {
    // three separate loops so that all NP independent (id -> row) chains are in flight together
    long e[T::NP];
    float4 xe[T::NP];
#pragma unroll
    for (int ps = 0; ps < T::NP; ++ps) {
      const long row = tile * T::R + ps * T::RP + ty;
      e[ps] = (tile < ntiles && row < L.rows) ? (long)__ldg(L.ent + row) : -1;
    }
#pragma unroll
    for (int ps = 0; ps < T::NP; ++ps) {
      const long row = tile * T::R + ps * T::RP + ty;
      xe[ps] = x[ps] = f4zero();
      if (e[ps] >= 0) {
        xe[ps] = ldg4(erow(a.E, e[ps], D) + tx * 4);
        x[ps] = ldg4(a.u + fastdiv(row, L.rpp_magic) * D + tx * 4);
      }
    }
#pragma unroll
    for (int ps = 0; ps < T::NP; ++ps) x[ps] = f4add(x[ps], xe[ps]);
  };
  gather(cs.local);
  for (long t = cs.local; t < ntiles; t += cs.count) {
    const long row0 = t * T::R;
#pragma unroll
    for (int ps = 0; ps < T::NP; ++ps) umma::store_split<D>(a_hi, a_lo, ps * T::RP + ty, tx, x[ps]);
    umma::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_sync();
      umma::issue_3xtf32<D>(tmem, umma::smem_u32(a_hi), umma::smem_u32(a_lo), umma::smem_u32(w_hi), umma::smem_u32(w_lo), D,
                            true);
      umma::commit(bar);
    }
    gather(t + cs.count);
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
    // epilogue: accumulator rows (thread = row) -> row-major staging tile in the A_hi buffer (free once the MMAs have
    // completed) -> coalesced 16-byte stores, LPR lanes per row
    float* stg = reinterpret_cast<float*>(a_hi);
#pragma unroll
    for (int cc = 0; cc < T::HC; cc += 16) {
      float v[16];
      umma::tmem_ld16(tbase + cc, v);
#pragma unroll
      for (int j = 0; j < 16; j += 4) st4(&stg[er * T::SLD + c0 + cc + j], make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
    }
    umma::fence_before_sync();
    __syncthreads();
    const float4 b = ldg4(L.b + tx * 4);
#pragma unroll
    for (int ps = 0; ps < T::NP; ++ps) {
      const int r = ps * T::RP + ty;
      const long row = row0 + r;
      if (row < L.rows) st4a(L.T + row * D + tx * 4, f4add(ld4(&stg[r * T::SLD + tx * 4]), b), L.stream);
    }
    __syncthreads();                                      // staging tile is the next A operand
  }
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, T::TCOLS);
}

// ---------------------------------------------------------------------------------------------------------
// one aggregator iteration, forward, all levels  (aggregators.py:98-146; model.py:295-306) -- see agg_fwd_kernel
// ---------------------------------------------------------------------------------------------------------
template <int D>
inline size_t agg_fwd_tc_smem(bool has_leaf, int K, int n_rel) {
  return 2 * TT<D>::TILE + (has_leaf ? 4 : 2) * TT<D>::WTILE + AggSmem<D>::align16(sizeof(int2) * TT<D>::R * (K | 1)) +
         AggSmem<D>::align16(sizeof(float) * n_rel) + 32;
}

template <int D, bool HAS_LEAF>
__global__ void __launch_bounds__(TT<D>::NT) agg_fwd_tc_kernel(AggArgs a) {
  pdl_enter();
  using T = TT<D>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int K = a.K, KP = padded_k(K);
  unsigned char* a_hi = smem_raw;
  unsigned char* a_lo = a_hi + T::TILE;
  unsigned char* wa_hi = a_lo + T::TILE;
  unsigned char* wa_lo = wa_hi + T::WTILE;
  unsigned char* wt_hi = wa_lo + T::WTILE;
  unsigned char* wt_lo = wt_hi + (HAS_LEAF ? T::WTILE : 0);
  int2* nb_s = reinterpret_cast<int2*>(wt_lo + (HAS_LEAF ? T::WTILE : 0));
  float* s_s = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(nb_s) +
                                        AggSmem<D>::align16(sizeof(int2) * T::R * KP));
  uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_s) +
                                              AggSmem<D>::align16(sizeof(float) * a.n_rel));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  int* sched = reinterpret_cast<int*>(tmem_slot + 1);        // [2]
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, tx = tid % T::LPR, ty = tid / T::LPR;
  if (tid == 0) sched[0] = sched_next(a.tl);
  if (warp == 0) umma::tmem_alloc(tmem_slot, T::TCOLS);
  if (tid == 32) {
    umma::mbar_init(bar, 1);
    umma::fence_barrier_init();
  }
  umma::stage_weight_t<D>(wa_hi, wa_lo, a.Wa, D, tid, T::NT);
  if (HAS_LEAF) umma::stage_weight_t<D>(wt_hi, wt_lo, a.Wt, D, tid, T::NT);
  for (int i = tid; i < a.n_rel; i += T::NT) s_s[i] = a.s[i];
  const float invK = 1.f / (float)K;
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  uint32_t phase = 0;
  const int er = 32 * (warp % 4) + lane, c0 = (warp / 4) * T::HC;
  const uint32_t tbase = tmem + ((uint32_t)(32 * (warp % 4)) << 16) + (uint32_t)c0;

  for (int it = 0;; ++it) {
    const long t = sched[it & 1];
    if (t >= a.tl.tile_end[a.tl.nlev - 1]) break;
    int nxt = 0;
    if (tid == 0) nxt = sched_next(a.tl);                    // consumed at the end of this tile
    int lvl = 0;
    while (t >= a.tl.tile_end[lvl]) ++lvl;
    const AggLevel& L = a.lv[lvl];
    const bool leaf = HAS_LEAF && L.leaf;
    const bool ent_mode = leaf && a.Se != nullptr;
    const long row0 = (t - (lvl ? a.tl.tile_end[lvl - 1] : 0)) * T::R;
    // ---- stage phase ----
    if (!ent_mode) {
      stage_rows128<false>(L.ent, a.adj, s_s, row0, L.rows, K, KP, nb_s, nullptr, warp, lane);
      __syncthreads();
    }
    // ---- neighbour phase: thread-mapped, result = A operand of the first map ----
    if (ent_mode) {
      // per-entity leaf mode: S_e was computed once per distinct entity; all NP (id -> row) chains in flight together
      long e[T::NP];
      float4 xs[T::NP], xu[T::NP];
#pragma unroll
      for (int ps = 0; ps < T::NP; ++ps) {
        const long row = row0 + ps * T::RP + ty;
        e[ps] = row < L.rows ? (long)__ldg(L.ent + row) : -1;
      }
#pragma unroll
      for (int ps = 0; ps < T::NP; ++ps) {
        const long row = row0 + ps * T::RP + ty;
        xs[ps] = xu[ps] = f4zero();
        if (e[ps] >= 0) {
          xs[ps] = ldg4(a.Se + e[ps] * D + tx * 4);
          xu[ps] = ldg4(a.u + fastdiv(row, L.rpp_magic) * D + tx * 4);
        }
      }
#pragma unroll
      for (int ps = 0; ps < T::NP; ++ps) {
        const float4 o = f4add(xs[ps], xu[ps]);                // S + u: recomputed by the backward, not stored
        umma::store_split<D>(a_hi, a_lo, ps * T::RP + ty, tx, o);
      }
    } else {
#pragma unroll 1
      for (int ps = 0; ps < T::NP; ++ps) {
        const int r = ps * T::RP + ty;
        const long row = row0 + r;
        float4 o = f4zero();
        if (row < L.rows) {
          const int2* nb = nb_s + r * KP;
          float4 acc = f4zero();
          if (leaf) {
            const float4 uv = ldg4(a.u + fastdiv(row, L.rpp_magic) * D + tx * 4);
#pragma unroll 8
            for (int k = 0; k < K; ++k) {
              const int2 v = nb[k];
              acc = f4fma(__int_as_float(v.x), ldg4(erow(a.E, v.y, D) + tx * 4), acc);
            }
            o = f4add(acc, uv);
            st4a(L.SU + row * D + tx * 4, o, L.stream);
          } else {
            const float4 sv = ld4a(L.self + row * D + tx * 4, L.stream);
            const float* base = L.child + row * K * D + tx * 4;
#pragma unroll 8
            for (int k = 0; k < K; ++k) acc = f4fma(__int_as_float(nb[k].x), ld4a(base + (long)k * D, L.stream), acc);
            o = f4fma(invK, acc, sv);
            st4a(L.Y + row * D + tx * 4, o, L.stream);
          }
        }
        umma::store_split<D>(a_hi, a_lo, r, tx, o);
      }
    }
    float* stg = reinterpret_cast<float*>(a_hi);             // epilogue staging tile (A_hi is free after the MMAs)
    if (leaf) {
      // agg = ((S + u) . W_t[L] + b_t[L]) / K ;  Y = self + agg  -> A operand of the second map
      float4 sv[T::NP];
#pragma unroll
      for (int ps = 0; ps < T::NP; ++ps) {                   // residual rows: in flight during the MMA round trip
        const long row = row0 + ps * T::RP + ty;
        sv[ps] = row < L.rows ? ld4a(L.self + row * D + tx * 4, L.stream) : f4zero();
      }
      mma_round<D>(tmem, a_hi, a_lo, wt_hi, wt_lo, bar, phase, tid);
#pragma unroll
      for (int cc = 0; cc < T::HC; cc += 16) {
        float v[16];
        umma::tmem_ld16(tbase + cc, v);
#pragma unroll
        for (int j = 0; j < 16; j += 4) st4(&stg[er * T::SLD + c0 + cc + j], make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
      }
      umma::fence_before_sync();
      __syncthreads();
      const float4 b = ldg4(a.bt + tx * 4);
      float4 y[T::NP];
#pragma unroll
      for (int ps = 0; ps < T::NP; ++ps) {
        const int r = ps * T::RP + ty;
        const long row = row0 + r;
        y[ps] = f4zero();
        if (row < L.rows) {
          y[ps] = f4fma(invK, f4add(ld4(&stg[r * T::SLD + tx * 4]), b), sv[ps]);
          st4a(L.Y + row * D + tx * 4, y[ps], L.stream);
        }
      }
      __syncthreads();                                       // everybody has read the staging tile
#pragma unroll
      for (int ps = 0; ps < T::NP; ++ps) umma::store_split<D>(a_hi, a_lo, ps * T::RP + ty, tx, y[ps]);
    }
    // V = relu(Y . W_a + b_a)
    mma_round<D>(tmem, a_hi, a_lo, wa_hi, wa_lo, bar, phase, tid);
#pragma unroll
    for (int cc = 0; cc < T::HC; cc += 16) {
      float v[16];
      umma::tmem_ld16(tbase + cc, v);
#pragma unroll
      for (int j = 0; j < 16; j += 4) st4(&stg[er * T::SLD + c0 + cc + j], make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
    }
    umma::fence_before_sync();
    __syncthreads();
    {
      const float4 b = ldg4(a.ba + tx * 4);
#pragma unroll
      for (int ps = 0; ps < T::NP; ++ps) {
        const int r = ps * T::RP + ty;
        const long row = row0 + r;
        if (row < L.rows) {
          const float4 z = f4add(ld4(&stg[r * T::SLD + tx * 4]), b);
          st4a(L.V + row * D + tx * 4, make_float4(fmaxf(z.x, 0.f), fmaxf(z.y, 0.f), fmaxf(z.z, 0.f), fmaxf(z.w, 0.f)), L.stream);
        }
      }
    }
    if (tid == 0) sched[(it + 1) & 1] = nxt;
    __syncthreads();
  }
  sched_exit(a.tl);
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, T::TCOLS);
}

}  // namespace mvin
