// level.cuh -- the KG side of the MVIN hot path: fused gather-attend-aggregate kernels, one launch per
// aggregator iteration covering EVERY level of that iteration.
//
// Replaces (reference, src/model/MVIN/): model.py:267-283 (entity gather + user-oriented transform),
// aggregators.py:98-146 (attention over K, mean of weighted neighbours, (self+agg).W+b, ReLU) as driven by the
// loop at model.py:286-307, and the TF autodiff of the same ops (weight gradients included: dW is accumulated in
// registers by the CTA that already holds the tile, never by a second pass over the activations).
//
// Layout: every activation buffer is row-major [rows, D] fp32 with rows = B * K^h (pair-major, child k of node
// j at row j*K+k, model.py:251).  A launch carries up to MAX_LV level descriptors; the grid is partitioned
// between the levels (cta_end[]), each CTA owns tiles of R = 64 consecutive rows of ONE level and walks them
// persistently; the d x d weights live in shared memory for the CTA's lifetime.  Two thread mappings per tile:
//   * "warp per row" for the neighbour phase: a warp reads the node's packed adjacency record (K ids + K
//     relation ids, contiguous), does the K-softmax with shuffles, then streams the K child rows with G = 32/LPR
//     rows in flight per load instruction (LPR = D/4 lanes x 16 B cover one row).
//   * "register tile" for the dense maps: thread (ty, tx) owns TM rows x 4 columns, FP32 FFMA (see gemm.cuh for
//     why not TF32), and TMW x 4 entries of each weight gradient.
#pragma once
#include "common.cuh"

namespace mvin {

template <int D>
struct TC {
  static constexpr int LPR = D / 4;                                   // float4 lanes per row
  static constexpr int NT = (D >= 64) ? 512 : (D >= 16) ? 256 : 128;  // threads per CTA
  static constexpr int NTY = NT / LPR;                                // thread rows of the register tile
  static constexpr int R = 64;                                        // rows per tile
  static constexpr int TM = R / NTY;                                  // rows per thread
  static constexpr int LD = D + 4;                                    // padded leading dimension of a smem row tile
  static constexpr int NW = NT / 32;                                  // warps per CTA
  static constexpr int G = 32 / LPR;                                  // rows one warp load instruction covers
  static constexpr int TMW = (D / NTY) > 0 ? (D / NTY) : 1;           // dW rows per thread (SIMT path)
  static constexpr int LDW = D + 8;                                   // padded leading dimension of a smem weight
  static constexpr int WSZ = D * LDW;                                 // floats of one smem weight
  // tensor-core path (3xTF32 mma.sync, d >= 32): the 64 x D output tile is split into 4 row blocks of 16 rows and
  // NW/4 column groups, one (row block, column group) per warp
  static constexpr bool TCORE = D >= 32;
  static constexpr int CGRP = NW / 4 > 0 ? NW / 4 : 1;                // column groups
  static constexpr int NTW = D / CGRP / 8 > 0 ? D / CGRP / 8 : 1;     // n8 tiles per warp
  static constexpr int MT = D / 16 > 0 ? D / 16 : 1;                  // m16 tiles of a D x D weight gradient
  static constexpr int TPW = (MT * (D / 8)) / NW > 0 ? (MT * (D / 8)) / NW : 1;   // dW n8 tiles per warp
  static constexpr int DWN = TCORE ? TPW : TMW;                       // per-thread dW accumulator rows of 4 floats
};

constexpr int MAX_K = 64;
constexpr int MAX_LV = 3;

// rows per pair are a runtime K^h: q = n / d through a 64-bit multiply-high with M = floor(2^64 / d) + 1 (exact for
// n, d < 2^32; d = 1 is flagged by M = 0) instead of the ~100-instruction 64-bit software division
MVIN_DEV long fastdiv(long n, unsigned long long M) {
  return M == 0 ? n : (long)__umul64hi((unsigned long long)n, M);
}

// which level does this CTA work on?  CTAs [cta_end[l-1], cta_end[l]) own level l.
struct CtaSlice { int level, local, count; };
MVIN_DEV CtaSlice cta_slice(const int* cta_end, int nlev) {
  int l = 0, begin = 0;
  while (l + 1 < nlev && (int)blockIdx.x >= cta_end[l]) { begin = cta_end[l]; ++l; }
  CtaSlice s;
  s.level = l;
  s.local = (int)blockIdx.x - begin;
  s.count = cta_end[l] - begin;
  return s;
}

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
      const float4 w = ld4(&Ws[(k + kk) * TC<D>::LDW + tx * 4]);
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

// ---- 3xTF32 tensor-core tile GEMMs (mma.sync.m16n8k8): fp32 operands are split x = hi + lo with hi, lo in
// TF32, and  a.b ~= hi_a.hi_b + hi_a.lo_b + lo_a.hi_b  is accumulated in fp32 -- fp32-level accuracy (the dropped
// lo.lo term is ~2^-22 relative), which plain TF32 does not give (parity is 1e-4 on the final scores).
// hi = x truncated to TF32 (one LOP3; the residual x - hi is then exact in fp32), lo = x - hi handed to the tensor
// core as is: the MMA reads the TF32 bits of its operands only, i.e. truncates lo to 10 mantissa bits, leaving a
// ~2^-20 relative error per product.  (cvt.rna.tf32.f32 costs ~8 SASS instructions on sm_100 and made the split,
// not the MMA, the bottleneck.)
MVIN_DEV void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
MVIN_DEV void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
MVIN_DEV void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], const uint32_t (&bh)[2],
                   const uint32_t (&bl)[2]) {
  mma_tf32(c, al, bh);
  mma_tf32(c, ah, bl);
  mma_tf32(c, ah, bh);
}

// acc[i][0..3] (thread-mapped, rows ty*TM+i, cols tx*4..) += (As . Ws)[row][col];  As [64][LD] is CLOBBERED (it
// receives the product), Ws is [D][LDW].  Every thread of the CTA must call it; As must be complete and
// synchronised on entry; the caller synchronises before overwriting As again.
template <int D>
MVIN_DEV void tile_mm(float* __restrict__ As, const float* __restrict__ Ws, int ty, int tx,
                      float (&acc)[TC<D>::TM][4]) {
  using C = TC<D>;
  if constexpr (!C::TCORE) {
    mm_tile<D>(As, Ws, ty, tx, acc);
  } else {
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, g = lane / 4, t = lane % 4;
    const int r0 = (warp % 4) * 16, n0 = (warp / 4) * (C::NTW * 8);
    float c[C::NTW][4];
#pragma unroll
    for (int j = 0; j < C::NTW; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
#pragma unroll 2
    for (int k0 = 0; k0 < D; k0 += 8) {
      uint32_t ah[4], al[4];
      split_tf32(As[(r0 + g) * C::LD + k0 + t], ah[0], al[0]);
      split_tf32(As[(r0 + g + 8) * C::LD + k0 + t], ah[1], al[1]);
      split_tf32(As[(r0 + g) * C::LD + k0 + t + 4], ah[2], al[2]);
      split_tf32(As[(r0 + g + 8) * C::LD + k0 + t + 4], ah[3], al[3]);
#pragma unroll
      for (int j = 0; j < C::NTW; ++j) {
        uint32_t bh[2], bl[2];
        split_tf32(Ws[(k0 + t) * C::LDW + n0 + j * 8 + g], bh[0], bl[0]);
        split_tf32(Ws[(k0 + t + 4) * C::LDW + n0 + j * 8 + g], bh[1], bl[1]);
        mma3(c[j], ah, al, bh, bl);
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < C::NTW; ++j) {
      *reinterpret_cast<float2*>(&As[(r0 + g) * C::LD + n0 + j * 8 + 2 * t]) = make_float2(c[j][0], c[j][1]);
      *reinterpret_cast<float2*>(&As[(r0 + g + 8) * C::LD + n0 + j * 8 + 2 * t]) = make_float2(c[j][2], c[j][3]);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const float4 v = ld4(&As[(ty * C::TM + i) * C::LD + tx * 4]);
      acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
    }
  }
}

// dW += A^T G over the 64 rows of a tile (A = As [64][LD], G = Gs [64][LD]); register-resident accumulator.
//   SIMT path:  dw[i] = dW[ty*TMW+i][tx*4 .. +3]
//   tensor-core path: warp owns m16 tile (warp % MT) and TPW consecutive n8 tiles; dw[j] is the mma C fragment
template <int D>
MVIN_DEV void dw_tile(const float* __restrict__ As, const float* __restrict__ Gs, int ty, int tx,
                      float (&dw)[TC<D>::DWN][4]) {
  using C = TC<D>;
  constexpr int LD = C::LD, R = C::R;
  if constexpr (!C::TCORE) {
    constexpr int TMW = C::TMW;
    if (ty * TMW >= D) return;
#pragma unroll 4
    for (int r = 0; r < R; ++r) {
      const float4 gv = ld4(&Gs[r * LD + tx * 4]);
      float av[TMW];
      if constexpr (TMW % 4 == 0) {
#pragma unroll
        for (int q = 0; q < TMW / 4; ++q) {
          const float4 a4 = ld4(&As[r * LD + ty * TMW + q * 4]);
          av[q * 4 + 0] = a4.x; av[q * 4 + 1] = a4.y; av[q * 4 + 2] = a4.z; av[q * 4 + 3] = a4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < TMW; ++i) av[i] = As[r * LD + ty * TMW + i];
      }
#pragma unroll
      for (int i = 0; i < TMW; ++i) {
        dw[i][0] = fmaf(av[i], gv.x, dw[i][0]);
        dw[i][1] = fmaf(av[i], gv.y, dw[i][1]);
        dw[i][2] = fmaf(av[i], gv.z, dw[i][2]);
        dw[i][3] = fmaf(av[i], gv.w, dw[i][3]);
      }
    }
  } else {
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, g = lane / 4, t = lane % 4;
    const int i0 = (warp % C::MT) * 16, j0 = (warp / C::MT) * (C::TPW * 8);
#pragma unroll 2
    for (int k0 = 0; k0 < R; k0 += 8) {
      uint32_t ah[4], al[4];
      split_tf32(As[(k0 + t) * LD + i0 + g], ah[0], al[0]);
      split_tf32(As[(k0 + t) * LD + i0 + g + 8], ah[1], al[1]);
      split_tf32(As[(k0 + t + 4) * LD + i0 + g], ah[2], al[2]);
      split_tf32(As[(k0 + t + 4) * LD + i0 + g + 8], ah[3], al[3]);
#pragma unroll
      for (int j = 0; j < C::TPW; ++j) {
        uint32_t bh[2], bl[2];
        split_tf32(Gs[(k0 + t) * LD + j0 + j * 8 + g], bh[0], bl[0]);
        split_tf32(Gs[(k0 + t + 4) * LD + j0 + j * 8 + g], bh[1], bl[1]);
        mma3(dw[j], ah, al, bh, bl);
      }
    }
  }
}

// CTA-level flush of the register-resident weight gradient to global memory
template <int D>
MVIN_DEV void dw_flush(float (&dw)[TC<D>::DWN][4], float* __restrict__ dW, int ty, int tx) {
  using C = TC<D>;
  if constexpr (!C::TCORE) {
    constexpr int TMW = C::TMW;
    if (ty * TMW >= D) return;
#pragma unroll
    for (int i = 0; i < TMW; ++i)
      red_add4(dW + (long)(ty * TMW + i) * D + tx * 4, make_float4(dw[i][0], dw[i][1], dw[i][2], dw[i][3]));
  } else {
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, g = lane / 4, t = lane % 4;
    const int i0 = (warp % C::MT) * 16, j0 = (warp / C::MT) * (C::TPW * 8);
#pragma unroll
    for (int j = 0; j < C::TPW; ++j) {
      red_add2(dW + (long)(i0 + g) * D + j0 + j * 8 + 2 * t, dw[j][0], dw[j][1]);
      red_add2(dW + (long)(i0 + g + 8) * D + j0 + j * 8 + 2 * t, dw[j][2], dw[j][3]);
    }
  }
}
// red[D] (shared, zeroed) += per-thread column partials, then one global atomic per column
template <int D>
MVIN_DEV void bias_flush(float4 part, float* __restrict__ red, float* __restrict__ db, int tid, int tx) {
  atomicAdd(&red[tx * 4 + 0], part.x);
  atomicAdd(&red[tx * 4 + 1], part.y);
  atomicAdd(&red[tx * 4 + 2], part.z);
  atomicAdd(&red[tx * 4 + 3], part.w);
  __syncthreads();
  if (tid < D) atomicAdd(db + tid, red[tid]);
  __syncthreads();
}

template <int D>
MVIN_DEV void load_weight(float* __restrict__ Ws, const float* __restrict__ Wg, int tid) {
  for (int i = tid * 4; i < D * D; i += TC<D>::NT * 4) st4(&Ws[(i / D) * TC<D>::LDW + i % D], ldg4(&Wg[i]));
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

// accumulate per-pair sums of the rows of a smem tile into du[B, D]; rows of one pair are contiguous, so each
// thread walks a strip of the tile and flushes one atomic per (pair, column) run.
template <int D>
MVIN_DEV void tile_rows_to_pairs(const float* __restrict__ As, float* __restrict__ du, long row0, long rows,
                                 unsigned long long rpp_magic, int tid) {
  using C = TC<D>;
  constexpr int PARTS = (C::NT / D) < C::R ? (C::NT / D) : C::R;
  constexpr int RPS = C::R / PARTS;                   // rows per strip
  const int col = tid % D, part = tid / D;
  if (part >= PARTS) return;
  long cur = -1;
  float acc = 0.f;
#pragma unroll 4
  for (int r = part * RPS; r < (part + 1) * RPS; ++r) {
    const long row = row0 + r;
    if (row >= rows) break;
    const long b = fastdiv(row, rpp_magic);
    if (b != cur) {
      if (cur >= 0) atomicAdd(du + cur * D + col, acc);
      cur = b;
      acc = 0.f;
    }
    acc += As[r * C::LD + col];
  }
  if (cur >= 0) atomicAdd(du + cur * D + col, acc);
}

// ---------------------------------------------------------------------------------------------------------
// user-oriented transform  T = (E[ent] + u) . W_t[h] + b_t[h]      (model.py:270-283), levels h < L, one launch
// ---------------------------------------------------------------------------------------------------------
struct TransformLevel {
  const int32_t* ent;   // [rows]
  const float* W;       // [D, D]   W_t[h]   (backward: W_t[h] transposed)
  const float* b;       // [D]
  float* T;             // fwd out [rows, D]
  const float* g1;      // bwd in  [rows, D]  dT = g1 (+ g2)
  const float* g2;      // bwd in, optional
  float* dW;            // bwd out [D, D]  (accumulated)
  float* db;            // bwd out [D]
  long rows;
  int rpp;              // rows per pair = K^h
  unsigned long long rpp_magic;
};
struct TransformArgs {
  TransformLevel lv[MAX_LV];
  int nlev;
  int cta_end[MAX_LV];
  ETab E;               // entity table
  const float* u;       // [B, D]  user_o
  GTab dE;              // bwd: entity-table gradient (scatter-add)
  float* du;            // bwd: [B, D] (accumulated)
};

template <int D>
__global__ void __launch_bounds__(TC<D>::NT) transform_fwd_kernel(TransformArgs a) {
  using C = TC<D>;
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;
  float* As = Ws + C::WSZ;
  const int tid = threadIdx.x, tx = tid % C::LPR, ty = tid / C::LPR;
  const CtaSlice cs = cta_slice(a.cta_end, a.nlev);
  const TransformLevel& L = a.lv[cs.level];
  load_weight<D>(Ws, L.W, tid);
  const float4 bias = ldg4(L.b + tx * 4);
  const long ntiles = (L.rows + C::R - 1) / C::R;
  for (long t = cs.local; t < ntiles; t += cs.count) {
    const long row0 = t * C::R;
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const int r = ty * C::TM + i;
      const long row = row0 + r;
      float4 x = f4zero();
      if (row < L.rows) {
        const long e = L.ent[row];
        x = f4add(ldg4(erow(a.E, e, D) + tx * 4), ldg4(a.u + fastdiv(row, L.rpp_magic) * D + tx * 4));
      }
      st4(&As[r * C::LD + tx * 4], x);
    }
    __syncthreads();
    float acc[C::TM][4];
#pragma unroll
    for (int i = 0; i < C::TM; ++i) { acc[i][0] = bias.x; acc[i][1] = bias.y; acc[i][2] = bias.z; acc[i][3] = bias.w; }
    tile_mm<D>(As, Ws, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const long row = row0 + ty * C::TM + i;
      if (row < L.rows) st4(L.T + row * D + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    }
    __syncthreads();
  }
}

// backward: dT = g1 (+ g2);  dW_t[h] += XU^T dT with XU = E[ent] + u recomputed;  db += sum dT;
//           gx = dT . W_t[h]^T ;  dE[ent] += gx ;  du[b] += sum_rows gx
template <int D>
__global__ void __launch_bounds__(TC<D>::NT) transform_bwd_kernel(TransformArgs a) {
  using C = TC<D>;
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;
  float* As = Ws + C::WSZ;                 // dT tile, later gx tile
  float* Xs = As + C::R * C::LD;          // XU tile
  float* red = Xs + C::R * C::LD;         // [D]
  const int tid = threadIdx.x, tx = tid % C::LPR, ty = tid / C::LPR;
  const CtaSlice cs = cta_slice(a.cta_end, a.nlev);
  const TransformLevel& L = a.lv[cs.level];
  load_weight<D>(Ws, L.W, tid);
  if (tid < D) red[tid] = 0.f;
  float dw[C::DWN][4];
#pragma unroll
  for (int i = 0; i < C::DWN; ++i) dw[i][0] = dw[i][1] = dw[i][2] = dw[i][3] = 0.f;
  float4 bpart = f4zero();
  __syncthreads();
  const long ntiles = (L.rows + C::R - 1) / C::R;
  for (long t = cs.local; t < ntiles; t += cs.count) {
    const long row0 = t * C::R;
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const int r = ty * C::TM + i;
      const long row = row0 + r;
      float4 g = f4zero(), x = f4zero();
      if (row < L.rows) {
        g = ld4(L.g1 + row * D + tx * 4);
        if (L.g2) g = f4add(g, ld4(L.g2 + row * D + tx * 4));
        const long e = L.ent[row];
        x = f4add(ldg4(erow(a.E, e, D) + tx * 4), ldg4(a.u + fastdiv(row, L.rpp_magic) * D + tx * 4));
      }
      bpart = f4add(bpart, g);
      st4(&As[r * C::LD + tx * 4], g);
      st4(&Xs[r * C::LD + tx * 4], x);
    }
    __syncthreads();
    dw_tile<D>(Xs, As, ty, tx, dw);
    float acc[C::TM][4];
#pragma unroll
    for (int i = 0; i < C::TM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    tile_mm<D>(As, Ws, ty, tx, acc);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const int r = ty * C::TM + i;
      const long row = row0 + r;
      const float4 gx = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      if (row < L.rows) red_add4(grow_of(a.dE, L.ent[row], D) + tx * 4, gx);
      st4(&As[r * C::LD + tx * 4], gx);
    }
    __syncthreads();
    tile_rows_to_pairs<D>(As, a.du, row0, L.rows, L.rpp_magic, tid);
    __syncthreads();
  }
  dw_flush<D>(dw, L.dW, ty, tx);
  bias_flush<D>(bpart, red, L.db, tid, tx);
}

template <int D>
constexpr size_t transform_fwd_smem() { return sizeof(float) * (TC<D>::WSZ + TC<D>::R * TC<D>::LD); }
template <int D>
constexpr size_t transform_bwd_smem() { return sizeof(float) * (TC<D>::WSZ + 2 * TC<D>::R * TC<D>::LD + D); }

// ---------------------------------------------------------------------------------------------------------
// one aggregator iteration, forward, all levels  (aggregators.py:98-146; model.py:295-306)
//   inner level:  agg = (1/K) sum_k p_k child[row*K+k]          (children are the rows of the next level)
//   leaf level :  S = sum_k p_k E[adj[e][k]];  agg = ((S + u) . W_t[L] + b_t[L]) / K     (hoisted transform)
//   Y = self + agg;   V = relu(Y . W_a + b_a)
// ---------------------------------------------------------------------------------------------------------
struct AggLevel {
  const int32_t* ent;   // [rows] entity id of each node of this level
  const float* child;   // inner: [rows*K, D]
  const float* self;    // [rows, D]
  float* SU;            // leaf: [rows, D]  S + u
  float* Y;             // [rows, D]
  float* V;             // [rows, D]
  long rows;
  int rpp;
  unsigned long long rpp_magic;
  int leaf;
};
struct AggArgs {
  AggLevel lv[MAX_LV];
  int nlev;
  int cta_end[MAX_LV];
  const int32_t* adj;   // packed adjacency [n_entity][2][K]
  const float* s;       // [n_rel] relation scores of this aggregator
  ETab E;               // leaf: entity table
  const float* u;       // leaf: [B, D]
  const float* Wt;      // leaf: W_t[L] [D, D]
  const float* bt;      // leaf: b_t[L]
  const float* Wa;      // [D, D]
  const float* ba;      // [D]
  const float* Se;      // leaf, entity mode: [n_entity, D] per-entity S = sum_k p_k E[n_k] (leaf_entity_fwd_kernel)
  int K, n_rel;
};

template <int D, bool HAS_LEAF>
__global__ void __launch_bounds__(TC<D>::NT) agg_fwd_kernel(AggArgs a) {
  using C = TC<D>;
  extern __shared__ __align__(16) float smem[];
  float* Wa_s = smem;
  float* Wt_s = Wa_s + C::WSZ;
  float* As = Wt_s + (HAS_LEAF ? C::WSZ : 0);
  float* pw = As + C::R * C::LD;                           // [NW][MAX_K]
  int* idw = reinterpret_cast<int*>(pw + C::NW * MAX_K);   // [NW][MAX_K]
  float* s_s = reinterpret_cast<float*>(idw + C::NW * MAX_K);
  const int tid = threadIdx.x, tx = tid % C::LPR, ty = tid / C::LPR;
  const int warp = tid / 32, lane = tid % 32, g = lane / C::LPR, c = lane % C::LPR;
  const CtaSlice cs = cta_slice(a.cta_end, a.nlev);
  const AggLevel& L = a.lv[cs.level];
  const bool leaf = HAS_LEAF && L.leaf;
  load_weight<D>(Wa_s, a.Wa, tid);
  if (leaf) load_weight<D>(Wt_s, a.Wt, tid);
  for (int i = tid; i < a.n_rel; i += C::NT) s_s[i] = a.s[i];
  const float4 ba = ldg4(a.ba + tx * 4);
  float4 bt = f4zero();
  if (leaf) bt = ldg4(a.bt + tx * 4);
  const int K = a.K;
  const float invK = 1.f / (float)K;
  float* pw_w = pw + warp * MAX_K;
  int* idw_w = idw + warp * MAX_K;
  __syncthreads();

  const long ntiles = (L.rows + C::R - 1) / C::R;
  for (long t = cs.local; t < ntiles; t += cs.count) {
    const long row0 = t * C::R;
    const bool ent_mode = leaf && a.Se != nullptr;
    if (ent_mode) {
      // entity mode: S depends on the node's entity only and was computed once per distinct entity
      // (leaf_entity_kernel); thread-mapped row loads, all TM of them in flight at once
#pragma unroll
      for (int i = 0; i < C::TM; ++i) {
        const int r = ty * C::TM + i;
        const long row = row0 + r;
        float4 o = f4zero();
        if (row < L.rows) {
          const long e = L.ent[row];
          o = f4add(ldg4(a.Se + e * D + tx * 4), ldg4(a.u + fastdiv(row, L.rpp_magic) * D + tx * 4));
          st4(L.SU + row * D + tx * 4, o);
        }
        st4(&As[r * C::LD + tx * 4], o);
      }
    }
    // ---- neighbour phase: warp per row ----
    for (int r = warp; r < C::R && !ent_mode; r += C::NW) {
      const long row = row0 + r;
      if (row < L.rows) {
        const long e = L.ent[row];
        const Att at = attend(a.adj + e * 2 * K, K, s_s, lane);
        pw_w[lane] = at.p0;
        pw_w[lane + 32] = at.p1;
        if (leaf) { idw_w[lane] = at.id0; idw_w[lane + 32] = at.id1; }
        __syncwarp();
        float4 acc = f4zero();
        if (leaf) {
#pragma unroll 4
          for (int k = g; k < K; k += C::G) acc = f4fma(pw_w[k], ldg4(erow(a.E, idw_w[k], D) + c * 4), acc);
        } else {
          const float* base = L.child + row * K * D + c * 4;
#pragma unroll 4
          for (int k = g; k < K; k += C::G) acc = f4fma(pw_w[k], ldg4(base + (long)k * D), acc);
        }
        acc = cross_group_sum4<C::LPR>(acc);
        if (g == 0) {
          float4 o;
          if (leaf) {
            o = f4add(acc, ldg4(a.u + fastdiv(row, L.rpp_magic) * D + c * 4));
            st4(L.SU + row * D + c * 4, o);
          } else {
            o = f4fma(invK, acc, ld4(L.self + row * D + c * 4));
            st4(L.Y + row * D + c * 4, o);
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
    if (leaf) {
#pragma unroll
      for (int i = 0; i < C::TM; ++i) { acc[i][0] = bt.x; acc[i][1] = bt.y; acc[i][2] = bt.z; acc[i][3] = bt.w; }
      tile_mm<D>(As, Wt_s, ty, tx, acc);
      __syncthreads();
#pragma unroll
      for (int i = 0; i < C::TM; ++i) {
        const int r = ty * C::TM + i;
        const long row = row0 + r;
        float4 y = f4zero();
        if (row < L.rows) {
          y = f4fma(invK, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]), ld4(L.self + row * D + tx * 4));
          st4(L.Y + row * D + tx * 4, y);
        }
        st4(&As[r * C::LD + tx * 4], y);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < C::TM; ++i) { acc[i][0] = ba.x; acc[i][1] = ba.y; acc[i][2] = ba.z; acc[i][3] = ba.w; }
    tile_mm<D>(As, Wa_s, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const long row = row0 + ty * C::TM + i;
      if (row < L.rows)
        st4(L.V + row * D + tx * 4, make_float4(fmaxf(acc[i][0], 0.f), fmaxf(acc[i][1], 0.f), fmaxf(acc[i][2], 0.f),
                                                fmaxf(acc[i][3], 0.f)));
    }
    __syncthreads();
  }
}

template <int D, bool HAS_LEAF>
constexpr size_t agg_fwd_smem(int n_rel) {
  return sizeof(float) * ((HAS_LEAF ? 2 : 1) * TC<D>::WSZ + TC<D>::R * TC<D>::LD + 2 * TC<D>::NW * MAX_K + n_rel);
}

// ---------------------------------------------------------------------------------------------------------
// one aggregator iteration, backward, all levels (Appendix B of SURVEY.md; tests/fused_model.py is the CPU twin)
//   gout = g1 (+ g2) ;  gz = gout * [V > 0] ;  dW_a += Y^T gz ;  db_a += sum gz
//   gs = gz . W_a^T ;  dself = gs ;  grow = gs / K
//   inner: dchild[row*K+k] = p_k * grow ; dp_k = grow . child_k
//   leaf : dW_t[L] += SU^T grow ; db_t[L] += sum grow ;  gsu = grow . W_t[L]^T ;  du[b] += gsu ;
//          dE[n_k] += p_k * gsu ;  dp_k = gsu . E[n_k]
//   dlogit_k = p_k (dp_k - sum_j p_j dp_j) ;  ds[rel_k] += dlogit_k
// The gradient of a node's output arrives from up to two producers (its own aggregator step one iteration later
// and its parent's dchild); they are summed on load (g1 + g2) instead of read-modify-written, so every level of an
// iteration is independent and shares one launch.
// ---------------------------------------------------------------------------------------------------------
struct AggBwdLevel {
  const int32_t* ent;
  const float* child;   // inner
  const float* V;       // [rows, D] forward output (ReLU mask)
  const float* Y;       // [rows, D] forward GEMM input
  const float* SU;      // leaf
  const float* g1;      // [rows, D]
  const float* g2;      // optional
  float* dself;         // [rows, D]
  float* dchild;        // inner: [rows*K, D]
  long rows;
  int rpp;
  unsigned long long rpp_magic;
  int leaf;
};
struct AggBwdArgs {
  AggBwdLevel lv[MAX_LV];
  int nlev;
  int cta_end[MAX_LV];
  const int32_t* adj;
  const float* s;
  ETab E;               // leaf
  const float* WaT;     // W_a transposed
  const float* WtT;     // leaf: W_t[L] transposed
  float* dWa;           // [D, D]
  float* dba;           // [D]
  float* dWt;           // leaf: [D, D]
  float* dbt;           // leaf: [D]
  GTab dE;              // leaf
  float* du;            // leaf: [B, D]
  float* ds;            // [n_rel]
  float* GSe;           // leaf, entity mode: [n_entity, D] per-entity sum of gsu (consumed by leaf_entity_bwd_kernel)
  int K, n_rel;
};

template <int D, bool HAS_LEAF>
__global__ void __launch_bounds__(TC<D>::NT) agg_bwd_kernel(AggBwdArgs a) {
  using C = TC<D>;
  extern __shared__ __align__(16) float smem[];
  float* Wa_s = smem;
  float* Wt_s = Wa_s + C::WSZ;
  float* Gs = Wt_s + (HAS_LEAF ? C::WSZ : 0);
  float* Ys = Gs + C::R * C::LD;
  float* pw = Ys + C::R * C::LD;                       // [NW][MAX_K]
  float* dpw = pw + C::NW * MAX_K;                     // [NW][MAX_K]
  int* idw = reinterpret_cast<int*>(dpw + C::NW * MAX_K);
  float* s_s = reinterpret_cast<float*>(idw + C::NW * MAX_K);
  float* ds_s = s_s + a.n_rel;
  float* red = ds_s + a.n_rel;                         // [2][D]
  const int tid = threadIdx.x, tx = tid % C::LPR, ty = tid / C::LPR;
  const int warp = tid / 32, lane = tid % 32, g = lane / C::LPR, c = lane % C::LPR;
  const CtaSlice cs = cta_slice(a.cta_end, a.nlev);
  const AggBwdLevel& L = a.lv[cs.level];
  const bool leaf = HAS_LEAF && L.leaf;
  load_weight<D>(Wa_s, a.WaT, tid);
  if (leaf) load_weight<D>(Wt_s, a.WtT, tid);
  for (int i = tid; i < a.n_rel; i += C::NT) { s_s[i] = a.s[i]; ds_s[i] = 0.f; }
  for (int i = tid; i < 2 * D; i += C::NT) red[i] = 0.f;
  const int K = a.K;
  const float invK = 1.f / (float)K;
  float* pw_w = pw + warp * MAX_K;
  float* dpw_w = dpw + warp * MAX_K;
  int* idw_w = idw + warp * MAX_K;
  float dwa[C::DWN][4], dwt[HAS_LEAF ? C::DWN : 1][4];
#pragma unroll
  for (int i = 0; i < C::DWN; ++i) dwa[i][0] = dwa[i][1] = dwa[i][2] = dwa[i][3] = 0.f;
#pragma unroll
  for (int i = 0; i < (HAS_LEAF ? C::DWN : 1); ++i) dwt[i][0] = dwt[i][1] = dwt[i][2] = dwt[i][3] = 0.f;
  float4 bpa = f4zero(), bpt = f4zero();
  __syncthreads();

  const long ntiles = (L.rows + C::R - 1) / C::R;
  for (long t = cs.local; t < ntiles; t += cs.count) {
    const long row0 = t * C::R;
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const int r = ty * C::TM + i;
      const long row = row0 + r;
      float4 gz = f4zero(), y = f4zero();
      if (row < L.rows) {
        float4 go = ld4(L.g1 + row * D + tx * 4);
        if (L.g2) go = f4add(go, ld4(L.g2 + row * D + tx * 4));
        const float4 v = ld4(L.V + row * D + tx * 4);
        gz = make_float4(v.x > 0.f ? go.x : 0.f, v.y > 0.f ? go.y : 0.f, v.z > 0.f ? go.z : 0.f,
                         v.w > 0.f ? go.w : 0.f);
        y = ld4(L.Y + row * D + tx * 4);
      }
      bpa = f4add(bpa, gz);
      st4(&Gs[r * C::LD + tx * 4], gz);
      st4(&Ys[r * C::LD + tx * 4], y);
    }
    __syncthreads();
    dw_tile<D>(Ys, Gs, ty, tx, dwa);
    float acc[C::TM][4];
#pragma unroll
    for (int i = 0; i < C::TM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    tile_mm<D>(Gs, Wa_s, ty, tx, acc);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const int r = ty * C::TM + i;
      const long row = row0 + r;
      const float4 gs = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      const float4 grow = f4scale(gs, invK);
      float4 su = f4zero();
      if (row < L.rows) {
        st4(L.dself + row * D + tx * 4, gs);
        if (leaf) su = ld4(L.SU + row * D + tx * 4);
      }
      st4(&Gs[r * C::LD + tx * 4], grow);
      if (leaf) {
        bpt = f4add(bpt, grow);
        st4(&Ys[r * C::LD + tx * 4], su);
      }
    }
    __syncthreads();
    if (leaf) {
      if constexpr (HAS_LEAF) dw_tile<D>(Ys, Gs, ty, tx, dwt);
#pragma unroll
      for (int i = 0; i < C::TM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      tile_mm<D>(Gs, Wt_s, ty, tx, acc);
      __syncthreads();
#pragma unroll
      for (int i = 0; i < C::TM; ++i) {
        const float4 gsu = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        st4(&Gs[(ty * C::TM + i) * C::LD + tx * 4], gsu);
        if (a.GSe != nullptr) {
          // entity mode: the leaf scatter and the softmax gradient are linear in gsu and depend on the entity only
          const long row = row0 + ty * C::TM + i;
          if (row < L.rows) red_add4(a.GSe + (long)L.ent[row] * D + tx * 4, gsu);
        }
      }
      __syncthreads();
      tile_rows_to_pairs<D>(Gs, a.du, row0, L.rows, L.rpp_magic, tid);
    }
    // ---- neighbour phase: warp per row ----
    for (int r = warp; r < C::R && !(leaf && a.GSe != nullptr); r += C::NW) {
      const long row = row0 + r;
      if (row >= L.rows) break;
      const long e = L.ent[row];
      const Att at = attend(a.adj + e * 2 * K, K, s_s, lane);
      pw_w[lane] = at.p0;
      pw_w[lane + 32] = at.p1;
      if (leaf) { idw_w[lane] = at.id0; idw_w[lane + 32] = at.id1; }
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
          if (leaf) {
            const long n = idw_w[k];
            part = f4dot(gr, ldg4(erow(a.E, n, D) + c * 4));
            red_add4(grow_of(a.dE, n, D) + c * 4, f4scale(gr, pk));
          } else {
            const long cr = (row * K + k) * D + c * 4;
            part = f4dot(gr, ld4(L.child + cr));
            st4(L.dchild + cr, f4scale(gr, pk));
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
  dw_flush<D>(dwa, a.dWa, ty, tx);
  bias_flush<D>(bpa, red, a.dba, tid, tx);
  if (leaf) {
    if constexpr (HAS_LEAF) dw_flush<D>(dwt, a.dWt, ty, tx);
    bias_flush<D>(bpt, red + D, a.dbt, tid, tx);
  }
}

template <int D, bool HAS_LEAF>
constexpr size_t agg_bwd_smem(int n_rel) {
  return sizeof(float) *
         ((HAS_LEAF ? 2 : 1) * TC<D>::WSZ + 2 * TC<D>::R * TC<D>::LD + 3 * TC<D>::NW * MAX_K + 2 * n_rel + 2 * D);
}

// ---------------------------------------------------------------------------------------------------------
// entity mode of the leaf level.  S_e = sum_k p_k(e) E[adj[e][k]] depends on the depth-(L-1) node's ENTITY only
// (attention is relation-only), and a batch re-uses the same depth-(L-1) entities many times
// (2-150x at C2..C4), so the K-row gather, its scatter-add and the softmax gradient run once per DISTINCT entity
// marked in `stamp` (set by the expansion kernel) instead of once per (pair, node):
//   fwd:  Se[e] = sum_k p_k E[n_k]
//   bwd:  g = GSe[e] (sum of gsu over the nodes holding e);  dE[n_k] += p_k g;  dp_k = g . E[n_k];
//         dlogit_k = p_k (dp_k - sum_j p_j dp_j);  ds[rel_k] += dlogit_k
// One warp per entity, grid-stride.
// ---------------------------------------------------------------------------------------------------------
struct LeafEntArgs {
  const int32_t* stamp; // [n_entity] != 0: entity occurs at depth L-1 in this batch
  const int32_t* adj;
  const float* s;       // [n_rel] relation scores of aggregator 0
  ETab E;
  float* Se;            // [n_entity, D]
  float* GSe;           // [n_entity, D]
  GTab dE;
  float* ds;            // [n_rel]
  int n_entity, K, n_rel;
};
constexpr int LEAF_NT = 256, LEAF_NW = LEAF_NT / 32;
inline size_t leaf_entity_smem(int n_rel) { return sizeof(float) * (3 * LEAF_NW * MAX_K + 2 * n_rel); }

template <int D, bool BWD>
__global__ void __launch_bounds__(LEAF_NT) leaf_entity_kernel(LeafEntArgs a) {
  constexpr int LPR = D / 4, G = 32 / LPR;
  extern __shared__ __align__(16) float smem[];
  float* pw = smem;                                        // [NW][MAX_K]
  float* dpw = pw + LEAF_NW * MAX_K;                       // [NW][MAX_K]
  int* idw = reinterpret_cast<int*>(dpw + LEAF_NW * MAX_K);
  float* s_s = reinterpret_cast<float*>(idw + LEAF_NW * MAX_K);
  float* ds_s = s_s + a.n_rel;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, g = lane / LPR, c = lane % LPR;
  for (int i = tid; i < a.n_rel; i += LEAF_NT) { s_s[i] = a.s[i]; ds_s[i] = 0.f; }
  __syncthreads();
  float* pw_w = pw + warp * MAX_K;
  float* dpw_w = dpw + warp * MAX_K;
  int* idw_w = idw + warp * MAX_K;
  const int K = a.K;
  for (long e = (long)blockIdx.x * LEAF_NW + warp; e < a.n_entity; e += (long)gridDim.x * LEAF_NW) {
    if (a.stamp[e] == 0) continue;
    const Att at = attend(a.adj + e * 2 * K, K, s_s, lane);
    pw_w[lane] = at.p0;
    pw_w[lane + 32] = at.p1;
    idw_w[lane] = at.id0;
    idw_w[lane + 32] = at.id1;
    __syncwarp();
    if (!BWD) {
      float4 acc = f4zero();
#pragma unroll 4
      for (int k = g; k < K; k += G) acc = f4fma(pw_w[k], ldg4(erow(a.E, idw_w[k], D) + c * 4), acc);
      acc = cross_group_sum4<LPR>(acc);
      if (g == 0) st4(a.Se + e * D + c * 4, acc);
    } else {
      const float4 gr = ld4(a.GSe + e * D + c * 4);
#pragma unroll 4
      for (int k0 = 0; k0 < K; k0 += G) {
        const int k = k0 + g;
        const bool valid = k < K;
        float part = 0.f;
        if (valid) {
          const long n = idw_w[k];
          part = f4dot(gr, ldg4(erow(a.E, n, D) + c * 4));
          red_add4(grow_of(a.dE, n, D) + c * 4, f4scale(gr, pw_w[k]));
        }
        part = group_sum<LPR>(part);
        if (valid && c == 0) dpw_w[k] = part;
      }
      __syncwarp();
      const float dp0 = lane < K ? dpw_w[lane] : 0.f;
      const float dp1 = lane + 32 < K ? dpw_w[lane + 32] : 0.f;
      const float dot = warp_sum(at.p0 * dp0 + at.p1 * dp1);
      if (lane < K) atomicAdd(&ds_s[at.rel0], at.p0 * (dp0 - dot));
      if (lane + 32 < K) atomicAdd(&ds_s[at.rel1], at.p1 * (dp1 - dot));
    }
    __syncwarp();
  }
  if (BWD) {
    __syncthreads();
    for (int i = tid; i < a.n_rel; i += LEAF_NT)
      if (ds_s[i] != 0.f) atomicAdd(a.ds + i, ds_s[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// grouped weight gradient for the B-row maps (mix layer, user MLP):
//   dW_g[i][j] += sum_r A_g[r][i] G[r][j];  db[j] += sum_r G[r][j] (group 0 only)    blockIdx.y = group g
// ---------------------------------------------------------------------------------------------------------
constexpr int MAX_DW_GROUPS = 8;
struct DwArgs {
  const float* A[MAX_DW_GROUPS];
  long lda[MAX_DW_GROUPS];
  float* dW[MAX_DW_GROUPS];
  const float* G;       // [rows, D]
  float* db;            // [D] or nullptr
  long rows;
};

template <int D>
__global__ void __launch_bounds__(TC<D>::NT) dw_kernel(DwArgs a) {
  using C = TC<D>;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Gs = As + C::R * C::LD;
  float* red = Gs + C::R * C::LD;
  const int tid = threadIdx.x, tx = tid % C::LPR, ty = tid / C::LPR;
  const int grp = blockIdx.y;
  const float* A = a.A[grp];
  const long lda = a.lda[grp];
  const bool do_bias = (grp == 0 && a.db != nullptr);
  if (tid < D) red[tid] = 0.f;
  float dw[C::DWN][4];
#pragma unroll
  for (int i = 0; i < C::DWN; ++i) dw[i][0] = dw[i][1] = dw[i][2] = dw[i][3] = 0.f;
  float4 bpart = f4zero();
  __syncthreads();
  const long ntiles = (a.rows + C::R - 1) / C::R;
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long row0 = t * C::R;
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const int r = ty * C::TM + i;
      const long row = row0 + r;
      const bool ok = row < a.rows;
      const float4 gv = ok ? ld4(a.G + row * D + tx * 4) : f4zero();
      bpart = f4add(bpart, gv);
      st4(&As[r * C::LD + tx * 4], ok ? ld4(A + row * lda + tx * 4) : f4zero());
      st4(&Gs[r * C::LD + tx * 4], gv);
    }
    __syncthreads();
    dw_tile<D>(As, Gs, ty, tx, dw);
    __syncthreads();
  }
  dw_flush<D>(dw, a.dW[grp], ty, tx);
  if (do_bias) bias_flush<D>(bpart, red, a.db, tid, tx);
}

template <int D>
constexpr size_t dw_smem() { return sizeof(float) * (2 * TC<D>::R * TC<D>::LD + D); }

}  // namespace mvin
