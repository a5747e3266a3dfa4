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
// between the levels (transform kernels) or pulls tiles of R consecutive rows from a shared list (aggregator
// kernels) and walks them persistently; the d x d weights live in shared memory for the CTA's lifetime.  Three phases per tile:
//   * "stage" (warp per row, all rows of the warp in flight): read the node's packed adjacency record (K ids + K
//     relation ids, contiguous), K-softmax with shuffles, park (p_k, id_k[, rel_k]) in shared memory.
//   * "neighbour" (thread-mapped: LPR = D/4 lanes x 16 B cover one row, thread (ty, tx) owns row ty*TM+i and
//     columns 4tx..4tx+3): stream the K child / entity rows of the thread's row with every 16-byte load
//     independent of the others (ids and weights come from shared memory), 8 in flight per thread.  The backward
//     needs one dot product per (row, k) across the LPR lanes: W of them are reduced together by a butterfly
//     reduce-scatter (W-1 shuffles for W sums).
//   * "dense" (register tile / tensor cores): the d x d maps and the weight gradients, FP32 FFMA for d < 32 and
//     3xTF32 mma.sync for d >= 32 (see gemm.cuh for why not plain TF32).
#pragma once
#include "common.cuh"

namespace mvin {

constexpr int MAX_K = 64;
constexpr int MAX_LV = 3;

template <int D>
struct TC {
  static constexpr int LPR = D / 4;                                   // float4 lanes per row
  static constexpr int NT = (D >= 128) ? 512 : (D >= 16) ? 256 : 128; // threads per CTA
  static constexpr int NTY = NT / LPR;                                // thread rows of the register tile
  static constexpr int R = (D >= 32) ? 32 : 64;                       // rows per tile
  static constexpr int TM = R / NTY;                                  // rows per thread
  static constexpr int LD = D + 4;                                    // padded leading dimension of a smem row tile
  static constexpr int NW = NT / 32;                                  // warps per CTA
  static constexpr int G = 32 / LPR;                                  // rows one warp load instruction covers
  static constexpr int RPW = R / NW;                                  // rows per warp in the stage phase
  static constexpr int TMW = (D / NTY) > 0 ? (D / NTY) : 1;           // dW rows per thread (SIMT path)
  static constexpr int LDW = D + 8;                                   // padded leading dimension of a smem weight
  static constexpr int WSZ = D * LDW;                                 // floats of one smem weight
  static constexpr int W = LPR < 8 ? LPR : 8;                         // dot products reduced together (backward)
  static constexpr int NKW = MAX_K / W;                               // chunks of W neighbours
  // tensor-core path (3xTF32 mma.sync, d >= 32): the R x D output tile is split into RB row blocks of 16 rows and
  // CGRP column groups, one (row block, column group) per warp
  static constexpr bool TCORE = D >= 32;
  static constexpr int RB = R / 16;                                   // row blocks
  static constexpr int CGRP = NW / RB > 0 ? NW / RB : 1;              // column groups
  static constexpr int NTW = D / CGRP / 8 > 0 ? D / CGRP / 8 : 1;     // n8 tiles per warp
  static constexpr int MT = D / 16 > 0 ? D / 16 : 1;                  // m16 tiles of a D x D weight gradient
  static constexpr int TPW = (MT * (D / 8)) / NW > 0 ? (MT * (D / 8)) / NW : 1;   // dW n8 tiles per warp
  static constexpr int DWN = TCORE ? TPW : TMW;                       // per-thread dW accumulator rows of 4 floats
  static_assert(TM >= 1 && TM * NTY == R && RPW >= 1 && RPW * NW == R, "bad tile");
};

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

// Dynamic tile scheduler of the aggregator kernels.  The levels of one launch have very different costs per tile
// (an inner level streams K child rows per row, the per-entity leaf mode streams none), so tiles are not assigned
// statically: the launch's tiles form one list (levels in the order given, heaviest first) and every CTA pulls the
// next index from a global counter.  The counter pair lives in the handle, starts at {0, 0} and is reset by the
// last CTA to leave, so no memset is needed between launches.
struct TileList {
  long tile_end[MAX_LV];   // cumulative tile counts of the levels
  int nlev;
  int* ctr;                // [2] {next tile, CTAs finished}
};
MVIN_DEV int sched_next(const TileList& tl) { return atomicAdd(tl.ctr, 1); }
MVIN_DEV void sched_exit(const TileList& tl) {
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(tl.ctr + 1, 1) == (int)gridDim.x - 1) {
      tl.ctr[0] = 0;
      tl.ctr[1] = 0;
      __threadfence();
    }
  }
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

// acc[i][0..3] (thread-mapped, rows ty*TM+i, cols tx*4..) += (As . Ws)[row][col];  As [R][LD] is CLOBBERED (it
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
    const int r0 = (warp % C::RB) * 16, n0 = (warp / C::RB) * (C::NTW * 8);
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

// dW += A^T G over the R rows of a tile (A = As [R][LD], G = Gs [R][LD]); register-resident accumulator.
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
// db[D] += column sums of the per-thread partials: warp shuffle over the rows of a warp, one smem slot per warp,
// then one global atomic per column.  `scratch` is >= NW * D floats of shared memory nobody else is using; ends
// with a __syncthreads().
template <int D>
MVIN_DEV void bias_flush(float4 part, float* __restrict__ scratch, float* __restrict__ db, int tid) {
  using C = TC<D>;
  const int warp = tid / 32, lane = tid % 32, g = lane / C::LPR, c = lane % C::LPR;
  part = cross_group_sum4<C::LPR>(part);
  __syncthreads();
  if (g == 0) st4(&scratch[warp * D + c * 4], part);
  __syncthreads();
  if (tid < D) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < C::NW; ++w) s += scratch[w * D + tid];
    atomicAdd(db + tid, s);
  }
  __syncthreads();
}

template <int D>
MVIN_DEV void load_weight(float* __restrict__ Ws, const float* __restrict__ Wg, int tid) {
  for (int i = tid * 4; i < D * D; i += TC<D>::NT * 4) st4(&Ws[(i / D) * TC<D>::LDW + i % D], ldg4(&Wg[i]));
}

// ---- stage phase ------------------------------------------------------------------------------------------
// attention of a node over its K sampled neighbours: p_k = softmax_k(s[rel_k])  (aggregators.py:121-139 with the
// user / self thirds of the logit cancelled).  Each warp handles its RPW rows of the tile, CH
// of them at a time with all their adjacency loads issued before the first softmax; lane l handles k = l and
// k = l + 32.  Output: nb_s[r][k] = (p_k, id_k) and, for the backward, rel_s[r][k].
// UNIFORM (User_orient_rela = 0, aggregators.py:148-152): no attention, every neighbour weighs 1 / K.
template <int D, bool WITH_REL, bool UNIFORM = false>
MVIN_DEV void stage_tile(const int32_t* __restrict__ ent, const int32_t* __restrict__ adj,
                         const float* __restrict__ s_s, long row0, long rows, int K, int KP, int2* __restrict__ nb_s,
                         uint16_t* __restrict__ rel_s, int warp, int lane) {
  using C = TC<D>;
  constexpr int CH = C::RPW < 4 ? C::RPW : 4;
#pragma unroll 1
  for (int j0 = 0; j0 < C::RPW; j0 += CH) {
    int id0[CH], id1[CH], rl0[CH], rl1[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const long row = row0 + warp * C::RPW + j0 + j;
      id0[j] = id1[j] = rl0[j] = rl1[j] = 0;
      if (row < rows) {
        const int32_t* arow = adj + (long)__ldg(ent + row) * 2 * K;
        if (lane < K) { id0[j] = __ldg(arow + lane); rl0[j] = __ldg(arow + K + lane); }
        if (lane + 32 < K) { id1[j] = __ldg(arow + lane + 32); rl1[j] = __ldg(arow + K + lane + 32); }
      }
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int r = warp * C::RPW + j0 + j;
      if (row0 + r >= rows) continue;                       // warp-uniform
      const float l0 = lane < K ? s_s[rl0[j]] : -INFINITY;
      const float l1 = lane + 32 < K ? s_s[rl1[j]] : -INFINITY;
      const float mx = warp_max(fmaxf(l0, l1));
      const float e0 = lane < K ? expf(l0 - mx) : 0.f;
      const float e1 = lane + 32 < K ? expf(l1 - mx) : 0.f;
      const float inv = 1.f / warp_sum(e0 + e1);
      const float p0 = UNIFORM ? 1.f / (float)K : e0 * inv, p1 = UNIFORM ? 1.f / (float)K : e1 * inv;
      if (lane < K) {
        nb_s[r * KP + lane] = make_int2(__float_as_int(p0), id0[j]);
        if (WITH_REL) rel_s[r * KP + lane] = (uint16_t)rl0[j];
      }
      if (lane + 32 < K) {
        nb_s[r * KP + lane + 32] = make_int2(__float_as_int(p1), id1[j]);
        if (WITH_REL) rel_s[r * KP + lane + 32] = (uint16_t)rl1[j];
      }
    }
  }
}
MVIN_DEV int padded_k(int K) { return K | 1; }   // odd row pitch of the staged neighbour records (bank spread)

// ---- per-warp two-slot ring of neighbour rows staged by bulk asynchronous copies (common.cuh, bulk::) ----------------
// One step of the thread-mapped neighbour phase touches, per warp, G = 32 / LPR parent rows x W children = G W <= 32
// table rows of 4 D bytes.  Lane (g, j) issues ONE cp.async.bulk for child j of the warp's g-th parent into the slot of
// the NEXT step (completion: mbarrier transaction count) while the lanes reduce the current slot from shared memory;
// 2 x 512 W bytes per warp.  d <= 64 (at d = 128 the ring of a 16-warp CTA does not fit beside the weights).
template <int D>
struct RowRing {
  using C = TC<D>;
  static constexpr bool ENABLED = D <= 64;
  static constexpr int ROWS = C::G * C::W;              // rows per step
  static constexpr int SLOT = ROWS * D;                 // floats per slot
  static MVIN_HD size_t bytes() { return ENABLED ? (size_t)C::NW * (2 * SLOT * sizeof(float) + 2 * sizeof(uint64_t)) + 16 : 0; }
  float* buf;          // this warp's [2][SLOT]
  uint64_t* bar;       // this warp's [2]
  uint32_t uses0, uses1;   // completed waits per slot (the parity to wait for is uses & 1)
  uint32_t qi, qw;         // steps issued / waited

  // carve the CTA's ring out of `base` (16-byte aligned), initialise the barriers; ends with a __syncthreads()
  MVIN_DEV void init(unsigned char* base, int warp, int lane) {
    buf = reinterpret_cast<float*>(base) + (size_t)warp * 2 * SLOT;
    bar = reinterpret_cast<uint64_t*>(base + (size_t)C::NW * 2 * SLOT * sizeof(float)) + warp * 2;
    uses0 = uses1 = qi = qw = 0;
    if (lane == 0) { bulk::mbar_init(bar, 1); bulk::mbar_init(bar + 1, 1); }
    bulk::fence_mbar_init();
    __syncthreads();
  }
  // rows of the warp's groups: r_first + g * r_stride; children k0 .. k0 + W - 1
  MVIN_DEV void issue(const float* __restrict__ tab, const int2* __restrict__ nb_s, int KP, int r_first, int r_stride, int k0,
                      int K, long row0, long rows, int lane) {
    const int slot = qi & 1;
    const int g = lane / C::W, j = lane % C::W;
    const int r = r_first + g * r_stride;
    const bool valid = lane < ROWS && (row0 + r) < rows && (k0 + j) < K;
    const unsigned m = __ballot_sync(FULL_MASK, valid);   // also: every lane has finished reading this slot
    bulk::fence_proxy_async();
    if (lane == 0) bulk::mbar_expect(bar + slot, (uint32_t)__popc(m) * D * (uint32_t)sizeof(float));
    __syncwarp();
    if (valid)
      bulk::copy_g2s(buf + slot * SLOT + (g * C::W + j) * D, tab + (long)nb_s[r * KP + k0 + j].y * D, D * sizeof(float),
                     bar + slot);
    ++qi;
  }
  MVIN_DEV const float* wait() {
    const int slot = qw & 1;
    if (slot == 0) { bulk::mbar_wait(bar, uses0 & 1); ++uses0; } else { bulk::mbar_wait(bar + 1, uses1 & 1); ++uses1; }
    ++qw;
    return buf + slot * SLOT;
  }
};

// W partial sums per lane -> complete sums across the LPR lanes of a row: butterfly reduce-scatter over the low
// log2(W) lane bits (W - 1 shuffles), then an all-reduce over the remaining bits.  Returns the total of
// v[lane % W]; lanes that agree on lane % W (within a row group) hold the same value.
template <int W, int LPR>
MVIN_DEV float reduce_scatter(float (&v)[W], int lane) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < o; ++j) {
      const float send = up ? v[j] : v[j + o];
      const float keep = up ? v[j + o] : v[j];
      v[j] = keep + __shfl_xor_sync(FULL_MASK, send, o);
    }
  }
  float t = v[0];
#pragma unroll
  for (int o = W; o < LPR; o <<= 1) t += __shfl_xor_sync(FULL_MASK, t, o);
  return t;
}

// du[pair] += sum over the rows of the pair of v (thread-mapped rows ty*TM+i; rows past the end of the level must
// hold zeros).  Rows of one pair are contiguous (rpp = K^h per pair): when a warp's G*TM consecutive rows cannot
// straddle two pairs they are summed with shuffles first, one 16-byte red per warp and column chunk.
template <int D>
MVIN_DEV void pair_accumulate(const float4 (&v)[TC<D>::TM], float* __restrict__ du, long row0, long rows, int rpp,
                              unsigned long long rpp_magic, int ty, int tx, int lane) {
  using C = TC<D>;
  constexpr int RW = C::G * C::TM;
  if (rpp % RW == 0) {
    float4 s = v[0];
#pragma unroll
    for (int i = 1; i < C::TM; ++i) s = f4add(s, v[i]);
    s = cross_group_sum4<C::LPR>(s);
    const long first = row0 + (long)(ty - lane / C::LPR) * C::TM;
    if (lane / C::LPR == 0 && first < rows) red_add4(du + fastdiv(first, rpp_magic) * D + tx * 4, s);
  } else {
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const long row = row0 + ty * C::TM + i;
      if (row < rows) red_add4(du + fastdiv(row, rpp_magic) * D + tx * 4, v[i]);
    }
  }
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
  int stream;           // level buffers >> L2: streaming (evict-first) activation accesses
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
  pdl_enter();
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
      if (row < L.rows) st4a(L.T + row * D + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]), L.stream);
    }
    __syncthreads();
  }
}

// backward: dT = g1 (+ g2);  dW_t[h] += XU^T dT with XU = E[ent] + u recomputed;  db += sum dT;
//           gx = dT . W_t[h]^T ;  dE[ent] += gx ;  du[b] += sum_rows gx
template <int D>
__global__ void __launch_bounds__(TC<D>::NT) transform_bwd_kernel(TransformArgs a) {
  pdl_enter();
  using C = TC<D>;
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;
  float* As = Ws + C::WSZ;                 // dT tile, later gx tile
  float* Xs = As + C::R * C::LD;          // XU tile
  const int tid = threadIdx.x, tx = tid % C::LPR, ty = tid / C::LPR;
  const CtaSlice cs = cta_slice(a.cta_end, a.nlev);
  const TransformLevel& L = a.lv[cs.level];
  load_weight<D>(Ws, L.W, tid);
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
        g = ld4a(L.g1 + row * D + tx * 4, L.stream);
        if (L.g2) g = f4add(g, ld4a(L.g2 + row * D + tx * 4, L.stream));
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
    float4 gx[C::TM];
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const long row = row0 + ty * C::TM + i;
      gx[i] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);   // zero for rows past the end (dT = 0)
      if (row < L.rows) red_add4(grow_of(a.dE, L.ent[row], D) + tx * 4, gx[i]);
    }
    pair_accumulate<D>(gx, a.du, row0, L.rows, L.rpp, L.rpp_magic, ty, tx, tid % 32);
    __syncthreads();
  }
  dw_flush<D>(dw, L.dW, ty, tx);
  bias_flush<D>(bpart, As, L.db, tid);
}

template <int D>
constexpr size_t transform_fwd_smem() { return sizeof(float) * (TC<D>::WSZ + TC<D>::R * TC<D>::LD); }
template <int D>
constexpr size_t transform_bwd_smem() { return sizeof(float) * (TC<D>::WSZ + 2 * TC<D>::R * TC<D>::LD); }

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
  int stream;           // level buffers >> L2: streaming (evict-first) activation accesses
  // table mode (table.cuh): the children are not rows of a buffer but  relu(tab[id_k] + Cp[pair])  with id_k from the
  // node's adjacency record -- iteration 0 of the child level evaluated on the fly from its per-entity table
  int virt;
  const float* tab;     // [n_entity, D]  A_{h+1}
  const float* Cp;      // [B, D]         C_{h+1}
  int preagg;           // Y already holds self + agg (group.cuh evaluated the neighbour phase per entity group)
};
struct AggArgs {
  AggLevel lv[MAX_LV];
  TileList tl;
  const int32_t* adj;   // packed adjacency [n_entity][2][K]
  const float* s;       // [n_rel] relation scores of this aggregator
  ETab E;               // leaf: entity table
  const float* u;       // leaf: [B, D]
  const float* Wt;      // leaf: W_t[L] [D, D]
  const float* bt;      // leaf: b_t[L]
  const float* Wa;      // [D, D]
  const float* ba;      // [D]
  const float* Se;      // leaf, entity mode: [n_entity, D] per-entity S = sum_k p_k E[n_k] (leaf_entity_fwd_kernel)
  int ring;             // the launch carries RowRing shared memory: virt levels stage their table rows with bulk copies
  const float* Xpart;   // leaf, exchange mode (exchange.cuh): [xG][rows, D] per-owner partial sums S^(g), S = sum_g
  int xG;
  int K, n_rel;
};

// shared-memory carve-up shared by the host-side size functions and the kernels (byte offsets, 16-byte aligned)
template <int D>
struct AggSmem {
  using C = TC<D>;
  static MVIN_HD size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
  // forward: Wa | [Wt] | As | nb[R][KP] int2 | s[n_rel]
  static MVIN_HD size_t fwd(bool has_leaf, int K, int n_rel) {
    return sizeof(float) * ((has_leaf ? 2 : 1) * C::WSZ + C::R * C::LD) + align16(sizeof(int2) * C::R * (K | 1)) +
           sizeof(float) * n_rel;
  }
  // backward: WaT | [WtT] | Gs | Ys | nb[R][KP] int2 | rel[R][KP] u16 | s[n_rel] | ds[NWH][n_rel]
  static MVIN_HD int ds_copies(int n_rel) { return n_rel <= 128 ? C::NW : 1; }
  static MVIN_HD size_t bwd(bool has_leaf, int K, int n_rel) {
    return sizeof(float) * ((has_leaf ? 2 : 1) * C::WSZ + 2 * C::R * C::LD) + align16(sizeof(int2) * C::R * (K | 1)) +
           align16(sizeof(uint16_t) * C::R * (K | 1)) + sizeof(float) * n_rel * (1 + ds_copies(n_rel));
  }
};

// UNIFORM (User_orient_rela = 0): agg = mean_k(child_k), the weights 1 / K replace the attention AND the mean's 1 / K
template <int D, bool HAS_LEAF, bool UNIFORM = false>
__global__ void __launch_bounds__(TC<D>::NT) agg_fwd_kernel(AggArgs a) {
  pdl_enter();
  using C = TC<D>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = a.K, KP = padded_k(K);
  float* Wa_s = reinterpret_cast<float*>(smem_raw);
  float* Wt_s = Wa_s + C::WSZ;
  float* As = Wt_s + (HAS_LEAF ? C::WSZ : 0);
  int2* nb_s = reinterpret_cast<int2*>(As + C::R * C::LD);
  float* s_s = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(nb_s) +
                                        AggSmem<D>::align16(sizeof(int2) * C::R * KP));
  const int tid = threadIdx.x, tx = tid % C::LPR, ty = tid / C::LPR;
  const int warp = tid / 32, lane = tid % 32;
  __shared__ int sched[2];
  if (tid == 0) sched[0] = sched_next(a.tl);
  load_weight<D>(Wa_s, a.Wa, tid);
  if (HAS_LEAF) load_weight<D>(Wt_s, a.Wt, tid);
  for (int i = tid; i < a.n_rel; i += C::NT) s_s[i] = a.s[i];
  const float4 ba = ldg4(a.ba + tx * 4);
  float4 bt = f4zero();
  if (HAS_LEAF) bt = ldg4(a.bt + tx * 4);
  const float invK = UNIFORM ? 1.f : 1.f / (float)K;
  RowRing<D> ring;
  const bool use_ring = RowRing<D>::ENABLED && a.ring;
  if (use_ring)
    ring.init(reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(s_s + a.n_rel) + 15) & ~(uintptr_t)15), warp, lane);
  __syncthreads();

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
    const long row0 = (t - (lvl ? a.tl.tile_end[lvl - 1] : 0)) * C::R;
    // ---- stage phase: (p_k, id_k) of every row of the tile ----
    const bool x_mode = leaf && a.Xpart != nullptr;
    if (!ent_mode && !x_mode && !L.preagg) {
      stage_tile<D, false, UNIFORM>(L.ent, a.adj, s_s, row0, L.rows, K, KP, nb_s, nullptr, warp, lane);
      __syncthreads();
    }
    // ---- neighbour phase: thread-mapped ----
    if (L.virt && use_ring) {
      // children = relu(tab[id_k] + Cp[pair]), the table rows staged one step ahead by bulk copies (RowRing)
      const int gq = lane / C::LPR, nc = (K + C::W - 1) / C::W;
      ring.issue(L.tab, nb_s, KP, warp * C::G * C::TM, C::TM, 0, K, row0, L.rows, lane);
#pragma unroll
      for (int i = 0; i < C::TM; ++i) {
        const int r = ty * C::TM + i;
        const long row = row0 + r;
        const bool valid = row < L.rows;
        float4 sv = f4zero(), cv = f4zero();
        if (valid) {
          sv = ld4a(L.self + row * D + tx * 4, L.stream);
          cv = ldg4(L.Cp + fastdiv(row, L.rpp_magic) * D + tx * 4);
        }
        const int2* nb = nb_s + r * KP;
        float4 acc = f4zero();
        for (int c = 0; c < nc; ++c) {
          if (c + 1 < nc) ring.issue(L.tab, nb_s, KP, warp * C::G * C::TM + i, C::TM, (c + 1) * C::W, K, row0, L.rows, lane);
          else if (i + 1 < C::TM) ring.issue(L.tab, nb_s, KP, warp * C::G * C::TM + i + 1, C::TM, 0, K, row0, L.rows, lane);
          const float* slot = ring.wait() + gq * C::W * D + tx * 4;
#pragma unroll
          for (int j = 0; j < C::W; ++j) {
            const int k = c * C::W + j;
            if (valid && k < K) {
              const float4 x = f4add(ld4(slot + j * D), cv);
              acc = f4fma(__int_as_float(nb[k].x), make_float4(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f), fmaxf(x.z, 0.f), fmaxf(x.w, 0.f)), acc);
            }
          }
        }
        float4 o = f4zero();
        if (valid) {
          o = f4fma(invK, acc, sv);
          st4a(L.Y + row * D + tx * 4, o, L.stream);
        }
        st4(&As[r * C::LD + tx * 4], o);
      }
    } else {
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const int r = ty * C::TM + i;
      const long row = row0 + r;
      float4 o = f4zero();
      if (row < L.rows) {
        if (ent_mode) {
          // entity mode: S depends on the node's entity only and was computed once per distinct entity
          // (leaf_entity_kernel)
          const long e = __ldg(L.ent + row);
          o = f4add(ldg4(a.Se + e * D + tx * 4), ldg4(a.u + fastdiv(row, L.rpp_magic) * D + tx * 4));   // not stored:
                                                               // the backward recomputes S + u from the same table
        } else if (leaf) {
          const float4 uv = ldg4(a.u + fastdiv(row, L.rpp_magic) * D + tx * 4);
          const int2* nb = nb_s + r * KP;
          float4 acc = f4zero();
          if (x_mode) {
            // the owners of the leaf rows reduced their share already (exchange.cuh): S = sum over the owners' partials
            for (int g = 0; g < a.xG; ++g) acc = f4add(acc, ld4(a.Xpart + ((long)g * L.rows + row) * D + tx * 4));
          } else {
#pragma unroll 8
            for (int k = 0; k < K; ++k) {
              const int2 v = nb[k];
              acc = f4fma(__int_as_float(v.x), ldg4(erow(a.E, v.y, D) + tx * 4), acc);
            }
          }
          o = f4add(acc, uv);
          st4a(L.SU + row * D + tx * 4, o, L.stream);
        } else if (L.preagg) {
          o = ld4a(L.Y + row * D + tx * 4, L.stream);
        } else if (L.virt) {
          const float4 sv = ld4a(L.self + row * D + tx * 4, L.stream);
          const float4 cv = ldg4(L.Cp + fastdiv(row, L.rpp_magic) * D + tx * 4);
          const int2* nb = nb_s + r * KP;
          const float* base = L.tab + tx * 4;
          float4 acc = f4zero();
#pragma unroll 8
          for (int k = 0; k < K; ++k) {
            const int2 v = nb[k];
            const float4 x = f4add(ldg4(base + (long)v.y * D), cv);
            acc = f4fma(__int_as_float(v.x), make_float4(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f), fmaxf(x.z, 0.f), fmaxf(x.w, 0.f)), acc);
          }
          o = f4fma(invK, acc, sv);
          st4a(L.Y + row * D + tx * 4, o, L.stream);
        } else {
          const float4 sv = ld4a(L.self + row * D + tx * 4, L.stream);
          const int2* nb = nb_s + r * KP;
          const float* base = L.child + row * K * D + tx * 4;
          float4 acc = f4zero();
#pragma unroll 8
          for (int k = 0; k < K; ++k) acc = f4fma(__int_as_float(nb[k].x), ld4a(base + (long)k * D, L.stream), acc);
          o = f4fma(invK, acc, sv);
          st4a(L.Y + row * D + tx * 4, o, L.stream);
        }
      }
      st4(&As[r * C::LD + tx * 4], o);
    }
    }
    __syncthreads();
    // ---- dense phase: register tile / tensor cores ----
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
          y = f4fma(invK, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]), ld4a(L.self + row * D + tx * 4, L.stream));
          st4a(L.Y + row * D + tx * 4, y, L.stream);
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
        st4a(L.V + row * D + tx * 4, make_float4(fmaxf(acc[i][0], 0.f), fmaxf(acc[i][1], 0.f), fmaxf(acc[i][2], 0.f),
                                                fmaxf(acc[i][3], 0.f)), L.stream);
    }
    if (tid == 0) sched[(it + 1) & 1] = nxt;
    __syncthreads();
  }
  sched_exit(a.tl);
}

template <int D, bool HAS_LEAF>
inline size_t agg_fwd_smem(int K, int n_rel) { return AggSmem<D>::fwd(HAS_LEAF, K, n_rel); }

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
  float* gp;            // inner, defer: [rows, D] grow = gs / K of this level's rows (consumed by agg_bwd_leaf_tc_kernel)
  long rows;
  int rpp;
  unsigned long long rpp_magic;
  int leaf;
  int defer;            // inner: the children's share (dchild, dp_k, ds) is evaluated by the tcgen05 leaf kernel of the
                        // child level (level_tcb.cuh); this level only leaves `gp`
  int no_dself;         // defer levels of the fused group mode (group.cuh): dself = K gp is rebuilt by the consumer, not stored
  int stream;           // level buffers >> L2: streaming (evict-first) activation accesses
  // table mode (table.cuh): children = relu(tab[id_k] + Cp[pair]) recomputed; their pre-activation gradient
  // p_k grow * [child > 0] is summed per entity (dtab) and per pair (dCs)
  int virt;
  const float* tab;     // [n_entity, D]
  const float* Cp;      // [B, D]
  float* dtab;          // [n_entity, D] (+=)
  float* dCs;           // [B, D] (+=)
};
struct AggBwdArgs {
  AggBwdLevel lv[MAX_LV];
  TileList tl;
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
  const float* Se;      // leaf, entity mode: [n_entity, D] (S + u is recomputed, not stored)
  const float* u;       // leaf, entity mode: [B, D]
  int ring;             // as in AggArgs
  float* Xgsu;          // leaf, exchange mode (exchange.cuh): [rows, D] gsu = dL/dS of every leaf-level node, left for the
                        // owners of the leaf rows (their scatter-add and the softmax gradient happen on the owner)
  int K, n_rel;
};

// RING: table-gather levels stage their rows through the bulk-copy ring (RowRing; d <= 64).  The ring's shared memory
// limits the SM to two CTAs, so that instantiation is compiled for two (128 registers instead of 85).
// UNIFORM (User_orient_rela = 0): as in agg_fwd_kernel; the relation scores get no gradient.
template <int D, bool HAS_LEAF, bool RING = false, bool UNIFORM = false>
__global__ void __launch_bounds__(TC<D>::NT, (RING ? 2 : D <= 64 ? 3 : 1)) agg_bwd_kernel(AggBwdArgs a) {
  pdl_enter();
  using C = TC<D>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = a.K, KP = padded_k(K);
  float* Wa_s = reinterpret_cast<float*>(smem_raw);
  float* Wt_s = Wa_s + C::WSZ;
  float* Gs = Wt_s + (HAS_LEAF ? C::WSZ : 0);
  float* Ys = Gs + C::R * C::LD;
  int2* nb_s = reinterpret_cast<int2*>(Ys + C::R * C::LD);
  uint16_t* rel_s = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(nb_s) +
                                                AggSmem<D>::align16(sizeof(int2) * C::R * KP));
  float* s_s = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(rel_s) +
                                        AggSmem<D>::align16(sizeof(uint16_t) * C::R * KP));
  float* ds_s = s_s + a.n_rel;                         // [NWH][n_rel]: one private copy per warp when n_rel is small
  const int NWH = AggSmem<D>::ds_copies(a.n_rel);
  const int tid = threadIdx.x, tx = tid % C::LPR, ty = tid / C::LPR;
  const int warp = tid / 32, lane = tid % 32;
  __shared__ int sched[2];
  if (tid == 0) sched[0] = sched_next(a.tl);
  load_weight<D>(Wa_s, a.WaT, tid);
  if (HAS_LEAF) load_weight<D>(Wt_s, a.WtT, tid);
  for (int i = tid; i < a.n_rel; i += C::NT) s_s[i] = a.s[i];
  for (int i = tid; i < NWH * a.n_rel; i += C::NT) ds_s[i] = 0.f;
  const float invK = UNIFORM ? 1.f : 1.f / (float)K;
  float* ds_w = ds_s + (NWH > 1 ? warp : 0) * a.n_rel;
  float dwa[C::DWN][4], dwt[HAS_LEAF ? C::DWN : 1][4];
#pragma unroll
  for (int i = 0; i < C::DWN; ++i) dwa[i][0] = dwa[i][1] = dwa[i][2] = dwa[i][3] = 0.f;
#pragma unroll
  for (int i = 0; i < (HAS_LEAF ? C::DWN : 1); ++i) dwt[i][0] = dwt[i][1] = dwt[i][2] = dwt[i][3] = 0.f;
  float4 bpa = f4zero(), bpt = f4zero();
  bool had_leaf = false;
  RowRing<D> ring;
  constexpr bool use_ring = RING && RowRing<D>::ENABLED;
  if constexpr (use_ring)
    ring.init(reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ds_s + NWH * a.n_rel) + 15) & ~(uintptr_t)15), warp, lane);
  __syncthreads();

  for (int it = 0;; ++it) {
    const long t = sched[it & 1];
    if (t >= a.tl.tile_end[a.tl.nlev - 1]) break;
    int nxt = 0;
    if (tid == 0) nxt = sched_next(a.tl);                    // consumed at the end of this tile
    int lvl = 0;
    while (t >= a.tl.tile_end[lvl]) ++lvl;
    const AggBwdLevel& L = a.lv[lvl];
    const bool leaf = HAS_LEAF && L.leaf;
    const bool ent_mode = leaf && a.GSe != nullptr;
    had_leaf |= leaf;
    const long row0 = (t - (lvl ? a.tl.tile_end[lvl - 1] : 0)) * C::R;
    // ---- stage phase (its loads overlap the tile loads below) ----
    const bool x_mode = leaf && a.Xgsu != nullptr;
    const bool nbr_phase = !ent_mode && !L.defer && !x_mode;
    if (nbr_phase) stage_tile<D, true, UNIFORM>(L.ent, a.adj, s_s, row0, L.rows, K, KP, nb_s, rel_s, warp, lane);
#pragma unroll
    for (int i = 0; i < C::TM; ++i) {
      const int r = ty * C::TM + i;
      const long row = row0 + r;
      float4 gz = f4zero(), y = f4zero();
      if (row < L.rows) {
        float4 go = ld4a(L.g1 + row * D + tx * 4, L.stream);
        if (L.g2) go = f4add(go, ld4a(L.g2 + row * D + tx * 4, L.stream));
        const float4 v = ld4a(L.V + row * D + tx * 4, L.stream);
        gz = make_float4(v.x > 0.f ? go.x : 0.f, v.y > 0.f ? go.y : 0.f, v.z > 0.f ? go.z : 0.f,
                         v.w > 0.f ? go.w : 0.f);
        y = ld4a(L.Y + row * D + tx * 4, L.stream);
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
        if (!L.no_dself) st4a(L.dself + row * D + tx * 4, gs, L.stream);
        if (ent_mode)
          su = f4add(ldg4(a.Se + (long)__ldg(L.ent + row) * D + tx * 4), ldg4(a.u + fastdiv(row, L.rpp_magic) * D + tx * 4));
        else if (leaf) su = ld4a(L.SU + row * D + tx * 4, L.stream);
        if (L.defer) st4(L.gp + row * D + tx * 4, grow);
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
      float4 gsu[C::TM];
#pragma unroll
      for (int i = 0; i < C::TM; ++i) {
        const int r = ty * C::TM + i;
        const long row = row0 + r;
        gsu[i] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);   // zero for rows past the end
        if (ent_mode) {
          // entity mode: the leaf scatter and the softmax gradient are linear in gsu and depend on the entity only
          if (row < L.rows) red_add4(a.GSe + (long)__ldg(L.ent + row) * D + tx * 4, gsu[i]);
        } else if (x_mode) {
          if (row < L.rows) st4(a.Xgsu + row * D + tx * 4, gsu[i]);
        } else {
          st4(&Gs[r * C::LD + tx * 4], gsu[i]);                            // same thread reads it back below
        }
      }
      pair_accumulate<D>(gsu, a.du, row0, L.rows, L.rpp, L.rpp_magic, ty, tx, lane);
    }
    // ---- neighbour phase: thread-mapped.  gr = dL/d(sum_k p_k x_k) of the thread's row; per neighbour k:
    //      dx_k = p_k gr (stored / scattered), dp_k = gr . x_k (W dot products reduced together) ----
    if (nbr_phase) {
      const int kl = lane & (C::W - 1);
      float4 cs[C::TM];                                      // virt: per-row sum of the children's pre-activation gradients
      const bool rv = L.virt && use_ring;                    // table rows staged one step ahead by bulk copies (RowRing)
      const int gq = lane / C::LPR;
      if (rv) ring.issue(L.tab, nb_s, KP, warp * C::G * C::TM, C::TM, 0, K, row0, L.rows, lane);
#pragma unroll
      for (int i = 0; i < C::TM; ++i) {
        const int r = ty * C::TM + i;
        const long row = row0 + r;
        const bool valid = row < L.rows;
        const float4 gr = ld4(&Gs[r * C::LD + tx * 4]);
        cs[i] = f4zero();
        float4 cv = f4zero();
        if (L.virt && valid) cv = ldg4(L.Cp + fastdiv(row, L.rpp_magic) * D + tx * 4);
        const int2* nb = nb_s + r * KP;
        float dp[C::NKW];
#pragma unroll
        for (int c = 0; c < C::NKW; ++c) {
          dp[c] = 0.f;
          if (c * C::W < K) {                                   // CTA-uniform
            float4 x[C::W];
            int2 v[C::W];
            const float* slot = nullptr;
            if (rv) {
              if ((c + 1) * C::W < K) ring.issue(L.tab, nb_s, KP, warp * C::G * C::TM + i, C::TM, (c + 1) * C::W, K, row0, L.rows, lane);
              else if (i + 1 < C::TM) ring.issue(L.tab, nb_s, KP, warp * C::G * C::TM + i + 1, C::TM, 0, K, row0, L.rows, lane);
              slot = ring.wait() + gq * C::W * D + tx * 4;
            }
#pragma unroll
            for (int j = 0; j < C::W; ++j) {
              const int k = c * C::W + j;
              x[j] = f4zero();
              v[j] = make_int2(0, 0);
              if (valid && k < K) {
                v[j] = nb[k];
                if (leaf) x[j] = ldg4(erow(a.E, v[j].y, D) + tx * 4);
                else if (L.virt) {
                  const float4 t = f4add(rv ? ld4(slot + j * D) : ldg4(L.tab + (long)v[j].y * D + tx * 4), cv);
                  x[j] = make_float4(fmaxf(t.x, 0.f), fmaxf(t.y, 0.f), fmaxf(t.z, 0.f), fmaxf(t.w, 0.f));
                } else x[j] = ld4a(L.child + (row * K + k) * D + tx * 4, L.stream);
              }
            }
            float part[C::W];
#pragma unroll
            for (int j = 0; j < C::W; ++j) {
              const int k = c * C::W + j;
              part[j] = f4dot(gr, x[j]);
              if (valid && k < K) {
                float4 dx = f4scale(gr, __int_as_float(v[j].x));
                if (leaf) red_add4(grow_of(a.dE, v[j].y, D) + tx * 4, dx);
                else if (L.virt) {
                  dx = make_float4(x[j].x > 0.f ? dx.x : 0.f, x[j].y > 0.f ? dx.y : 0.f, x[j].z > 0.f ? dx.z : 0.f,
                                   x[j].w > 0.f ? dx.w : 0.f);
                  red_add4(L.dtab + (long)v[j].y * D + tx * 4, dx);
                  cs[i] = f4add(cs[i], dx);
                } else st4a(L.dchild + (row * K + k) * D + tx * 4, dx, L.stream);
              }
            }
            dp[c] = reduce_scatter<C::W, C::LPR>(part, lane);   // dp of k = c*W + kl
          }
        }
        // softmax backward: dlogit_k = p_k (dp_k - sum_j p_j dp_j);  ds[rel_k] += dlogit_k
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < C::NKW; ++c) {
          const int k = c * C::W + kl;
          if (valid && k < K) dot = fmaf(__int_as_float(nb[k].x), dp[c], dot);
        }
#pragma unroll
        for (int o = C::W / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(FULL_MASK, dot, o);
        if (!UNIFORM && valid && tx < C::W) {
#pragma unroll
          for (int c = 0; c < C::NKW; ++c) {
            const int k = c * C::W + tx;
            if (k < K) atomicAdd(&ds_w[rel_s[r * KP + k]], __int_as_float(nb[k].x) * (dp[c] - dot));
          }
        }
      }
      if (L.virt) pair_accumulate<D>(cs, L.dCs, row0, L.rows, L.rpp, L.rpp_magic, ty, tx, lane);   // CTA-uniform branch
    }
    if (tid == 0) sched[(it + 1) & 1] = nxt;
    __syncthreads();
  }
  for (int i = tid; i < a.n_rel; i += C::NT) {
    float s = 0.f;
    for (int w = 0; w < NWH; ++w) s += ds_s[w * a.n_rel + i];
    if (s != 0.f) atomicAdd(a.ds + i, s);
  }
  dw_flush<D>(dwa, a.dWa, ty, tx);
  bias_flush<D>(bpa, Gs, a.dba, tid);
  if constexpr (HAS_LEAF) {
    if (had_leaf) {                                          // CTA-uniform
      dw_flush<D>(dwt, a.dWt, ty, tx);
      bias_flush<D>(bpt, Gs, a.dbt, tid);
    }
  }
  sched_exit(a.tl);
}

template <int D, bool HAS_LEAF>
inline size_t agg_bwd_smem(int K, int n_rel) { return AggSmem<D>::bwd(HAS_LEAF, K, n_rel); }

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
  int chunk;            // entities per warp visit (power of two <= 32): small graphs get one warp per few entities
};
constexpr int LEAF_NT = 256, LEAF_NW = LEAF_NT / 32;
MVIN_HD int leaf_ds_copies(int n_rel) { return n_rel <= 128 ? LEAF_NW : 1; }
inline size_t leaf_entity_smem(int n_rel) {
  return sizeof(float) * (3 * LEAF_NW * MAX_K + (1 + leaf_ds_copies(n_rel)) * n_rel);
}

// adjacency record of one entity spread over a warp: lane l holds k = l and k = l + 32
struct AdjRec { int id0, id1, rel0, rel1; };
MVIN_DEV AdjRec load_adj(const int32_t* __restrict__ adj, long e, int K, int lane) {
  AdjRec r{0, 0, 0, 0};
  const int32_t* arow = adj + e * 2 * K;
  if (lane < K) { r.id0 = __ldg(arow + lane); r.rel0 = __ldg(arow + K + lane); }
  if (lane + 32 < K) { r.id1 = __ldg(arow + lane + 32); r.rel1 = __ldg(arow + K + lane + 32); }
  return r;
}

// One warp per chunk of `chunk` (<= 32) consecutive entities: one coalesced read of their stamps, then the marked ones in
// turn (the adjacency record of the next marked entity is fetched while the current one is processed).
template <int D, bool BWD>
__global__ void __launch_bounds__(LEAF_NT) leaf_entity_kernel(LeafEntArgs a) {
  pdl_enter();
  constexpr int LPR = D / 4, G = 32 / LPR;
  extern __shared__ __align__(16) float smem[];
  float* pw = smem;                                        // [NW][MAX_K]
  float* dpw = pw + LEAF_NW * MAX_K;                       // [NW][MAX_K]
  int* idw = reinterpret_cast<int*>(dpw + LEAF_NW * MAX_K);
  float* s_s = reinterpret_cast<float*>(idw + LEAF_NW * MAX_K);
  float* ds_s = s_s + a.n_rel;                             // [NWH][n_rel]
  const int NWH = leaf_ds_copies(a.n_rel);
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, g = lane / LPR, c = lane % LPR;
  for (int i = tid; i < a.n_rel; i += LEAF_NT) s_s[i] = a.s[i];
  for (int i = tid; i < NWH * a.n_rel; i += LEAF_NT) ds_s[i] = 0.f;
  __syncthreads();
  float* pw_w = pw + warp * MAX_K;
  float* dpw_w = dpw + warp * MAX_K;
  int* idw_w = idw + warp * MAX_K;
  float* ds_w = ds_s + (NWH > 1 ? warp : 0) * a.n_rel;
  const int K = a.K;
  const int chunk = a.chunk;
  for (long c0 = ((long)blockIdx.x * LEAF_NW + warp) * chunk; c0 < a.n_entity; c0 += (long)gridDim.x * LEAF_NW * chunk) {
    const long me = c0 + lane;
    unsigned mask = __ballot_sync(FULL_MASK, lane < chunk && me < a.n_entity && __ldg(a.stamp + me) != 0);
    AdjRec nxt{0, 0, 0, 0};
    if (mask) nxt = load_adj(a.adj, c0 + (__ffs(mask) - 1), K, lane);
    while (mask) {
      const long e = c0 + (__ffs(mask) - 1);
      mask &= mask - 1;
      const AdjRec rec = nxt;
      if (mask) nxt = load_adj(a.adj, c0 + (__ffs(mask) - 1), K, lane);
      const float l0 = lane < K ? s_s[rec.rel0] : -INFINITY;
      const float l1 = lane + 32 < K ? s_s[rec.rel1] : -INFINITY;
      const float mx = warp_max(fmaxf(l0, l1));
      const float e0 = lane < K ? expf(l0 - mx) : 0.f;
      const float e1 = lane + 32 < K ? expf(l1 - mx) : 0.f;
      const float inv = 1.f / warp_sum(e0 + e1);
      const float p0 = e0 * inv, p1 = e1 * inv;
      pw_w[lane] = p0;
      pw_w[lane + 32] = p1;
      idw_w[lane] = rec.id0;
      idw_w[lane + 32] = rec.id1;
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
        const float dot = warp_sum(p0 * dp0 + p1 * dp1);
        if (lane < K) atomicAdd(&ds_w[rec.rel0], p0 * (dp0 - dot));
        if (lane + 32 < K) atomicAdd(&ds_w[rec.rel1], p1 * (dp1 - dot));
      }
      __syncwarp();
    }
  }
  if (BWD) {
    __syncthreads();
    for (int i = tid; i < a.n_rel; i += LEAF_NT) {
      float s = 0.f;
      for (int w = 0; w < NWH; ++w) s += ds_s[w * a.n_rel + i];
      if (s != 0.f) atomicAdd(a.ds + i, s);
    }
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
  const float* Gg[MAX_DW_GROUPS];   // optional per-group G (overrides G)
  float* db;            // [D] or nullptr
  long rows;
};

template <int D>
__global__ void __launch_bounds__(TC<D>::NT) dw_kernel(DwArgs a) {
  pdl_enter();
  using C = TC<D>;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Gs = As + C::R * C::LD;
  const int tid = threadIdx.x, tx = tid % C::LPR, ty = tid / C::LPR;
  const int grp = blockIdx.y;
  const float* A = a.A[grp];
  const float* G = a.Gg[grp] ? a.Gg[grp] : a.G;
  const long lda = a.lda[grp];
  const bool do_bias = (grp == 0 && a.db != nullptr);
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
      const float4 gv = ok ? ld4(G + row * D + tx * 4) : f4zero();
      bpart = f4add(bpart, gv);
      st4(&As[r * C::LD + tx * 4], ok ? ld4(A + row * lda + tx * 4) : f4zero());
      st4(&Gs[r * C::LD + tx * 4], gv);
    }
    __syncthreads();
    dw_tile<D>(As, Gs, ty, tx, dw);
    __syncthreads();
  }
  dw_flush<D>(dw, a.dW[grp], ty, tx);
  if (do_bias) bias_flush<D>(bpart, As, a.db, tid);
}

template <int D>
constexpr size_t dw_smem() { return sizeof(float) * (2 * TC<D>::R * TC<D>::LD); }

}  // namespace mvin
