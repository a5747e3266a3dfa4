// gemm_tc.cuh -- tcgen05 (3xTF32) versions of the three relation-batched contractions of the user side, d in {32, 64}:
//   Q[b, r, :]   = RK[r]^T v_b                              (model.py:211-220 refactored: v^T R_m h_m = Q[b, r_m] . h_m)
//   dv[b, :]     = sum_r RK[r] dQ[b, r, :]  -> dE[item_b]   (backward of the same, item side)
//   dRK[r]      += sum_b v_b dQ[b, r, :]^T                  (backward, relation-matrix side)
// With a large relation-KGE table (n_rel d beyond the shared memory of the fused user kernel) these are
// [B, d] x [d, n_rel d] products: at C4 2.6 GMAC each, 18 % of the step on the fp32 SIMT kernel of gemm.cuh.  Here one
// CTA owns a 128-row tile of the batch and walks the relations: operands are split x = hi + lo (hi = x truncated to TF32)
// and staged by the threads as K-major no-swizzle core-matrix tiles (umma.cuh), one elected thread issues the three
// tcgen05.mma products per k-step into a tensor-memory accumulator, completion arrives on an mbarrier
// (tcgen05.commit), the epilogue reads the accumulator back with tcgen05.ld.  Staging of the next relation overlaps the
// tensor-core work of the current one (two weight buffers; two accumulators where every relation has its own output).
#pragma once
#include "umma.cuh"

namespace mvin {

constexpr int RG_NT = 256;

struct RelGemmArgs {
  const float* V;        // [B, D]           v = E[item]
  const float* RK;       // [n_rel, D, D]
  float* Q;              // q:   out [B, n_rel, D]
  const float* dQ;       // dv / drk: in [B, n_rel, D]
  float* dE;             // dv:  entity-gradient table rows (+=) -- or a dense [B, D] buffer when rows == nullptr
  const int32_t* rows;   // dv:  [B] item ids
  float* dRK;            // drk: [n_rel, D, D] (+=)
  const unsigned char* RKs;   // q / dv: relation matrices pre-split and pre-staged by rel_stage_kernel (below)
  int B, n_rel;
};

// The relation matrices as the tensor core reads them, built ONCE per step (parameters only): per relation four operand
// images of OpLayout<D>::bytes(D) bytes -- {transposed, plain} x {hi, lo} -- so that the GEMM kernels fetch a relation's
// operand with one bulk asynchronous copy (cp.async.bulk -> mbarrier complete_tx) instead of re-splitting and re-staging
// it with all threads for every batch tile.
template <int D>
MVIN_HD constexpr size_t rel_stage_bytes() { return 4 * (size_t)umma::OpLayout<D>::bytes(D); }
template <int D>
__global__ void __launch_bounds__(RG_NT) rel_stage_kernel(const float* __restrict__ RK, unsigned char* __restrict__ out) {
  pdl_enter();
  using L = umma::OpLayout<D>;
  const int r = blockIdx.x;
  unsigned char* base = out + (size_t)r * rel_stage_bytes<D>();
  if (blockIdx.y == 0) umma::stage_weight_t<D>(base, base + L::bytes(D), RK + (long)r * D * D, D, threadIdx.x, RG_NT);
  else umma::stage_weight<D>(base + 2 * L::bytes(D), base + 3 * L::bytes(D), RK + (long)r * D * D, D, threadIdx.x, RG_NT);
}


template <int D>
inline size_t relq_tc_smem() { return 2 * umma::OpLayout<D>::bytes(128) + 6 * umma::OpLayout<D>::bytes(D) + 64; }
template <int D>
inline size_t reldv_tc_smem() { return 4 * umma::OpLayout<D>::bytes(128) + 4 * umma::OpLayout<D>::bytes(D) + 64; }
template <int D>
inline size_t reldrk_tc_smem() { return 2 * umma::OpLayout<128>::bytes(128) + 2 * umma::OpLayout<128>::bytes(D) + 64; }

// accumulator rows 32 (warp % 4) + lane, columns [c0, c0 + D / 2) of this warp
template <int D, typename F>
MVIN_DEV void rg_epilogue(uint32_t tmem_acc, int warp, F&& store) {
  const int c0 = (warp / 4) * (D / 2);
#pragma unroll
  for (int cc = 0; cc < D / 2; cc += 16) {
    float v[16];
    umma::tmem_ld16(tmem_acc + ((uint32_t)(32 * (warp % 4)) << 16) + (uint32_t)(c0 + cc), v);
    store(c0 + cc, v);
  }
}

// ---- Q[b, r, :] = v_b . RK[r]  : A = V tile (staged once), B(n = j, k = i) = RK[r][i][j] (transposed staging) ----------
template <int D>
__global__ void __launch_bounds__(RG_NT) relq_tc_kernel(RelGemmArgs a) {
  pdl_enter();
  using L = umma::OpLayout<D>;
  constexpr int LPR = D / 4, TCOLS = 2 * D < 32 ? 32 : 2 * D;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* a_hi = smem_raw;
  unsigned char* a_lo = a_hi + L::bytes(128);
  unsigned char* w_base = a_lo + L::bytes(128);            // [3][hi, lo]: ring of bulk-copied relation operands
  uint64_t* bar = reinterpret_cast<uint64_t*>(w_base + 6 * L::bytes(D));   // [2] MMA done, [3] operand landed
  uint64_t* wbar = bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 3);
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  if (warp == 0) umma::tmem_alloc(tmem_slot, TCOLS);
  if (tid == 32) {
    umma::mbar_init(bar, 1);
    umma::mbar_init(bar + 1, 1);
    for (int i = 0; i < 3; ++i) umma::mbar_init(wbar + i, 1);
    umma::fence_barrier_init();
  }
  const long row0 = (long)blockIdx.x * 128;
  const int n_it = ((int)a.n_rel - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;
  // operand of iteration j (relation blockIdx.y + j gridDim.y): {transposed hi, lo} are contiguous in the staged image
  auto fetch = [&](int j) {
    const int r = blockIdx.y + j * gridDim.y;
    bulk::mbar_expect(wbar + j % 3, 2 * L::bytes(D));
    bulk::copy_g2s(w_base + (j % 3) * 2 * L::bytes(D), a.RKs + (size_t)r * rel_stage_bytes<D>(), 2 * L::bytes(D), wbar + j % 3);
  };
  {
    constexpr int NI = 128 * LPR / RG_NT;
    float4 x[NI];
#pragma unroll
    for (int q = 0; q < NI; ++q) {
      const int i = tid + q * RG_NT, r = i / LPR, k4 = i % LPR;
      x[q] = f4zero();
      if (row0 + r < a.B) x[q] = ldg4(a.V + (row0 + r) * D + k4 * 4);
    }
#pragma unroll
    for (int q = 0; q < NI; ++q) {
      const int i = tid + q * RG_NT;
      umma::store_split<D>(a_hi, a_lo, i / LPR, i % LPR, x[q]);
    }
  }
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  if (tid == 0) {
    if (n_it > 0) fetch(0);
    if (n_it > 1) fetch(1);
  }
  const uint32_t tmem = *tmem_slot;
  const long my_row = row0 + 32 * (warp % 4) + lane;
  auto epilogue = [&](int e_it) {
    const int buf = e_it & 1;
    const int r = blockIdx.y + e_it * gridDim.y;
    umma::mbar_wait(bar + buf, (uint32_t)((e_it >> 1) & 1));
    umma::fence_after_sync();
    float* out = a.Q + (my_row * a.n_rel + r) * D;
    rg_epilogue<D>(tmem + buf * D, warp, [&](int col, const float (&v)[16]) {
      if (my_row < a.B) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) st4(out + col + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
      }
    });
  };
  for (int it = 0; it < n_it; ++it) {
    const int buf = it & 1;
    if (it > 0) {                                          // accumulator `buf` was read by the epilogue of iteration it - 2
      umma::fence_before_sync();
      __syncthreads();
      umma::fence_after_sync();
    }
    if (tid == 0) {
      bulk::mbar_wait(wbar + it % 3, (uint32_t)((it / 3) & 1));
      unsigned char* w_hi = w_base + (it % 3) * 2 * L::bytes(D);
      umma::issue_3xtf32<D>(tmem + buf * D, umma::smem_u32(a_hi), umma::smem_u32(a_lo), umma::smem_u32(w_hi),
                            umma::smem_u32(w_hi + L::bytes(D)), D, true);
      umma::commit(bar + buf);
    }
    if (it > 0) epilogue(it - 1);                          // overlaps the tensor-core work just issued
    // the operand slot of iteration it - 1 is free once its MMAs are complete (awaited by the epilogue just above)
    if (tid == 0 && it + 2 < n_it && it > 0) fetch(it + 2);
    if (tid == 0 && it == 0 && n_it > 2) { /* slot 2 has never been used */ fetch(2); }
  }
  if (n_it > 0) epilogue(n_it - 1);
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, TCOLS);
}

// ---- dv[b, :] = sum_r dQ[b, r, :] . RK[r]^T : A = dQ[., r, .] tile, B(n = i, k = j) = RK[r][i][j]; one accumulator -----
template <int D>
__global__ void __launch_bounds__(RG_NT) reldv_tc_kernel(RelGemmArgs a) {
  pdl_enter();
  using L = umma::OpLayout<D>;
  constexpr int LPR = D / 4, TCOLS = D < 32 ? 32 : D;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* a_base = smem_raw;                        // [2][hi, lo]
  unsigned char* w_base = a_base + 4 * L::bytes(128);      // [2][hi, lo]
  uint64_t* bar = reinterpret_cast<uint64_t*>(w_base + 4 * L::bytes(D));   // [2] MMA done, [2] operand landed
  uint64_t* wbar = bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 2);
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  if (warp == 0) umma::tmem_alloc(tmem_slot, TCOLS);
  if (tid == 32) {
    for (int i = 0; i < 4; ++i) umma::mbar_init(bar + i, 1);
    umma::fence_barrier_init();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const long row0 = (long)blockIdx.x * 128;
  int it = 0;
  for (int r = blockIdx.y; r < a.n_rel; r += gridDim.y, ++it) {
    const int buf = it & 1;
    unsigned char* a_hi = a_base + buf * 2 * L::bytes(128);
    unsigned char* a_lo = a_hi + L::bytes(128);
    unsigned char* w_hi = w_base + buf * 2 * L::bytes(D);
    unsigned char* w_lo = w_hi + L::bytes(D);
    if (it >= 2) umma::mbar_wait(bar + buf, (uint32_t)(((it - 2) >> 1) & 1));   // the MMAs that read these buffers are done
    if (tid == 0) {
      // the relation's operand {plain hi, lo}: one bulk copy from the pre-staged image, in flight while the threads split
      // and stage the dQ tile
      bulk::mbar_expect(wbar + buf, 2 * L::bytes(D));
      bulk::copy_g2s(w_hi, a.RKs + (size_t)r * rel_stage_bytes<D>() + 2 * L::bytes(D), 2 * L::bytes(D), wbar + buf);
    }
    {
      // all of the thread's loads first, then the splits and shared-memory stores: a load -> store loop would pay one
      // L2 round trip per trip (measured: 5.6 us per relation before, most of it this loop)
      constexpr int NI = 128 * LPR / RG_NT;
      float4 x[NI];
#pragma unroll
      for (int q = 0; q < NI; ++q) {
        const int i = tid + q * RG_NT, rr = i / LPR, k4 = i % LPR;
        x[q] = f4zero();
        if (row0 + rr < a.B) x[q] = ld4(a.dQ + ((row0 + rr) * a.n_rel + r) * D + k4 * 4);
      }
#pragma unroll
      for (int q = 0; q < NI; ++q) {
        const int i = tid + q * RG_NT;
        umma::store_split<D>(a_hi, a_lo, i / LPR, i % LPR, x[q]);
      }
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (tid == 0) {
      bulk::mbar_wait(wbar + buf, (uint32_t)((it >> 1) & 1));
      umma::issue_3xtf32<D>(tmem, umma::smem_u32(a_hi), umma::smem_u32(a_lo), umma::smem_u32(w_hi), umma::smem_u32(w_lo), D, it == 0);
      umma::commit(bar + buf);
    }
  }
  if (it > 0) {
    const int last = it - 1;
    umma::mbar_wait(bar + (last & 1), (uint32_t)((last >> 1) & 1));   // MMAs complete in issue order
    umma::fence_after_sync();
    const long my_row = row0 + 32 * (warp % 4) + lane;
    float* out = nullptr;
    if (my_row < a.B) out = a.rows ? a.dE + (long)a.rows[my_row] * D : a.dE + my_row * D;
    rg_epilogue<D>(tmem, warp, [&](int col, const float (&v)[16]) {
      if (out) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) red_add4(out + col + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
      }
    });
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, TCOLS);
}

// ---- dRK[r][i][j] += sum_b v[b][i] dQ[b][r][j] over one 128-row tile of the batch: A(m = i, k = b) = v^T (staged once,
// rows D .. 127 zero), B(n = j, k = b) = dQ[., r, .]^T; the partial of every (tile, relation) is reduced into dRK ---------
template <int RWS>
MVIN_DEV void stage_transposed_128(unsigned char* hi, unsigned char* lo, const float* __restrict__ src, long ld, long row0,
                                   long nrows, int tid) {
  // operand element (n, k) = src[(row0 + k) * ld + n], n < RWS, k < 128 (zero beyond nrows)
  // all loads of the thread first (4 NI independent scalar loads), then the splits and stores
  constexpr int NI = RWS * 32 / RG_NT;
  float x[NI][4];
#pragma unroll
  for (int q = 0; q < NI; ++q) {
    const int i = tid + q * RG_NT, n = i % RWS, k4 = i / RWS;   // consecutive threads -> consecutive n: coalesced row segments
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long b = row0 + k4 * 4 + j;
      x[q][j] = b < nrows ? __ldg(src + b * ld + n) : 0.f;
    }
  }
#pragma unroll
  for (int q = 0; q < NI; ++q) {
    const int i = tid + q * RG_NT;
    umma::store_split<128>(hi, lo, i % RWS, i / RWS, make_float4(x[q][0], x[q][1], x[q][2], x[q][3]));
  }
}

template <int D>
__global__ void __launch_bounds__(RG_NT) reldrk_tc_kernel(RelGemmArgs a) {
  pdl_enter();
  using L = umma::OpLayout<128>;
  constexpr int TCOLS = D < 32 ? 32 : D;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* a_hi = smem_raw;                          // v^T   [128 (m = i, zero beyond D)][128 (k = b)]
  unsigned char* a_lo = a_hi + L::bytes(128);
  unsigned char* b_hi = a_lo + L::bytes(128);              // dQ^T  [D (n = j)][128 (k = b)]
  unsigned char* b_lo = b_hi + L::bytes(D);
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_lo + L::bytes(D));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  if (warp == 0) umma::tmem_alloc(tmem_slot, TCOLS);
  if (tid == 32) {
    umma::mbar_init(bar, 1);
    umma::fence_barrier_init();
  }
  const long row0 = (long)blockIdx.x * 128;
  for (int i = tid; i < 2 * L::bytes(128) / 16; i += RG_NT) reinterpret_cast<float4*>(a_hi)[i] = f4zero();
  __syncthreads();
  stage_transposed_128<D>(a_hi, a_lo, a.V, D, row0, a.B, tid);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const int i_row = 32 * (warp % 4) + lane;                // accumulator row = i
  int it = 0;
  // every batch tile reduces into the same dRK[r]: the tiles walk the relations in rotated order, so that at any moment
  // the CTAs' reductions are spread over all relation matrices instead of piling onto one
  const int per = (a.n_rel + gridDim.y - 1) / gridDim.y;
  for (int q = 0; q < per; ++q) {
    const int r = ((q + (int)blockIdx.x) % per) * gridDim.y + blockIdx.y;
    if (r >= a.n_rel) continue;                            // CTA-uniform
    stage_transposed_128<D>(b_hi, b_lo, a.dQ + (long)r * D, (long)a.n_rel * D, row0, a.B, tid);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (tid == 0) {
      umma::issue_3xtf32<128>(tmem, umma::smem_u32(a_hi), umma::smem_u32(a_lo), umma::smem_u32(b_hi), umma::smem_u32(b_lo), D, true);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, (uint32_t)(it & 1));
    umma::fence_after_sync();
    float* out = a.dRK + ((long)r * D + i_row) * D;
    rg_epilogue<D>(tmem, warp, [&](int col, const float (&v)[16]) {
      if (i_row < D) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) red_add4(out + col + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
      }
    });
    umma::fence_before_sync();                             // the next relation overwrites the accumulator and the dQ^T tile
    __syncthreads();
    umma::fence_after_sync();
    ++it;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, TCOLS);
}

}  // namespace mvin
