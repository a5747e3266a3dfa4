// umma_bf.cuh -- tcgen05 building blocks for the BACKWARD d x d maps: split-bf16 operands.
//
// The backward of every dense map of the path (aggregators.py:108-116, model.py:277-281 under TF autodiff) needs the same
// activation tile in two products that contract over DIFFERENT axes:
//     dX[r][k]  = sum_n G[r][n] W[k][n]        contracts over features: G is the K-major operand  (K = features)
//     dW[i][j] += sum_r A[r][i] G[r][j]        contracts over rows:     A and G are MN-major operands (K = rows)
// For kind::tf32 the tensor core accepts an MN-major shared-memory operand only in the 128B_BASE32B swizzle, which no
// K-major form shares, so a tf32 tile would have to be staged twice (and the weight gradients stayed on mma.sync in
// round 1 for that reason).  For 16-bit types the un-swizzled "interleave" form -- core matrices of 8 x 16 bytes -- is
// valid for BOTH majors, and the K-major and the MN-major view of one staged tile are the same bytes with the two
// strides swapped.  So the backward stages each fp32 tile ONCE, as NP bf16 planes
//     x = b0 + b1 (+ b2),   b0 = bf16_rn(x), b1 = bf16_rn(x - b0), ...        (the residuals are exact in fp32)
// and evaluates a.g as the sum of the plane products with i + j < NP in the fp32 accumulator: NP = 2 keeps 16 mantissa
// bits per operand (measured error 6e-6 of the largest entry for K = 64 and for K = 2e5, against 1e-4 allowed on the
// gradients), NP = 3 all 24.  bf16 has the fp32 exponent range, so no per-tile scaling is needed.
//
// Layout of one plane of a [rows][KD] tile: element (r, f) at byte
//     (r / 8) * SBO + (f / 8) * LBO + (r % 8) * 16 + (f % 8) * 2,        SBO = (KD / 8) * LBO.
//   K-major view  (M or N = rows, K = features): descriptor(LBO field = LBO, SBO field = SBO), 16 features per MMA
//   MN-major view (M or N = features, K = rows): descriptor(LBO field = SBO, SBO field = LBO), 16 rows per MMA
// LBO = 144 for activation tiles (the 8-byte stores of the 16 lanes of a row then spread over the banks), 128 for
// weights (staged once per CTA).
#pragma once
#include <cuda_bf16.h>

#include "umma.cuh"

namespace mvin {
namespace umma {

template <int KD, int LBO_>
struct BfLayout {
  static constexpr int LBO = LBO_;
  static constexpr int SBO = (KD / 8) * LBO_;
  MVIN_HD static constexpr int bytes(int rows) { return rows / 8 * SBO; }
  // byte offset of features 4 tx .. 4 tx + 3 of row r (8 bytes)
  MVIN_DEV static int off4(int r, int tx) { return (r >> 3) * SBO + (tx >> 1) * LBO + (r & 7) * 16 + (tx & 1) * 8; }
};

// instruction descriptor of kind::f16 with bf16 operands and fp32 accumulation (bit 4: D = f32, bits 7-9 / 10-12:
// A / B = bf16, bit 15 / 16: A / B MN-major, bits 17-22: N / 8, bits 24-28: M / 16)
MVIN_HD constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

MVIN_DEV void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}

// 4 floats -> NP planes of 4 bf16 (8 bytes each) at `off` of consecutive planes `plane_bytes` apart
template <int NP>
MVIN_DEV void store_planes4(unsigned char* tile, int plane_bytes, int off, float4 x) {
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const __nv_bfloat162 lo = __float22bfloat162_rn(make_float2(x.x, x.y));
    const __nv_bfloat162 hi = __float22bfloat162_rn(make_float2(x.z, x.w));
    uint2 v;
    v.x = *reinterpret_cast<const uint32_t*>(&lo);
    v.y = *reinterpret_cast<const uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(tile + p * plane_bytes + off) = v;
    if (p + 1 < NP) {
      const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
      x = make_float4(x.x - a.x, x.y - a.y, x.z - b.x, x.w - b.y);
    }
  }
}

// stage a row-major [ROWS][KD] global matrix as NP planes (weights: LBO = 128)
template <int KD, int NP, int LBO_>
MVIN_DEV void stage_weight_bf(unsigned char* tile, const float* __restrict__ W, int rows, int tid, int nt) {
  using L = BfLayout<KD, LBO_>;
  for (int i = tid; i < rows * (KD / 4); i += nt) {
    const int r = i / (KD / 4), tx = i % (KD / 4);
    store_planes4<NP>(tile, L::bytes(rows), L::off4(r, tx), ldg4(W + (long)r * KD + tx * 4));
  }
}

// dX: D[128 x N] (+)= A[128 x KD] . B[N x KD]^T, both K-major; one thread issues.  a / b = shared-memory addresses of
// plane 0; planes are aps / bps bytes apart.
template <int KD, int NP, int LBO_A, int LBO_B>
MVIN_DEV void issue_dx_bf(uint32_t tmem_d, uint32_t a, int aps, uint32_t b, int bps, int N, bool first) {
  using LA = BfLayout<KD, LBO_A>;
  using LB = BfLayout<KD, LBO_B>;
  const uint32_t idesc = idesc_bf16(128, N);
  uint32_t acc = first ? 0u : 1u;
#pragma unroll
  for (int k = 0; k < KD / 16; ++k) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
#pragma unroll
      for (int j = 0; i + j < NP; ++j) {
        mma_bf16(tmem_d, make_desc(a + i * aps + k * 2 * LA::LBO, LA::LBO, LA::SBO),
                 make_desc(b + j * bps + k * 2 * LB::LBO, LB::LBO, LB::SBO), idesc, acc);
        acc = 1u;
      }
    }
  }
}

// dW: D[M x N] (+)= A[rows x M]^T . G[rows x N] with M, N = feature counts (M padded to 64) and K = `rows` tile rows;
// both operands are the MN-major views of K-major staged tiles of KD_A / KD_G features per row.
template <int KD_A, int KD_G, int NP, int LBO_>
MVIN_DEV void issue_dw_bf(uint32_t tmem_d, uint32_t a, int aps, uint32_t g, int gps, int rows, bool first) {
  using LA = BfLayout<KD_A, LBO_>;
  using LG = BfLayout<KD_G, LBO_>;
  const uint32_t idesc = idesc_bf16(KD_A < 64 ? 64 : KD_A, KD_G) | IDESC_A_MN | IDESC_B_MN;
  uint32_t acc = first ? 0u : 1u;
  for (int k = 0; k < rows / 16; ++k) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
#pragma unroll
      for (int j = 0; i + j < NP; ++j) {
        mma_bf16(tmem_d, make_desc(a + i * aps + k * 2 * LA::SBO, LA::SBO, LA::LBO),
                 make_desc(g + j * gps + k * 2 * LG::SBO, LG::SBO, LG::LBO), idesc, acc);
        acc = 1u;
      }
    }
  }
}

// 32 consecutive accumulator columns of this thread's lane
MVIN_DEV void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace umma

// ---------------------------------------------------------------------------------------------------------
// self-test kernel (mvin_test_umma_bf16): ONE CTA walks all 128-row tiles of G, A [M, D]:
//   C[r][k] = sum_n G[r][n] W[k][n]   (dX product, W row-major [D][D] as the K-major B operand)
//   dump[lane][j] = raw tensor-memory lanes of  dW[i][j] = sum_r A[r][i] G[r][j]  accumulated over the tiles
// ---------------------------------------------------------------------------------------------------------
template <int D, int NP>
__global__ void __launch_bounds__(256) umma_bf_test_kernel(const float* __restrict__ A, const float* __restrict__ G,
                                                           const float* __restrict__ W, float* __restrict__ C,
                                                           float* __restrict__ dump, long M) {
  pdl_enter();
  using L = umma::BfLayout<D, 144>;
  using LW = umma::BfLayout<D, 128>;
  constexpr int LPR = D / 4, RP = 256 / LPR, PASSES = 128 / RP;
  constexpr int APL = L::bytes(128), WPL = LW::bytes(D);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* g_t = smem_raw;                         // NP planes
  unsigned char* a_t = g_t + NP * APL;
  unsigned char* w_t = a_t + NP * APL + 1024;            // (+1 KB: the M = 64 view of a D = 32 tile reads past the planes)
  uint64_t* bar = reinterpret_cast<uint64_t*>(w_t + NP * WPL);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, tx = tid % LPR, ty = tid / LPR;
  if (warp == 0) umma::tmem_alloc(tmem_slot, 128);       // columns [0, 64): dX accumulator, [64, 128): dW accumulator
  if (tid == 32) {
    umma::mbar_init(bar, 1);
    umma::fence_barrier_init();
  }
  umma::stage_weight_bf<D, NP, 128>(w_t, W, D, tid, 256);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  uint32_t phase = 0;
  const long ntiles = (M + 127) / 128;
  for (long t = 0; t < ntiles; ++t) {
    const long row0 = t * 128;
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int r = ps * RP + ty;
      float4 x = f4zero(), g = f4zero();
      if (row0 + r < M) {
        x = ldg4(A + (row0 + r) * D + tx * 4);
        g = ldg4(G + (row0 + r) * D + tx * 4);
      }
      umma::store_planes4<NP>(a_t, APL, L::off4(r, tx), x);
      umma::store_planes4<NP>(g_t, APL, L::off4(r, tx), g);
    }
    umma::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_sync();
      umma::issue_dx_bf<D, NP, 144, 128>(tmem, umma::smem_u32(g_t), APL, umma::smem_u32(w_t), WPL, D, true);
      umma::issue_dw_bf<D, D, NP, 144>(tmem + 64, umma::smem_u32(a_t), APL, umma::smem_u32(g_t), APL, 128, t == 0);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
    const int r = 32 * (warp % 4) + lane, c0 = (warp / 4) * (D / 2);
#pragma unroll
    for (int cc = 0; cc < D / 2; cc += 16) {
      float v[16];
      umma::tmem_ld16(tmem + ((uint32_t)(32 * (warp % 4)) << 16) + (uint32_t)(c0 + cc), v);
      if (row0 + r < M) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) st4(C + (row0 + r) * D + c0 + cc + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
      }
    }
    umma::fence_before_sync();
    __syncthreads();                                     // planes are rewritten by the next tile
    umma::fence_after_sync();
  }
  {
    const int r = 32 * (warp % 4) + lane, c0 = (warp / 4) * (D / 2);
#pragma unroll
    for (int cc = 0; cc < D / 2; cc += 16) {
      float v[16];
      umma::tmem_ld16(tmem + ((uint32_t)(32 * (warp % 4)) << 16) + (uint32_t)(64 + c0 + cc), v);
#pragma unroll
      for (int j = 0; j < 16; ++j) dump[(long)r * D + c0 + cc + j] = v[j];
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 128);
}
template <int D, int NP>
inline size_t umma_bf_test_smem() {
  return 2 * NP * umma::BfLayout<D, 144>::bytes(128) + 1024 + NP * umma::BfLayout<D, 128>::bytes(D) + 16;
}

}  // namespace mvin
