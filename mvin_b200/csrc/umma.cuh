// umma.cuh -- 5th-generation tensor-core (tcgen05) building blocks for the d x d maps of the MVIN path.
//
// The dense maps inside the aggregator (aggregators.py:108-116, model.py:277-281) are [rows, d] x [d, d] products
// with d in {32, 64}: far too skinny to be compute-bound, but on the legacy mma.sync path their 3xTF32 split
// (fp32-level accuracy, see level.cuh) costs more issue slots than the row streaming around them.  Here one elected
// thread issues tcgen05.mma (kind::tf32, M = 128 rows, N = d, K = 8 per instruction) on operands staged in shared
// memory; the fp32 accumulator lives in tensor memory and is read back with tcgen05.ld for the epilogue.  fp32
// accuracy comes from the same three-product split: x = hi + lo with hi = x truncated to TF32 and lo = x - hi
// (exact in fp32; the tensor core reads the TF32 bits of lo), and  a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi
// accumulated in fp32 -- three MMAs into the same accumulator.
//
// Operand layout ("K-major, no swizzle" canonical form of the sm_100 shared-memory matrix descriptor): core
// matrices of 8 rows x 16 bytes (4 tf32) stored as 128 contiguous bytes; element (r, k) of a tile lives at byte
//     (r / 8) * SBO + (k / 4) * LBO + (r % 8) * 16 + (k % 4) * 4.
// LBO = 144 (not 128) so that the 16-byte stores of the LPR lanes that hold one row's k-chunks fall into different
// bank groups.  The same bytes read with the roles of the two strides swapped are the "MN-major" operand of the
// transposed product (weight gradients), so one staged tile serves both.
#pragma once
#include "common.cuh"

namespace mvin {
namespace umma {

MVIN_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int KD>
struct OpLayout {
  static constexpr int LBO = 144;
  static constexpr int SBO = (KD / 4) * LBO;
  MVIN_HD static constexpr int bytes(int rows) { return rows / 8 * SBO; }
  // byte offset of the 16-byte chunk k4 = k / 4 of row r
  MVIN_DEV static int off(int r, int k4) { return (r >> 3) * SBO + k4 * LBO + (r & 7) * 16; }
};

// shared-memory matrix descriptor: start address, leading / stride byte offsets (16-byte units), version 1 (sm_100),
// layout type 0 (no swizzle)
MVIN_DEV uint64_t make_desc(uint32_t saddr, int lbo, int sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor of kind::tf32 with fp32 accumulation, both operands K-major
MVIN_HD constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
constexpr uint32_t IDESC_A_MN = 1u << 15, IDESC_B_MN = 1u << 16;   // operand is MN-major

MVIN_DEV void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
MVIN_DEV void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
MVIN_DEV void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
MVIN_DEV void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
MVIN_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// generic-proxy shared-memory writes -> visible to the tensor core (async proxy)
MVIN_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
MVIN_DEV void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
MVIN_DEV void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tensor-memory allocation: one full warp; ncols a power of two >= 32; the base address lands in *dst (shared)
MVIN_DEV void tmem_alloc(uint32_t* dst, int ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
MVIN_DEV void tmem_dealloc(uint32_t taddr, int ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// accumulator read-back: warp w reads lanes 32 (w % 4) .. +31 (thread = lane = tile row), 16 consecutive columns
MVIN_DEV void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// x = hi + lo, hi = x truncated to TF32, lo = x - hi (exact)
MVIN_DEV void split4(float4 x, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
  hi.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
  hi.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
  hi.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
  lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
}
// store the 16-byte chunk k4 of row r of an operand tile (hi and lo copies)
template <int KD>
MVIN_DEV void store_split(unsigned char* hi_tile, unsigned char* lo_tile, int r, int k4, float4 x) {
  float4 hi, lo;
  split4(x, hi, lo);
  const int o = OpLayout<KD>::off(r, k4);
  *reinterpret_cast<float4*>(hi_tile + o) = hi;
  *reinterpret_cast<float4*>(lo_tile + o) = lo;
}
// stage a row-major [ROWS][KD] global matrix (a weight: element (n, k) at W[n*KD + k]) as a K-major operand
template <int KD>
MVIN_DEV void stage_weight(unsigned char* hi_tile, unsigned char* lo_tile, const float* __restrict__ W, int rows, int tid,
                           int nt) {
  for (int i = tid; i < rows * (KD / 4); i += nt) {
    const int r = i / (KD / 4), k4 = i % (KD / 4);
    store_split<KD>(hi_tile, lo_tile, r, k4, ldg4(W + (long)r * KD + k4 * 4));
  }
}
// same for the TRANSPOSE of a row-major matrix: operand element (n, k) = W[k*KD + n]
template <int KD>
MVIN_DEV void stage_weight_t(unsigned char* hi_tile, unsigned char* lo_tile, const float* __restrict__ W, int rows, int tid,
                             int nt) {
  for (int i = tid; i < rows * (KD / 4); i += nt) {
    const int r = i % rows, k4 = i / rows;           // consecutive threads -> consecutive r: coalesced reads of W rows
    const float4 x = make_float4(__ldg(W + (long)(k4 * 4 + 0) * KD + r), __ldg(W + (long)(k4 * 4 + 1) * KD + r),
                                 __ldg(W + (long)(k4 * 4 + 2) * KD + r), __ldg(W + (long)(k4 * 4 + 3) * KD + r));
    store_split<KD>(hi_tile, lo_tile, r, k4, x);
  }
}

// D[128 x N] (+)= A[128 x KD] . B[N x KD]^T in 3xTF32; one thread issues.  `first` = overwrite the accumulator.
template <int KD>
MVIN_DEV void issue_3xtf32(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int N, bool first) {
  using L = OpLayout<KD>;
  const uint32_t idesc = idesc_tf32(128, N);
  uint32_t acc = first ? 0u : 1u;
#pragma unroll
  for (int k = 0; k < KD / 8; ++k) {
    const uint32_t ko = k * 2 * L::LBO;
    const uint64_t dah = make_desc(a_hi + ko, L::LBO, L::SBO), dal = make_desc(a_lo + ko, L::LBO, L::SBO);
    const uint64_t dbh = make_desc(b_hi + ko, L::LBO, L::SBO), dbl = make_desc(b_lo + ko, L::LBO, L::SBO);
    mma_tf32(tmem_d, dal, dbh, idesc, acc);
    mma_tf32(tmem_d, dah, dbl, idesc, 1u);
    mma_tf32(tmem_d, dah, dbh, idesc, 1u);
    acc = 1u;
  }
}

}  // namespace umma

// ---------------------------------------------------------------------------------------------------------
// self-test kernel (mvin_test_umma_gemm): C[M, D] = A[M, D] . W[D, D]^T, one 128-row tile per CTA
// ---------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) umma_gemm_test_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                             float* __restrict__ C, long M) {
  pdl_enter();
  using L = umma::OpLayout<D>;
  constexpr int LPR = D / 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* a_hi = smem_raw;
  unsigned char* a_lo = a_hi + L::bytes(128);
  unsigned char* w_hi = a_lo + L::bytes(128);
  unsigned char* w_lo = w_hi + L::bytes(D);
  uint64_t* bar = reinterpret_cast<uint64_t*>(w_lo + L::bytes(D));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  if (warp == 0) umma::tmem_alloc(tmem_slot, D < 32 ? 32 : D);
  if (tid == 32) {
    umma::mbar_init(bar, 1);
    umma::fence_barrier_init();
  }
  umma::stage_weight<D>(w_hi, w_lo, W, D, tid, 256);
  const long row0 = (long)blockIdx.x * 128;
  for (int i = tid; i < 128 * LPR; i += 256) {
    const int r = i / LPR, k4 = i % LPR;
    float4 x = f4zero();
    if (row0 + r < M) x = ldg4(A + (row0 + r) * D + k4 * 4);
    umma::store_split<D>(a_hi, a_lo, r, k4, x);
  }
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    umma::issue_3xtf32<D>(tmem, umma::smem_u32(a_hi), umma::smem_u32(a_lo), umma::smem_u32(w_hi), umma::smem_u32(w_lo), D,
                          true);
    umma::commit(bar);
  }
  umma::mbar_wait(bar, 0);
  umma::fence_after_sync();
  // epilogue: warps 0-3 read columns [0, D/2), warps 4-7 columns [D/2, D); thread = row 32 (warp % 4) + lane
  const int r = 32 * (warp % 4) + lane;
  const int c0 = (warp / 4) * (D / 2);
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
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, D < 32 ? 32 : D);
}

// self-test / probe of the transposed product (weight gradients): dump[128][D] = raw accumulator lanes of
// dW = A^T . G accumulated over all 128-row tiles of A, G [M, D] by ONE CTA.
//   variant bit 0: A operand = MN-major view of the K-major tile (else a transposed, K-major staged copy)
//   variant bit 1: B operand likewise;   bit 2: MMA M = 128 instead of 64
template <int D>
__global__ void __launch_bounds__(256) umma_dw_test_kernel(const float* __restrict__ A, const float* __restrict__ G,
                                                           float* __restrict__ dump, long M, int variant) {
  pdl_enter();
  using L = umma::OpLayout<D>;
  using LT = umma::OpLayout<128>;                        // transposed copies: K = 128 rows of the tile
  constexpr int LPR = D / 4;
  const int MM = (variant & 4) ? 128 : 64;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* a_hi = smem_raw;                        // probe: single-product TF32 (hi parts only)
  unsigned char* g_hi = a_hi + L::bytes(128);
  unsigned char* at_hi = g_hi + L::bytes(128) + 1024;    // [64 (m, zero beyond D)][128 (k = row)]
  unsigned char* gt_hi = at_hi + LT::bytes(64);          // [D (n)][128]
  uint64_t* bar = reinterpret_cast<uint64_t*>(gt_hi + LT::bytes(D));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  if (warp == 0) umma::tmem_alloc(tmem_slot, D < 32 ? 32 : D);
  if (tid == 32) {
    umma::mbar_init(bar, 1);
    umma::fence_barrier_init();
  }
  for (int i = tid; i < LT::bytes(64) / 4; i += 256) reinterpret_cast<float*>(at_hi)[i] = 0.f;
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  uint32_t idesc = umma::idesc_tf32(MM, D);
  if (variant & 1) idesc |= umma::IDESC_A_MN;
  if (variant & 2) idesc |= umma::IDESC_B_MN;
  uint32_t phase = 0;
  const long ntiles = (M + 127) / 128;
  for (long t = 0; t < ntiles; ++t) {
    const long row0 = t * 128;
    for (int i = tid; i < 128 * LPR; i += 256) {
      const int r = i / LPR, k4 = i % LPR;
      float4 x = f4zero(), g = f4zero();
      if (row0 + r < M) {
        x = ldg4(A + (row0 + r) * D + k4 * 4);
        g = ldg4(G + (row0 + r) * D + k4 * 4);
      }
      float4 hi, lo;
      umma::split4(x, hi, lo);
      *reinterpret_cast<float4*>(a_hi + L::off(r, k4)) = hi;
      umma::split4(g, hi, lo);
      *reinterpret_cast<float4*>(g_hi + L::off(r, k4)) = hi;
      const float xs[4] = {x.x, x.y, x.z, x.w}, gs[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int mcol = k4 * 4 + j;
        const int o = LT::off(mcol, r / 4) + (r % 4) * 4;
        *reinterpret_cast<float*>(at_hi + o) = xs[j];
        *reinterpret_cast<float*>(gt_hi + o) = gs[j];
      }
    }
    umma::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_sync();
      uint32_t acc = t == 0 ? 0u : 1u;
      for (int k = 0; k < 128 / 8; ++k) {                  // K = rows of the tile, 8 per instruction
        // MN-major view: leading (K-group) offset = SBO of the K-major tile, stride (MN chunk) offset = its LBO
        const uint32_t kmn = k * L::SBO, kk = k * 2 * LT::LBO;
        const uint64_t da = (variant & 1) ? umma::make_desc(umma::smem_u32(a_hi) + kmn, L::SBO, L::LBO)
                                          : umma::make_desc(umma::smem_u32(at_hi) + kk, LT::LBO, LT::SBO);
        const uint64_t dg = (variant & 2) ? umma::make_desc(umma::smem_u32(g_hi) + kmn, L::SBO, L::LBO)
                                          : umma::make_desc(umma::smem_u32(gt_hi) + kk, LT::LBO, LT::SBO);
        umma::mma_tf32(tmem, da, dg, idesc, acc);
        acc = 1u;
      }
      umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
  }
  const int r = 32 * (warp % 4) + lane;
  const int c0 = (warp / 4) * (D / 2);
#pragma unroll
  for (int cc = 0; cc < D / 2; cc += 16) {
    float v[16];
    umma::tmem_ld16(tmem + ((uint32_t)(32 * (warp % 4)) << 16) + (uint32_t)(c0 + cc), v);
#pragma unroll
    for (int j = 0; j < 16; ++j) dump[(long)r * D + c0 + cc + j] = v[j];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, D < 32 ? 32 : D);
}
template <int D>
inline size_t umma_dw_test_smem() {
  return 2 * umma::OpLayout<D>::bytes(128) + 1024 + umma::OpLayout<128>::bytes(64) + umma::OpLayout<128>::bytes(D) + 16;
}

template <int D>
inline size_t umma_gemm_test_smem() {
  return 2 * umma::OpLayout<D>::bytes(128) + 2 * umma::OpLayout<D>::bytes(D) + 16;
}

}  // namespace mvin
