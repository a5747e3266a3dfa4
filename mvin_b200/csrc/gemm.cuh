// gemm.cuh -- small generic fp32 SIMT GEMM used for the dense side-maps of the MVIN path
// (Q = V.RK[r], user_o = O.W_user + b, item = concat.W_mix + b and their backward counterparts).
// FP32 FFMA on purpose: parity is 1e-4 relative on the final scores versus an fp32 reference, which plain
// TF32 tensor-core math does not meet (hard part 1).  These maps are tall and skinny
// (M = batch, N and K a few multiples of d), so the tile is chosen per problem for CTA count, not for reuse:
// the HBM/L2-bound gather kernels in level.cuh / ripple.cuh are the hot part.
#pragma once
#include "common.cuh"

namespace mvin {

struct GemmArgs {
  const float* A;         // A(m,k) = A[rowA(m) * sa_m + k * sa_k + batch * bsA], rowA(m) = a_rows ? a_rows[m] : m
  long sa_m, sa_k, bsA;
  const int32_t* a_rows;
  const float* B;         // B(k,n) = B[k * sb_k + n * sb_n + batch * bsB]
  long sb_k, sb_n, bsB;
  float* C;               // C(m,n) = C[rowC(m) * ldc + n + batch * bsC]
  long ldc, bsC;
  const int32_t* c_rows;
  const float* bias;      // [N] or nullptr; added once (k-split 0)
  int M, N, K;
  int nbatch, ksplit;
  int reduce;             // 1: sum over the nbatch batches inside the CTA (C has no batch dimension); ksplit then
                          //    splits the BATCH range over CTAs instead of K
  int accumulate;         // 1: atomicAdd into C (required when ksplit > 1 or c_rows has duplicates)
  float alpha;
};

constexpr int GEMM_THREADS = 256;

// CTA tile BM x BN, k-step BK; thread (ty, tx) owns TMR rows x 4 columns.
template <int BM, int BN, int BK>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_kernel(GemmArgs g) {
  pdl_enter();
  constexpr int TX = BN / 4, TY = GEMM_THREADS / TX, TMR = BM / TY;
  static_assert(TMR >= 1 && TMR * TY == BM, "bad tile");
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int batch = g.reduce ? 0 : blockIdx.z / g.ksplit, split = blockIdx.z % g.ksplit;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int kchunk = (g.K + g.ksplit - 1) / g.ksplit;
  kchunk = (kchunk + BK - 1) / BK * BK;
  int k_begin = split * kchunk;
  int k_end = min(g.K, k_begin + kchunk);
  int red_begin = 0, red_end = 1;
  if (g.reduce) {
    const int per = (g.nbatch + g.ksplit - 1) / g.ksplit;
    red_begin = split * per;
    red_end = min(g.nbatch, red_begin + per);
    k_begin = 0;
    k_end = g.K;
  }
  const float* A0 = g.A + (long)batch * g.bsA;
  const float* B0 = g.B + (long)batch * g.bsB;
  const bool a_kfast = (g.sa_k == 1), b_nfast = (g.sb_n == 1);

  float acc[TMR][4];
#pragma unroll
  for (int i = 0; i < TMR; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

  for (int red = red_begin; red < red_end; ++red) {
    const float* A = A0 + (long)red * g.bsA;
    const float* B = B0 + (long)red * g.bsB;
    for (int kt = k_begin; kt < k_end; kt += BK) {
      for (int idx = tid; idx < BM * BK; idx += GEMM_THREADS) {
        int mm, kk;
        if (a_kfast) { mm = idx / BK; kk = idx % BK; } else { mm = idx % BM; kk = idx / BM; }
        const int m = m0 + mm, k = kt + kk;
        float v = 0.f;
        if (m < g.M && k < k_end) {
          const long row = g.a_rows ? (long)g.a_rows[m] : (long)m;
          v = A[row * g.sa_m + (long)k * g.sa_k];
        }
        As[kk][mm] = v;
      }
      for (int idx = tid; idx < BN * BK; idx += GEMM_THREADS) {
        int nn, kb;
        if (b_nfast) { nn = idx % BN; kb = idx / BN; } else { nn = idx / BK; kb = idx % BK; }
        const int n = n0 + nn, k2 = kt + kb;
        float w = 0.f;
        if (n < g.N && k2 < k_end) w = B[(long)k2 * g.sb_k + (long)n * g.sb_n];
        Bs[kb][nn] = w;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 b = ld4(&Bs[kk][tx * 4]);
        float av[TMR];
        if constexpr (TMR == 4) {
          const float4 a = ld4(&As[kk][ty * 4]);
          av[0] = a.x; av[1] = a.y; av[2] = a.z; av[3] = a.w;
        } else {
#pragma unroll
          for (int i = 0; i < TMR; ++i) av[i] = As[kk][ty * TMR + i];
        }
#pragma unroll
        for (int i = 0; i < TMR; ++i) {
          acc[i][0] = fmaf(av[i], b.x, acc[i][0]);
          acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
          acc[i][2] = fmaf(av[i], b.z, acc[i][2]);
          acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
        }
      }
      __syncthreads();
    }
  }

  float* C = g.C + (long)batch * g.bsC;
#pragma unroll
  for (int i = 0; i < TMR; ++i) {
    const int m = m0 + ty * TMR + i;
    if (m >= g.M) continue;
    const long row = g.c_rows ? (long)g.c_rows[m] : (long)m;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = g.alpha * acc[i][j];
      if (g.bias && split == 0) v += g.bias[n];
      float* dst = &C[row * g.ldc + n];
      if (g.accumulate) atomicAdd(dst, v); else *dst = v;
    }
  }
}

}  // namespace mvin
