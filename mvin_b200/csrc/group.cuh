// group.cuh -- the table-gather level of the entity-table mode (table.cuh), evaluated per ENTITY GROUP.
//
// Reference (src/model/MVIN/): the neighbour aggregation of aggregator iteration 1 over the level whose children are the
// deepest level (aggregators.py:121-146 as called by model.py:295-306) and its TF autodiff.
//
// In the entity-table mode the children of a level-(H-2) node are relu(A[adj[e][k]] + C[pair]) -- the table rows and the
// attention p_k depend on the node's ENTITY e only, the pair only adds C.  A batch holds the same entity at that level
// many times (4.6x on average at C4, hundreds of times for popular entities), so the rows of the level are counting-sorted
// by entity once per step and each warp walks a window of 32 sorted rows: for every run of equal entities it gathers the
// K table rows ONCE into registers, applies them to every row of the run (element-wise: add C, ReLU, weight, sum), and in
// the backward accumulates the children's pre-activation gradients over the run in registers before ONE
// red.global.add.v4 per child -- table gathers and reductions drop by the mean run length.
//   forward :  Y[row] = self[row] + (1/K) sum_k p_k relu(A[n_k] + C[pair])              (then agg_fwd_kernel, `preagg`)
//   backward:  g = grow[row] (left by agg_bwd_kernel, `defer`);  x_k = relu(A[n_k] + C[pair]);
//              dA[n_k] += p_k g * [x_k > 0];  dCs[pair] += sum_k (same);  dp_k = g . x_k;
//              ds[rel_k] += p_k (sum_rows dp_k - sum_j p_j sum_rows dp_j)
// Lane (g, c) = (lane / LPR, lane % LPR) holds the float4 column c of the children k = q G + g, q < KPL <= 16.
#pragma once
#include "level.cuh"

namespace mvin {

constexpr int GRP_NT = 128, GRP_NW = GRP_NT / 32, GRP_KPL = 16, GRP_WIN = 32;

// can the level be evaluated per entity group?  (children per lane group must fit the register file)
MVIN_HD bool grp_supported(int D, int K) { return D >= 8 && (K * (D / 4) + 31) / 32 <= GRP_KPL; }

// ---- counting sort of the rows of a level by entity ------------------------------------------------------------
static __global__ void grp_count_kernel(const int32_t* __restrict__ ent, long rows, int32_t* __restrict__ cnt) {
  pdl_enter();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) atomicAdd(cnt + ent[i], 1);
}
// exclusive prefix sum over n ints in three launches: per-block scans (SCAN_PER_BLOCK elements each) + block totals,
// a single-block scan of the totals (<= SCAN_PER_BLOCK of them: n <= 16.7 M), and the final add
constexpr int SCAN_NT = 1024, SCAN_EPT = 4, SCAN_PER_BLOCK = SCAN_NT * SCAN_EPT;
MVIN_DEV int block_exclusive_scan(int v, int* __restrict__ warp_tot, int tid, int& total) {
  const int lane = tid % 32, warp = tid / 32;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(FULL_MASK, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_tot[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int t = lane < SCAN_NT / 32 ? warp_tot[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(FULL_MASK, t, o);
      if (lane >= o) t += y;
    }
    warp_tot[lane] = t;                                  // inclusive totals of the warps
  }
  __syncthreads();
  total = warp_tot[SCAN_NT / 32 - 1];
  const int before = warp ? warp_tot[warp - 1] : 0;
  __syncthreads();
  return before + x - v;
}
static __global__ void __launch_bounds__(SCAN_NT) scan_block_kernel(const int32_t* __restrict__ in, long n,
                                                                    int32_t* __restrict__ out, int32_t* __restrict__ block_tot) {
  pdl_enter();
  __shared__ int warp_tot[32];
  const int tid = threadIdx.x;
  const long base = (long)blockIdx.x * SCAN_PER_BLOCK + (long)tid * SCAN_EPT;
  int v[SCAN_EPT], s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_EPT; ++j) { v[j] = base + j < n ? in[base + j] : 0; s += v[j]; }
  int total;
  int ex = block_exclusive_scan(s, warp_tot, tid, total);
#pragma unroll
  for (int j = 0; j < SCAN_EPT; ++j) { if (base + j < n) out[base + j] = ex; ex += v[j]; }
  if (tid == 0 && block_tot) block_tot[blockIdx.x] = total;
}
static __global__ void __launch_bounds__(SCAN_NT) scan_add_kernel(int32_t* __restrict__ out, long n,
                                                                  const int32_t* __restrict__ block_off) {
  pdl_enter();
  const int add = block_off[blockIdx.x];
  const long base = (long)blockIdx.x * SCAN_PER_BLOCK + (long)threadIdx.x * SCAN_EPT;
#pragma unroll
  for (int j = 0; j < SCAN_EPT; ++j)
    if (base + j < n) out[base + j] += add;
}
// order[off[e] + (arrival index)] = row; esort likewise = e.  `fill` starts at zero (it is the count array re-zeroed).
static __global__ void grp_fill_kernel(const int32_t* __restrict__ ent, long rows, const int32_t* __restrict__ off,
                                       int32_t* __restrict__ fill, int32_t* __restrict__ order, int32_t* __restrict__ esort) {
  pdl_enter();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const int e = ent[i];
  const int pos = off[e] + atomicAdd(fill + e, 1);
  order[pos] = (int32_t)i;
  esort[pos] = e;
}

// ---- per-group evaluation ----------------------------------------------------------------------------------------
struct GroupArgs {
  const int32_t* order;   // [rows] row index of every sorted position
  const int32_t* esort;   // [rows] its entity
  const int32_t* adj;
  const float* s;         // [n_rel] relation scores of the aggregator iteration this level belongs to
  const float* tab;       // [n_entity, D] A_{h+1}
  const float* Cp;        // [B, D]        C_{h+1}
  const float* self;      // fwd: [rows, D]
  float* Y;               // fwd out: [rows, D] self + agg
  const float* gp;        // bwd: [rows, D] grow = gs / K of every row
  float* dtab;            // bwd: [n_entity, D] (+=)
  float* dCs;             // bwd: [B, D] (+=)
  float* ds;              // bwd: [n_rel] (+=)
  // `fuse` (H >= 3): the rows of this level are not materialised either -- self[row] = relu(tab_self[e] + Cp_self[pair])
  // is evaluated on the fly (per run / per row), and its gradient K gp * [self > 0] is summed per run into dtab_self
  // and per row into dCs_self, so that neither V[1][h] nor its two gradient buffers exist
  const float* tab_self;  // [n_entity, D] A_h
  const float* Cp_self;   // [B, D]        C_h
  float* dtab_self;       // bwd: [n_entity, D] (+=)
  float* dCs_self;        // bwd: [B, D] (+=)
  int fuse;
  long rows;
  unsigned long long rpp_magic;
  int K, n_rel;
};

inline size_t grp_smem(int n_rel, bool bwd) { return sizeof(float) * (size_t)n_rel * (bwd ? 1 + GRP_NW : 1); }

// SP (backward only): the children of a run are split over SP warps walking the same window (slot q' of warp `sub`
// is child slot q' SP + sub), which halves the per-lane register arrays and doubles the resident warps; every warp
// applies the softmax correction of ITS share of sum_k p_k dp_k to all K relations, so the warps never communicate.
// KC: compile-time number of child slots per lane group (4, 8 or 16, the smallest >= ceil(K / G)).  The slot loops are
// straight-line code over all KC slots -- slots beyond K carry p = 0 and a zero table row -- because a run-time bound
// turns every slot into its own basic block and keeps the scheduler from interleaving the 16 independent chains.
template <int D, bool BWD, int SP = 1, int KC = GRP_KPL>
__global__ void __launch_bounds__(GRP_NT, BWD ? (SP == 2 || KC <= 8 ? 4 : 2) : 4) virt_group_kernel(GroupArgs a) {
  pdl_enter();
  constexpr int LPR = D / 4, G = 32 / LPR, KPL_T = KC / SP;
  extern __shared__ __align__(16) float smem[];
  float* s_s = smem;
  float* ds_s = s_s + a.n_rel;                             // bwd: [NW][n_rel]
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, g = lane / LPR, c = lane % LPR;
  for (int i = tid; i < a.n_rel; i += GRP_NT) s_s[i] = a.s[i];
  if (BWD)
    for (int i = tid; i < GRP_NW * a.n_rel; i += GRP_NT) ds_s[i] = 0.f;
  __syncthreads();
  float* ds_w = ds_s + warp * a.n_rel;
  const int K = a.K;
  const int kpl = (K + G - 1) / G;                         // children per lane group (<= GRP_KPL)
  const float invK = 1.f / (float)K;
  const long nwin = (a.rows + GRP_WIN - 1) / GRP_WIN;
  const long gw = (long)blockIdx.x * GRP_NW + warp;
  const int sub = (int)(gw % SP);
  for (long w = gw / SP; w < nwin; w += (long)gridDim.x * GRP_NW / SP) {
    const long pos0 = w * GRP_WIN;
    const int nrow = (int)min((long)GRP_WIN, a.rows - pos0);
    const int my_e = lane < nrow ? __ldg(a.esort + pos0 + lane) : -1;
    const int my_row = lane < nrow ? __ldg(a.order + pos0 + lane) : 0;
    int i = 0;
    while (i < nrow) {                                     // warp-uniform
      const int e = __shfl_sync(FULL_MASK, my_e, i);
      // the run of equal entities starting at position i (positions are sorted, so it is contiguous)
      const unsigned same = __ballot_sync(FULL_MASK, my_e == e && lane >= i);
      const int n = __popc(same);
      // adjacency record, attention, table rows of the K children: once per run.  (Requesting the next run's record and
      // the next window's rows ahead of time was measured and made the kernel slower: 0.76 -> 0.81 ms at C4.)
      const AdjRec rec = load_adj(a.adj, e, K, lane);
      float p0, p1;
      {
        const float l0 = lane < K ? s_s[rec.rel0] : -INFINITY;
        const float l1 = lane + 32 < K ? s_s[rec.rel1] : -INFINITY;
        const float mx = warp_max(fmaxf(l0, l1));
        const float e0 = lane < K ? expf(l0 - mx) : 0.f;
        const float e1 = lane + 32 < K ? expf(l1 - mx) : 0.f;
        const float inv = 1.f / warp_sum(e0 + e1);
        p0 = e0 * inv;
        p1 = e1 * inv;
      }
      f2x2 Ap[KPL_T];                                      // table rows of the children, as packed fp32 pairs
      float pq[KPL_T];
      int idq[KPL_T];
#pragma unroll
      for (int q = 0; q < KPL_T; ++q) {
        const int k = (q * SP + sub) * G + g;              // child of this lane group
        const int src = k & 31;
        const int id_lo = __shfl_sync(FULL_MASK, rec.id0, src), id_hi = __shfl_sync(FULL_MASK, rec.id1, src);
        const float p_lo = __shfl_sync(FULL_MASK, p0, src), p_hi = __shfl_sync(FULL_MASK, p1, src);
        const bool on = q * SP + sub < kpl && k < K;
        idq[q] = k < 32 ? id_lo : id_hi;
        pq[q] = on ? (k < 32 ? p_lo : p_hi) : 0.f;
        Ap[q] = pack4(on ? ldg4(a.tab + (long)idq[q] * D + c * 4) : f4zero());
      }
      if (!BWD) {
        // rows of the run, the next row's vectors requested while the current row is reduced
        long row = __shfl_sync(FULL_MASK, my_row, i);
        const float4 a_self = a.fuse ? ldg4(a.tab_self + (long)e * D + c * 4) : f4zero();
        // sv: the self row, or (fuse) the pair vector C_h it is rebuilt from
        float4 cv = ldg4(a.Cp + fastdiv(row, a.rpp_magic) * D + c * 4);
        float4 sv = a.fuse ? ldg4(a.Cp_self + fastdiv(row, a.rpp_magic) * D + c * 4) : ld4(a.self + row * D + c * 4);
        for (int t = 0; t < n; ++t) {
          long row_n = row;
          float4 cv_n = cv, sv_n = sv;
          if (t + 1 < n) {
            row_n = __shfl_sync(FULL_MASK, my_row, i + t + 1);
            cv_n = ldg4(a.Cp + fastdiv(row_n, a.rpp_magic) * D + c * 4);
            sv_n = a.fuse ? ldg4(a.Cp_self + fastdiv(row_n, a.rpp_magic) * D + c * 4) : ld4(a.self + row_n * D + c * 4);
          }
          if (a.fuse) {
            const float4 x = f4add(a_self, sv);
            sv = make_float4(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f), fmaxf(x.z, 0.f), fmaxf(x.w, 0.f));
          }
          // packed fp32 pairs (FADD2 / FFMA2): x = A + C, acc += p relu(x)
          const f2x2 cvp = pack4(cv);
          f2x2 acc0{0ull, 0ull}, acc1{0ull, 0ull};
#pragma unroll
          for (int q = 0; q < KPL_T; q += 2) {
            {
              float4 x = unpack4(f2x2{add2(Ap[q].lo, cvp.lo), add2(Ap[q].hi, cvp.hi)});
              const unsigned long long pp = pk2(pq[q], pq[q]);
              acc0.lo = fma2(pp, pk2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)), acc0.lo);
              acc0.hi = fma2(pp, pk2(fmaxf(x.z, 0.f), fmaxf(x.w, 0.f)), acc0.hi);
            }
            if (KPL_T > 1) {
              float4 x = unpack4(f2x2{add2(Ap[q + 1].lo, cvp.lo), add2(Ap[q + 1].hi, cvp.hi)});
              const unsigned long long pp = pk2(pq[q + 1], pq[q + 1]);
              acc1.lo = fma2(pp, pk2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)), acc1.lo);
              acc1.hi = fma2(pp, pk2(fmaxf(x.z, 0.f), fmaxf(x.w, 0.f)), acc1.hi);
            }
          }
          const float4 acc = cross_group_sum4<LPR>(unpack4(f2x2{add2(acc0.lo, acc1.lo), add2(acc0.hi, acc1.hi)}));
          if (g == 0) st4(a.Y + row * D + c * 4, f4fma(invK, acc, sv));
          row = row_n; cv = cv_n; sv = sv_n;
        }
      } else {
        // W dot products are reduced together by the butterfly reduce-scatter of level.cuh (W - 1 shuffles for W sums):
        // lane c of a group ends up with dp of child slot q = ch W + c % W
        constexpr int W = LPR < KPL_T ? LPR : KPL_T, NCH = KPL_T / W;
        float myp[NCH];
        int myrel[NCH];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const int k = ((ch * W + c % W) * SP + sub) * G + g;
          const int src = k & 31;
          const float p_lo = __shfl_sync(FULL_MASK, p0, src), p_hi = __shfl_sync(FULL_MASK, p1, src);
          const int r_lo = __shfl_sync(FULL_MASK, rec.rel0, src), r_hi = __shfl_sync(FULL_MASK, rec.rel1, src);
          const bool mine = k < K && (ch * W + c % W) * SP + sub < kpl && c < W;
          myp[ch] = mine ? (k < 32 ? p_lo : p_hi) : 0.f;
          myrel[ch] = mine ? (k < 32 ? r_lo : r_hi) : 0;
        }
        f2x2 acc[KPL_T];
        float dpsl[NCH];
#pragma unroll
        for (int q = 0; q < KPL_T; ++q) acc[q] = f2x2{0ull, 0ull};
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) dpsl[ch] = 0.f;
        long row = __shfl_sync(FULL_MASK, my_row, i);
        long pair = fastdiv(row, a.rpp_magic);
        float4 cv = ldg4(a.Cp + pair * D + c * 4);
        float4 gr = ld4(a.gp + row * D + c * 4);
        const bool fuse = a.fuse != 0;
        const float4 a_self = fuse ? ldg4(a.tab_self + (long)e * D + c * 4) : f4zero();
        float4 cs_self = fuse ? ldg4(a.Cp_self + pair * D + c * 4) : f4zero();
        float4 acc_self = f4zero();
        const float Kf = (float)K;
        for (int t = 0; t < n; ++t) {
          long pair_n = pair;
          float4 cv_n = cv, gr_n = gr, cs_self_n = cs_self;
          if (t + 1 < n) {
            const long row_n = __shfl_sync(FULL_MASK, my_row, i + t + 1);
            pair_n = fastdiv(row_n, a.rpp_magic);
            cv_n = ldg4(a.Cp + pair_n * D + c * 4);
            gr_n = ld4(a.gp + row_n * D + c * 4);
            if (fuse) cs_self_n = ldg4(a.Cp_self + pair_n * D + c * 4);
          }
          if (fuse && sub == 0) {
            // gradient of the row's own iteration-0 output: dself = K gp, masked by self = relu(A_h[e] + C_h[pair]) > 0
            const float4 x = f4add(a_self, cs_self);
            const float4 ds4 = make_float4(x.x > 0.f ? Kf * gr.x : 0.f, x.y > 0.f ? Kf * gr.y : 0.f, x.z > 0.f ? Kf * gr.z : 0.f,
                                           x.w > 0.f ? Kf * gr.w : 0.f);
            acc_self = f4add(acc_self, ds4);
            if (g == 0) red_add4(a.dCs_self + pair * D + c * 4, ds4);
          }
          // acc[q] collects  sum_rows [x > 0] g  (the factor p_q is applied once, at the flush);  m collects  sum_q p_q [x > 0]
          // packed fp32 pairs (FADD2 / FMUL2 / FFMA2); the ReLU mask as 1.0 / 0.0 so that the accumulations are FMAs
          const f2x2 cvp = pack4(cv), grp = pack4(gr);
          f2x2 mp{0ull, 0ull};
          float part[KPL_T];
#pragma unroll
          for (int q = 0; q < KPL_T; ++q) {
            {
              const float4 x = unpack4(f2x2{add2(Ap[q].lo, cvp.lo), add2(Ap[q].hi, cvp.hi)});
              const unsigned long long d2 = fma2(grp.hi, pk2(fmaxf(x.z, 0.f), fmaxf(x.w, 0.f)),
                                                 mul2(grp.lo, pk2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f))));
              float d0, d1;
              upk2(d2, d0, d1);
              part[q] = d0 + d1;
              const unsigned long long k_lo = pk2(gt0(x.x), gt0(x.y)), k_hi = pk2(gt0(x.z), gt0(x.w));
              const unsigned long long pp = pk2(pq[q], pq[q]);
              acc[q].lo = fma2(k_lo, grp.lo, acc[q].lo);
              acc[q].hi = fma2(k_hi, grp.hi, acc[q].hi);
              mp.lo = fma2(k_lo, pp, mp.lo);
              mp.hi = fma2(k_hi, pp, mp.hi);
            }
          }
          float4 cs = unpack4(f2x2{mul2(grp.lo, mp.lo), mul2(grp.hi, mp.hi)});
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            {
              float v[W];
#pragma unroll
              for (int j = 0; j < W; ++j) v[j] = part[ch * W + j];
              dpsl[ch] += reduce_scatter<W, LPR>(v, lane);
            }
          }
          cs = cross_group_sum4<LPR>(cs);
          if (g == 0) red_add4(a.dCs + pair * D + c * 4, cs);
          pair = pair_n; cv = cv_n; gr = gr_n; cs_self = cs_self_n;
        }
        if (fuse && sub == 0 && g == 0) red_add4(a.dtab_self + (long)e * D + c * 4, acc_self);
        // flush the run: one reduction per child, the softmax gradient once per (run, child)
#pragma unroll
        for (int q = 0; q < KPL_T; ++q)
          if (q * SP + sub < kpl && (q * SP + sub) * G + g < K)
            red_add4(a.dtab + (long)idq[q] * D + c * 4, f4scale(unpack4(acc[q]), pq[q]));
        float dotp = 0.f;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) dotp = fmaf(myp[ch], dpsl[ch], dotp);     // myp = 0 on the lanes that do not own a slot
        dotp = warp_sum(dotp);                             // this warp's share of sum_k p_k sum_rows dp_k
        // ds[rel_k] += p_k sum_rows dp_k  for the warp's own children ...
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
          if (myp[ch] != 0.f) atomicAdd(&ds_w[myrel[ch]], myp[ch] * dpsl[ch]);
        // ... and  ds[rel_k] -= p_k dotp  for ALL K children (lane l holds k = l and k = l + 32 of the record)
        if (lane < K) atomicAdd(&ds_w[rec.rel0], -p0 * dotp);
        if (lane + 32 < K) atomicAdd(&ds_w[rec.rel1], -p1 * dotp);
      }
      i += n;
    }
  }
  if (BWD) {
    __syncthreads();
    for (int i = tid; i < a.n_rel; i += GRP_NT) {
      float s = 0.f;
      for (int w = 0; w < GRP_NW; ++w) s += ds_s[w * a.n_rel + i];
      if (s != 0.f) atomicAdd(a.ds + i, s);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Per-entity leaf aggregate in the same register-blocked form (replaces leaf_entity_kernel of level.cuh where the
// children fit: grp_supported):  Se[e] = sum_k p_k E[n_k]  and its backward  dE[n_k] += p_k g,  dp_k = g . E[n_k],
// ds[rel_k] += p_k (dp_k - sum_j p_j dp_j)  with g = GSe[e].  One warp per stamped entity; lane (g, c) holds column c of
// the children k = q G + g: all K row loads of an entity are in flight at once, the K dot products are reduced by one
// butterfly reduce-scatter, and the slot loops are straight-line code (compile-time KC).
// ---------------------------------------------------------------------------------------------------------
template <int D, bool BWD, int KC>
__global__ void __launch_bounds__(LEAF_NT) leaf_entity_reg_kernel(LeafEntArgs a) {
  pdl_enter();
  constexpr int LPR = D / 4, G = 32 / LPR;
  extern __shared__ __align__(16) float smem[];
  float* s_s = smem;
  float* ds_s = s_s + a.n_rel;                             // bwd: [NWH][n_rel]
  const int NWH = leaf_ds_copies(a.n_rel);
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, g = lane / LPR, c = lane % LPR;
  for (int i = tid; i < a.n_rel; i += LEAF_NT) s_s[i] = a.s[i];
  if (BWD)
    for (int i = tid; i < NWH * a.n_rel; i += LEAF_NT) ds_s[i] = 0.f;
  __syncthreads();
  float* ds_w = ds_s + (NWH > 1 ? warp : 0) * a.n_rel;
  const int K = a.K, chunk = a.chunk;
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
      float4 x[KC];
      float pq[KC];
      int idq[KC];
#pragma unroll
      for (int q = 0; q < KC; ++q) {
        const int k = q * G + g, src = k & 31;
        const int id_lo = __shfl_sync(FULL_MASK, rec.id0, src), id_hi = __shfl_sync(FULL_MASK, rec.id1, src);
        const float p_lo = __shfl_sync(FULL_MASK, p0, src), p_hi = __shfl_sync(FULL_MASK, p1, src);
        const bool on = k < K;
        idq[q] = k < 32 ? id_lo : id_hi;
        pq[q] = on ? (k < 32 ? p_lo : p_hi) : 0.f;
        x[q] = on ? ldg4(erow(a.E, idq[q], D) + c * 4) : f4zero();
      }
      if (!BWD) {
        float4 acc = f4zero();
#pragma unroll
        for (int q = 0; q < KC; ++q) acc = f4fma(pq[q], x[q], acc);
        acc = cross_group_sum4<LPR>(acc);
        if (g == 0) st4(a.Se + e * D + c * 4, acc);
      } else {
        constexpr int W = LPR < KC ? LPR : KC, NCH = KC / W;
        const float4 gr = ld4(a.GSe + e * D + c * 4);
        float part[KC];
#pragma unroll
        for (int q = 0; q < KC; ++q) {
          part[q] = f4dot(gr, x[q]);
          if (q * G + g < K) red_add4(grow_of(a.dE, idq[q], D) + c * 4, f4scale(gr, pq[q]));
        }
        float dotp = 0.f, mydp[NCH], myp[NCH];
        int myrel[NCH];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          float v[W];
#pragma unroll
          for (int j = 0; j < W; ++j) v[j] = part[ch * W + j];
          mydp[ch] = reduce_scatter<W, LPR>(v, lane);      // dp of child slot ch W + c % W
          const int k = (ch * W + c % W) * G + g, src = k & 31;
          const float p_lo = __shfl_sync(FULL_MASK, p0, src), p_hi = __shfl_sync(FULL_MASK, p1, src);
          const int r_lo = __shfl_sync(FULL_MASK, rec.rel0, src), r_hi = __shfl_sync(FULL_MASK, rec.rel1, src);
          const bool mine = k < K && c < W;
          myp[ch] = mine ? (k < 32 ? p_lo : p_hi) : 0.f;
          myrel[ch] = mine ? (k < 32 ? r_lo : r_hi) : 0;
          dotp = fmaf(myp[ch], mydp[ch], dotp);
        }
        dotp = warp_sum(dotp);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
          if (myp[ch] != 0.f) atomicAdd(&ds_w[myrel[ch]], myp[ch] * (mydp[ch] - dotp));
      }
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
inline size_t leaf_entity_reg_smem(int n_rel, bool bwd) {
  return sizeof(float) * (size_t)n_rel * (bwd ? 1 + leaf_ds_copies(n_rel) : 1);
}

}  // namespace mvin
