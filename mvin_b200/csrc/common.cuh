// common.cuh -- device helpers shared by the MVIN sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MVIN_DEV __device__ __forceinline__
#define MVIN_HD __host__ __device__ __forceinline__
#define FULL_MASK 0xffffffffu

namespace mvin {

// Programmatic dependent launch (mvin_capi.cu launches every kernel with programmatic stream serialisation): the first
// statement of every kernel.  griddepcontrol.wait blocks until the preceding kernel of the stream has completed and its
// writes are visible (a no-op for a normal launch); launch_dependents then lets the NEXT kernel's CTAs become resident
// and sit at their own wait while this grid runs, so launch latency, CTA rasterisation and block start-up leave the
// critical path.  The trigger comes after the wait, so at most two kernels of a stream are co-resident.
MVIN_DEV void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

MVIN_DEV float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
  return v;
}
MVIN_DEV float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
// sum over the lanes of a power-of-two sub-group of width W (lanes share the high bits of the lane id)
template <int W>
MVIN_DEV float group_sum(float v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
// sum across the 32/W sub-groups of a warp (lanes with the same low bits)
template <int W>
MVIN_DEV float cross_group_sum(float v) {
#pragma unroll
  for (int o = 16; o >= W; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
template <int W>
MVIN_DEV float4 cross_group_sum4(float4 v) {
  v.x = cross_group_sum<W>(v.x);
  v.y = cross_group_sum<W>(v.y);
  v.z = cross_group_sum<W>(v.z);
  v.w = cross_group_sum<W>(v.w);
  return v;
}

MVIN_DEV float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
MVIN_DEV float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
MVIN_DEV void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// activation rows of a level much larger than L2 are touched once per kernel: streaming (evict-first) accesses keep
// them from flushing the entity table and the weights out of L2
MVIN_DEV float4 ld4a(const float* p, bool cs) {
  return cs ? __ldcs(reinterpret_cast<const float4*>(p)) : __ldg(reinterpret_cast<const float4*>(p));
}
MVIN_DEV void st4a(float* p, float4 v, bool cs) {
  if (cs) __stcs(reinterpret_cast<float4*>(p), v); else *reinterpret_cast<float4*>(p) = v;
}
MVIN_DEV void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
MVIN_DEV float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
MVIN_DEV float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
MVIN_DEV float4 f4scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
MVIN_DEV float4 f4fma(float s, float4 a, float4 c) {
  return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}
MVIN_DEV float f4dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// Bulk asynchronous L2 prefetch of a contiguous global range (cp.async.bulk.prefetch: the bulk-copy / TMA engine pulls
// the range into L2 without occupying the issuing thread or any shared memory).  The row kernels walk level buffers
// tile by tile -- 128 rows x 4d bytes, contiguous -- and issue the prefetch of the NEXT tile's streams while the
// current tile is computed, so that the dependent 16-byte row loads of the next iteration hit L2.
// addr 16-byte aligned, bytes a multiple of 16.
MVIN_DEV void bulk_prefetch_l2(const void* addr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(addr), "r"(bytes) : "memory");
}

// ---- packed fp32 pairs (sm_100: FADD2 / FMUL2 / FFMA2 process two floats per instruction; exact fp32 arithmetic) --------
struct f2x2 { unsigned long long lo, hi; };                // a float4 as two 64-bit register pairs: (x, y), (z, w)
MVIN_DEV unsigned long long pk2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
MVIN_DEV void upk2(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
MVIN_DEV f2x2 pack4(float4 v) { return f2x2{pk2(v.x, v.y), pk2(v.z, v.w)}; }
MVIN_DEV float4 unpack4(f2x2 v) {
  float4 r;
  upk2(v.lo, r.x, r.y);
  upk2(v.hi, r.z, r.w);
  return r;
}
MVIN_DEV unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
MVIN_DEV unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
MVIN_DEV unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
MVIN_DEV float gt0(float x) {                              // 1.0f if x > 0 else 0.0f, one instruction
  float r;
  asm("set.gt.f32.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(x));
  return r;
}

// ---- bulk asynchronous row copies global -> shared (cp.async.bulk, the non-tensor path of the TMA engine) with
// mbarrier transaction-count completion.  The gather kernels stage the neighbour rows of the NEXT step into a per-warp
// shared-memory ring while the current step is reduced: one 4d-byte copy instruction per row issued by one lane,
// instead of d/4 dependent 16-byte register loads per row, and no registers / scoreboard slots held by the rows in flight.
namespace bulk {
MVIN_DEV uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
MVIN_DEV void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(bar)), "r"(count) : "memory");
}
MVIN_DEV void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one arrival that also announces `bytes` of pending copies
MVIN_DEV void mbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(bytes) : "memory");
}
MVIN_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "BW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra BD_%=;\n\t"
      "bra BW_%=;\n\t"
      "BD_%=:\n\t"
      "}\n" ::"r"(saddr(bar)), "r"(parity)
      : "memory");
}
// src, dst 16-byte aligned, bytes a multiple of 16; completion is signalled on `bar` (complete_tx)
MVIN_DEV void copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(saddr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(saddr(bar))
               : "memory");
}
// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy (bulk copy) writes
MVIN_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}  // namespace bulk

// vectorised reduction to global memory: one 16-byte red instead of four scalar atomics (sm_90+)
MVIN_DEV void red_add4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

MVIN_DEV void red_add2(float* p, float x, float y) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(x), "f"(y) : "memory");
}

// Entity table / entity-gradient table, optionally row-sharded over 2^shift shards: entity e lives in shard
// (e & mask) at local row (e >> shift).  `shards` is a device array of base pointers -- the local shard and, on a
// multi-GPU box, the peers' shards mapped through CUDA IPC (NVLink peer loads / peer reductions).  shards == nullptr
// means one contiguous table at `base`.
struct ETab { const float* base; const float* const* shards; int shift, mask; };
struct GTab { float* base; float* const* shards; int shift, mask; };
MVIN_DEV const float* erow(const ETab& t, long e, int D) {
  if (t.shards == nullptr) return t.base + e * D;
  const unsigned long long p = __ldg(reinterpret_cast<const unsigned long long*>(t.shards) + (e & t.mask));
  return reinterpret_cast<const float*>(p) + (e >> t.shift) * D;
}
MVIN_DEV float* grow_of(const GTab& t, long e, int D) {
  if (t.shards == nullptr) return t.base + e * D;
  const unsigned long long p = __ldg(reinterpret_cast<const unsigned long long*>(t.shards) + (e & t.mask));
  return reinterpret_cast<float*>(p) + (e >> t.shift) * D;
}

}  // namespace mvin
