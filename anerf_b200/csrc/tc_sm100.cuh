// Thin inline-PTX layer over the Blackwell (sm_100a) primitives this library uses:
// mbarrier, bulk async copy (TMA engine, 1-D), tcgen05 MMA / TMEM alloc / TMEM load, proxy fences.
// No CUTLASS dependency: operand layouts are produced by our own pack kernel and epilogues, so the
// shared-memory matrix descriptors are built by hand (K-major, no swizzle; see smem_desc()).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace anerf {

// ------------------------------------------------------------------------------------------------
// error reporting from device code: first failing wait records a code, then the kernel traps so
// that a protocol bug can never hang the GPU.
// ------------------------------------------------------------------------------------------------
struct DeviceStatus {
  unsigned int code;      // 0 = ok
  unsigned int where;     // site id
  unsigned int block;
  unsigned int thread;
};

enum : unsigned int {
  kErrWaitTimeout = 1,
};

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive with a count > 1 (one warp standing in for several expected arrivals)
__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// One probe of the barrier phase.  With the suspend-time hint the thread sleeps in hardware
// (SASS: TRYWAIT + NANOSLEEP.SYNCS) until the phase completes or ~`kSuspendNs` elapse, so a waiting
// thread does not burn issue slots of its SM sub-partition.
#ifndef ANERF_SUSPEND_NS
#define ANERF_SUSPEND_NS 100000
#endif
constexpr uint32_t kSuspendNs = ANERF_SUSPEND_NS;
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(kSuspendNs)
      : "memory");
  return ok != 0;
}
// Bounded wait: ~seconds at any clock, then record + trap (never spin forever on a protocol bug).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, DeviceStatus* st, unsigned int site) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (t0 == 0) { t0 = clock64(); continue; }
    if (clock64() - t0 > 6000000000LL) {
      if (st != nullptr && atomicCAS(&st->code, 0u, (unsigned)kErrWaitTimeout) == 0u) {
        st->where = site;
        st->block = blockIdx.x;
        st->thread = threadIdx.x;
        __threadfence_system();
      }
      __trap();
    }
  }
}

// ---- CTA pair (cluster of 2) ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// all threads of both CTAs
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster.
// Default (cta-scope release) semantics, as CUTLASS's ClusterBarrier::arrive(cta_id): what the arrival
// publishes is consumed by this CTA's own tensor core / async proxy (made visible by fence.proxy.async
// before the arrive); a cluster-scope release here costs a MEMBAR per chunk (measured: -15 % throughput).
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_cnt(uint64_t* bar, uint32_t cta, uint32_t count) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra], %2;\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta), "r"(count)
      : "memory");
}
// wait with cluster-scope acquire (for barriers that a peer CTA arrives on)
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(kSuspendNs)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, DeviceStatus* st, unsigned int site) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  long long t0 = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (t0 == 0) { t0 = clock64(); continue; }
    if (clock64() - t0 > 6000000000LL) {
      if (st != nullptr && atomicCAS(&st->code, 0u, (unsigned)kErrWaitTimeout) == 0u) {
        st->where = site;
        st->block = blockIdx.x;
        st->thread = threadIdx.x;
        __threadfence_system();
      }
      __trap();
    }
  }
}

// ---- fences -------------------------------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads, bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- bulk async copy global -> shared (TMA engine, SASS: UBLKCP) ----------------------------------
// size multiple of 16, both addresses 16-byte aligned; completion is signalled on `bar` as tx bytes.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- tensor memory ------------------------------------------------------------------------------
// Executed by one full warp.  Writes the TMEM base address to *smem_out.
// (cta_group::2: executed by one warp in EACH CTA of the pair)
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_out, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t of the warp owns lane
// 32*(warp%4)+t; register i is column col0+i).  Caller must tmem_ld_wait() before using v[].
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
// 32 lanes x 8 consecutive fp32 columns -> 8 registers per thread
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- tcgen05.mma ---------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand, no swizzle ("interleaved" canonical layout):
//   core matrix = 8 rows x 16 bytes, stored as 128 contiguous bytes;
//   SBO = byte distance between core matrices adjacent in the M/N direction,
//   LBO = byte distance between the two core matrices adjacent in K inside one K=16 (16-bit) slab.
// Bit layout (PTX ISA "matrix descriptor", sm_100): [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4,
// [46,48) version=1, [49,52) base offset=0, [61,64) swizzle mode=0.
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// Instruction descriptor for kind::f16: fp32 accumulate, A/B both K-major.
// fmt: 0 = fp16, 1 = bf16.  Bits: [4,6) D fmt (1=f32), [7,10) A fmt, [10,13) B fmt, [15] A major,
// [16] B major, [17,23) N>>3, [24,29) M>>4.
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t fmt_a, uint32_t fmt_b, uint32_t M, uint32_t N) {
  return (1u << 4) | (fmt_a << 7) | (fmt_b << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T over a CTA pair (M = 256: rows 0..127 in the leader's TMEM, 128..255
// in the peer's; each CTA supplies its own A rows and N/2 rows of B at the same shared-memory offsets).
// Issued by ONE thread of the leader CTA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on the mbarrier at this offset in BOTH CTAs of the pair when all previously issued MMAs of this
// thread have completed (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- split-precision operand packing --------------------------------------------------------------
// x = hi + lo with hi = round16(x), lo = round16(x - hi).  Three MMAs (hi*hi + lo*hi + hi*lo)
// then reproduce an fp32 product to ~2^-17 (bf16) / ~2^-22 (fp16); see DESIGN.md "precision".
template <int FMT> struct Split;
template <> struct Split<1> {  // bf16
  static __device__ __forceinline__ void pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);       // .x = a (low half), .y = b
    float2 hf = __bfloat1622float2(h);
    __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<uint32_t*>(&h);
    lo = *reinterpret_cast<uint32_t*>(&l);
  }
};
template <> struct Split<0> {  // fp16; conversions saturate at +-65504 so an out-of-range value never becomes inf/NaN
  static __device__ __forceinline__ uint32_t cvt2(float lo_half, float hi_half) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_half), "f"(lo_half));
    return r;
  }
  static __device__ __forceinline__ void pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = cvt2(a, b);
    float2 hf = __half22float2(*reinterpret_cast<__half2*>(&hi));
    lo = cvt2(a - hf.x, b - hf.y);
  }
};

// tcgen05 operand format code of a Split<FMT> (A and B must use the same format: a probe with bf16 hi
// parts and fp16 lo parts in one kind::f16 MMA raises 'illegal instruction' on sm_100a).
__host__ __device__ constexpr uint32_t fmt_hi(int FMT) { return FMT == 0 ? 0u : 1u; }
__host__ __device__ constexpr uint32_t fmt_lo(int FMT) { return FMT == 0 ? 0u : 1u; }

#endif  // __CUDACC__
}  // namespace anerf
