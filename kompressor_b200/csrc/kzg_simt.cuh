// Thin SIMT layer of the execute kernel (kzg_exec2.cuh): warp collectives, shared-memory accesses by 32-bit
// shared address, cp.async (LDGSTS) staging and the bulk asynchronous shared -> global copy (cp.async.bulk, SASS
// UBLKCP) with its proxy fence.  On the device every function is one or two PTX instructions.
//
// With KZG_SIMT_EMU defined (tests/hostsim only, never in libkzg.so) the same names resolve to a lock-step warp
// emulator on the CPU (tests/hostsim/simt_emu.h): 32 coroutines, one per lane, meet at every collective; asynchronous
// copies complete as late as their wait allows.  The kernel source is therefore testable without a GPU, including
// missing-barrier and early-reuse mistakes.
#pragma once
#include "kzg_common.cuh"

#if defined(KZG_SIMT_EMU)
#include "simt_emu.h"          // provides namespace simt for the host
#else

namespace simt {

#define KZG_FULL 0xFFFFFFFFu
__device__ __forceinline__ u32 w_shfl(u32 v, u32 src) { return __shfl_sync(KZG_FULL, v, src); }
__device__ __forceinline__ u32 w_shfl_up(u32 v, u32 d) { return __shfl_up_sync(KZG_FULL, v, d); }
__device__ __forceinline__ u32 w_ballot(bool p) { return __ballot_sync(KZG_FULL, p); }
__device__ __forceinline__ bool w_any(bool p) { return __any_sync(KZG_FULL, p); }
__device__ __forceinline__ u32 w_or(u32 v) { return __reduce_or_sync(KZG_FULL, v); }          // REDUX: OR over the warp
__device__ __forceinline__ void w_sync() { __syncwarp(); }
__device__ __forceinline__ void fence_block() { __threadfence_block(); }
__device__ __forceinline__ void fence_gpu() { __threadfence(); }

__device__ __forceinline__ u32 lds32(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ u32 lds8(u32 a) { u32 v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts32(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8(u32 a, u32 v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts32_if(bool p, u32 a, u32 v) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u32 [%0], %1;\n\t}" :: "r"(a), "r"(v), "r"((u32)p) : "memory");
}
__device__ __forceinline__ void sts8_if(bool p, u32 a, u32 v) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u8 [%0], %1;\n\t}" :: "r"(a), "r"(v), "r"((u32)p) : "memory");
}
__device__ __forceinline__ uint4 lds128(u32 a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(u32 a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(u32 a, u32 x, u32 y) { asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ uint2 lds64(u32 a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}

// cp.async (LDGSTS): 16 bytes global -> shared without a destination register; groups are per thread.
__device__ __forceinline__ void cp_async16_if(bool p, u32 dst, const void* src) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}" :: "r"(dst), "l"(src), "r"((u32)p) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// Bulk asynchronous copy shared -> global (16-byte aligned on both sides, size a multiple of 16), tracked by
// per-thread bulk groups.  fence_async() orders the warp's st.shared before the async proxy reads them.
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void* g, u32 s, u32 n) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(g), "r"(s), "r"(n) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }   // at most N groups pending
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }   // ... whose source has not been read yet
__device__ __forceinline__ void bulk_wait_all() { bulk_wait<0>(); }

// History this warp wrote earlier: read at L2 (the coherence point), never from a stale L1 line.
__device__ __forceinline__ u32 ldg32_hist(const u32* p) { return __ldcg(p); }
__device__ __forceinline__ u32 ldg8_hist(const u8* p) { return (u32)__ldcg(p); }
__device__ __forceinline__ u32 ldg8_ro(const u8* p) { return (u32)__ldg(p); }
__device__ __forceinline__ u32 ldg32_ro(const u32* p) { return __ldg(p); }
__device__ __forceinline__ u64 ldg64_stream(const u64* p) { return __ldcs(p); }
__device__ __forceinline__ void stg8(u8* p, u32 v) { *p = (u8)v; }
__device__ __forceinline__ void stg128(u8* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }
__device__ __forceinline__ u64 ldg64_ro(const u64* p) { return __ldg(reinterpret_cast<const unsigned long long*>(p)); }
__device__ __forceinline__ u32 atomic_add(u32* p, u32 v) { return atomicAdd(p, v); }

__device__ __forceinline__ u32 funnel_r(u32 lo, u32 hi, u32 sh) { return __funnelshift_r(lo, hi, sh); }
__device__ __forceinline__ u32 funnel_l(u32 lo, u32 hi, u32 sh) { return __funnelshift_l(lo, hi, sh); }   // high word of (hi:lo) << (sh & 31)
__device__ __forceinline__ u32 shl_clamp(u32 v, u32 n) { u32 r; asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n)); return r; }   // n >= 32 -> 0
__device__ __forceinline__ u32 shr_clamp(u32 v, u32 n) { u32 r; asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n)); return r; }
__device__ __forceinline__ u32 clz32(u32 v) { return (u32)__clz((int)v); }
__device__ __forceinline__ u32 ffs32(u32 v) { return (u32)__ffs((int)v); }

}  // namespace simt
#define KZG_SIMT_D __device__ __forceinline__
#endif
