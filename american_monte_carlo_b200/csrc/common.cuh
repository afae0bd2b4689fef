// Shared device helpers: streaming vector loads/stores, deterministic block reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace amc {

constexpr int kWarp = 32;

// ---- 128-bit streaming accesses ------------------------------------------------------------------------
// Path columns and the per-path state are touched once per launch: bypass L1 allocation so the 126 MB L2
// (which may still hold the column from the previous step's launch) is the only cache level involved.
__device__ __forceinline__ double2 ld_stream(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ld_stream(const int4* p) {
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ int2 ld_stream(const int2* p) {
    int2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
// state arrays are read AND written by the same launch: plain (coherent) load, no L1 allocation
__device__ __forceinline__ double2 ld_state(const double2* p) {
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(double2* p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_stream(int2* p, int2 v) {
    asm volatile("st.global.L1::no_allocate.v2.s32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// ---- reductions ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block reduction of NACC per-thread accumulators into out[0..NACC) (global memory, one row
// per block).  Order is fixed by (lane, warp) -> bitwise reproducible for a given launch geometry.
template <int NACC, int THREADS, int ROW = NACC>
__device__ __forceinline__ void block_reduce_store(const double (&acc)[NACC], double* smem /*[THREADS/32][NACC]*/,
                                                   double* out_row) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        double v = warp_sum(acc[a]);
        if (lane == 0) smem[warp * NACC + a] = v;
    }
    __syncthreads();
    if (threadIdx.x < ROW) {                // ROW > NACC: the unused tail of the row is written as zero
        double v = 0.0;
        if (threadIdx.x < NACC) {
#pragma unroll
            for (int w = 0; w < THREADS / 32; ++w) v += smem[w * NACC + threadIdx.x];
        }
        out_row[threadIdx.x] = v;
    }
}

}  // namespace amc

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------
// The sweep is a chain K3 -> K4 -> K3 -> ... of strictly dependent launches.  Launched with the
// programmatic-stream-serialization attribute, a kernel's blocks may become resident while its predecessor is
// still running; pdl_wait() then blocks until the predecessor has completed and flushed its memory.  Both are
// no-ops for a kernel launched the ordinary way.
namespace amc {
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
}  // namespace amc

// ---- TMA (bulk async copy) + mbarrier helpers, sm_90+/sm_100a -----------------------------------------------
// 1-D bulk copies need no tensor map: cp.async.bulk moves `bytes` (multiple of 16, 16-byte aligned on both
// sides) global -> shared and signals completion on an mbarrier via complete_tx.  SASS: UBLKCP.
namespace amc {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {      // never suspends
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// L2 eviction-priority policies (createpolicy): the 126 MB L2 is managed explicitly -- columns that the NEXT
// launch re-reads are kept (evict_last), data that is dead after this launch is streamed (evict_first).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_1d_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_addr(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void st_hint(double2* p, double2 v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(policy) : "memory");
}
__device__ __forceinline__ void st_hint(float2* p, float2 v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(policy) : "memory");
}
__device__ __forceinline__ void st_hint(double* p, double v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

}  // namespace amc

// ---- LL cells for the peer-memory all-reduce (see PeerArgs in kernels.h) ---------------------------------------
// A cell is 16 bytes {data_lo, flag, data_hi, flag}: every 8-byte half carries its own flag, so the protocol only
// needs 8-byte store atomicity (what NCCL's LL protocol relies on); volatile accesses go to L2 / the peer, never L1.
namespace amc {
__device__ __forceinline__ void st_ll(uint4* cell, double v, uint32_t flag) {
    const uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(cell), "r"(lo), "r"(flag), "r"(hi), "r"(flag)
                 : "memory");
}
__device__ __forceinline__ bool ld_ll(const uint4* cell, uint32_t flag, double& v) {
    uint32_t lo, f0, hi, f1;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(cell)
                 : "memory");
    v = __hiloint2double((int)hi, (int)lo);
    return f0 == flag && f1 == flag;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
}  // namespace amc
