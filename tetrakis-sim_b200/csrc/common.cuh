// Shared device helpers for the tetrakis-b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tk {

// The reference's update expression, reference tetrakis_sim/physics.py:120-126, evaluated left to
// right with one IEEE rounding per operation (no FMA contraction) -- bit-identical to CPython
// float arithmetic.  cm = coeff*inv_mass, cv = coeff*potential (each one rounded product).
__device__ __forceinline__ double leap_exact(double u, double up, double cm, double lap, double cv,
                                             double damping)
{
    double t = __dmul_rn(2.0, u);
    t = __dsub_rn(t, up);
    t = __dadd_rn(t, __dmul_rn(cm, lap));
    t = __dsub_rn(t, __dmul_rn(cv, u));
    t = __dsub_rn(t, __dmul_rn(damping, __dsub_rn(u, up)));
    return t;
}

// Sum 4 independent values across the 32 lanes of a warp with a transposing butterfly:
// 6 double shuffles instead of 20.  Fixed tree => deterministic.  Afterwards lane 8*q holds the
// total of v[q]; those four lanes store dst[0..3].
__device__ __forceinline__ void warp_sum4_store(const double v[4], int lane, double *dst)
{
    const bool h16 = (lane & 16) != 0;
    double a0 = h16 ? v[2] : v[0], a1 = h16 ? v[3] : v[1];
    const double s0 = h16 ? v[0] : v[2], s1 = h16 ? v[1] : v[3];
    a0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    a1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    const bool h8 = (lane & 8) != 0;
    double b = h8 ? a1 : a0;
    const double s = h8 ? a0 : a1;
    b += __shfl_xor_sync(0xffffffffu, s, 8);
    b += __shfl_xor_sync(0xffffffffu, b, 4);
    b += __shfl_xor_sync(0xffffffffu, b, 2);
    b += __shfl_xor_sync(0xffffffffu, b, 1);
    if ((lane & 7) == 0) dst[lane >> 3] = b;
}

// 256-bit global accesses (one 32-byte cell = 4 quadrant values per instruction; sm_100 LDG/STG.256).
__device__ __forceinline__ void ld256_nc(const double *p, double v[4])  // read-only for the kernel
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void ld256_cs(const double *p, double v[4])  // streamed once (evict-first)
{
    asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void st256_cs(double *p, const double v[4])
{
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
__device__ __forceinline__ void st256(double *p, const double v[4])
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

// 128-bit accesses for the fp32 state layout (one 16-byte cell = 4 quadrant values)
__device__ __forceinline__ float4 ld128_nc(const float *p)
{
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld128_cs(const float *p)
{
    float4 v;
    asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st128_cs(float *p, float4 v)
{
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st128(float *p, float4 v)
{
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// 256-bit fp32 accesses: two 16-byte cells per instruction
__device__ __forceinline__ void ld256f_nc(const float *p, float v[8])
{
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void ld256f_cs(const float *p, float v[8])
{
    asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p) : "memory");
}
__device__ __forceinline__ void st256f_cs(float *p, const float v[8])
{
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void st256f(float *p, const float v[8])
{
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}

// 16-byte asynchronous global->shared copy (LDGSTS); ok == false writes zeros and reads nothing.
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool ok)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    const int n = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(sa), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sa), "l"(gmem) : "memory");
}
// 8-byte asynchronous global->shared copy (LDGSTS.64): per-thread prefetch rings (kernels_cross.cuh)
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier
__device__ __forceinline__ void mbar_init(unsigned long long *b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes),
                    "r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity)
{
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 " @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
                 :: "r"((unsigned)__cvta_generic_to_shared(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace tk
