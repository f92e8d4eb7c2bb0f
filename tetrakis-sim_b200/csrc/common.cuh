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

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace tk
