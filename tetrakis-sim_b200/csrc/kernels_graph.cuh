// Generic-graph leapfrog step (any networkx graph) and the probe DFT.
//
// Graph step: reference tetrakis_sim/physics.py:106-129.  Sliced-ELL: rows are grouped in slices
// of 32 (one warp), each slice padded to its own maximum degree and stored column-major, so lane i
// reads idx[off + k*32 + i] -- coalesced -- and walks ITS neighbours in adjacency order.  Keeping
// that order and forbidding FMA contraction makes u(t) bit-identical to the reference.
// Padding entries point at a dummy slot (index n) that always holds +0.0; x + 0.0 == x for every
// partial sum the reference can produce (a sum that starts at +0.0 is never -0.0).
#pragma once
#include "common.cuh"

namespace tk {

template <bool KEN, bool PIPE = false>
__global__ void __launch_bounds__(256)
graph_step_kernel(long long n, const long long *__restrict__ slice_off,
                  const int *__restrict__ ell_idx, const double *__restrict__ deg,
                  const double *__restrict__ inv_mass, const double *__restrict__ potential,
                  const double *__restrict__ u, double *__restrict__ w, double coeff, double damping,
                  double *__restrict__ pen, double inv_coeff)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (i < n) {
    const long long slice = i >> 5;
    const int lane = (int)(i & 31);
    const long long off = slice_off[slice];
    const int width = (int)((slice_off[slice + 1] - off) >> 5);
    const int *idx = ell_idx + off + lane;
    double nsum = 0.0;  // physics.py:115
    int k = 0;
    if (PIPE) {
        // gather 4 ahead with the NEXT four indices already requested: the index stream (DRAM) and the gathers (cache)
        // of consecutive groups overlap instead of alternating; the adds stay strictly in adjacency order
        int j0 = 0, j1 = 0, j2 = 0, j3 = 0;
        if (width >= 4) { j0 = idx[0]; j1 = idx[32]; j2 = idx[64]; j3 = idx[96]; }
        for (; k + 4 <= width; k += 4) {
            const int i0 = j0, i1 = j1, i2 = j2, i3 = j3;
            if (k + 8 <= width) {
                j0 = idx[(size_t)(k + 4) * 32]; j1 = idx[(size_t)(k + 5) * 32];
                j2 = idx[(size_t)(k + 6) * 32]; j3 = idx[(size_t)(k + 7) * 32];
            }
            const double a = u[i0], b = u[i1], c = u[i2], d = u[i3];
            nsum = __dadd_rn(nsum, a);
            nsum = __dadd_rn(nsum, b);
            nsum = __dadd_rn(nsum, c);
            nsum = __dadd_rn(nsum, d);
        }
    } else {
        for (; k + 4 <= width; k += 4) {  // gather 4 ahead, add strictly in order
            const double a = u[idx[(size_t)k * 32]], b = u[idx[(size_t)(k + 1) * 32]];
            const double c = u[idx[(size_t)(k + 2) * 32]], d = u[idx[(size_t)(k + 3) * 32]];
            nsum = __dadd_rn(nsum, a);
            nsum = __dadd_rn(nsum, b);
            nsum = __dadd_rn(nsum, c);
            nsum = __dadd_rn(nsum, d);
        }
    }
    for (; k < width; ++k) nsum = __dadd_rn(nsum, u[idx[(size_t)k * 32]]);
    const double ui = u[i];
    const double lap = __dsub_rn(nsum, __dmul_rn(deg[i], ui));  // physics.py:119
    const double un = leap_exact(ui, w[i], __dmul_rn(coeff, inv_mass[i]), lap,
                                 __dmul_rn(coeff, potential[i]), damping);
    w[i] = un;
    if (KEN) e = (un - ui) * (un - ui) * inv_coeff / inv_mass[i] + un * (potential[i] * ui - lap);
    }
    if (KEN) {  // per-CTA partial of twice the staggered energy (fixed tree => deterministic)
        __shared__ double red[8];
        e = warp_sum(e);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
        __syncthreads();
        if (threadIdx.x == 0)
            pen[blockIdx.x] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
    }
}

// energy_out = 1/2 * sum(pen[0..n)) in a fixed order
__global__ void __launch_bounds__(256)
energy_reduce_kernel(const double *__restrict__ pen, long long n, double *__restrict__ energy_out,
                     const long long *__restrict__ ctr)
{
    if (ctr != nullptr) energy_out += ctr[2];  // graph replay: row from the device counter
    __shared__ double red[256];
    double a = 0.0;
    for (long long i = threadIdx.x; i < n; i += 256) a += pen[i];
    red[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *energy_out = 0.5 * red[0];
}

// |DFT| of real series, arbitrary length (reference physics.py:163: np.abs(np.fft.fft(values))).
// One CTA per output bin k; the twiddle angle is reduced exactly in integers ((k*n) mod N) before
// sincospi, so accuracy does not degrade with N.  Fixed-order tree => deterministic.
__global__ void __launch_bounds__(128)
dft_abs_kernel(const double *__restrict__ series, long long n, double *__restrict__ spectrum)
{
    const long long k = blockIdx.x;
    const double *x = series + (size_t)blockIdx.y * n;
    double re = 0.0, im = 0.0;
    for (long long j = threadIdx.x; j < n; j += blockDim.x) {
        const long long m = (k * j) % n;
        double s, c;
        sincospi(2.0 * (double)m / (double)n, &s, &c);
        re += x[j] * c;
        im -= x[j] * s;
    }
    __shared__ double sre[4], sim[4];
    re = warp_sum(re);
    im = warp_sum(im);
    if ((threadIdx.x & 31) == 0) {
        sre[threadIdx.x >> 5] = re;
        sim[threadIdx.x >> 5] = im;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double r = (sre[0] + sre[1]) + (sre[2] + sre[3]);
        const double i2 = (sim[0] + sim[1]) + (sim[2] + sim[3]);
        spectrum[(size_t)blockIdx.y * n + k] = hypot(r, i2);
    }
}

// np.argmax (first maximum) of each row; reference tetrakis_sim/batch.py:205.
__global__ void __launch_bounds__(256)
argmax_kernel(const double *__restrict__ spectrum, long long n, long long *__restrict__ out)
{
    const double *x = spectrum + (size_t)blockIdx.x * n;
    double bv = -1.0;
    long long bi = 0;
    for (long long j = threadIdx.x; j < n; j += blockDim.x)
        if (x[j] > bv) { bv = x[j]; bi = j; }
    __shared__ double sv[256];
    __shared__ long long si[256];
    sv[threadIdx.x] = bv;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double ov = sv[threadIdx.x + o];
            const long long oi = si[threadIdx.x + o];
            if (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) {
                sv[threadIdx.x] = ov;
                si[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = si[0];
}

}  // namespace tk
