// O(N log N) spectrum of long / many probe series: Bluestein's chirp-z on top of a power-of-two Stockham FFT.
//
// reference tetrakis_sim/physics.py:162-163: spectrum = np.abs(np.fft.fft(values)) with len(values) = steps + 1, never a
// power of two.  The direct O(N^2) kernel (dft_abs_kernel, kernels_graph.cuh) is exact and cheap for one short series;
// for BASELINE config 2's 10 001 samples times hundreds of probes (whole-layer spectra, notebooks cells 12-13) it is
// 10^10..10^12 sincospi evaluations.  With w_m = exp(-i pi m^2 / N) and nk = (n^2 + k^2 - (k-n)^2) / 2:
//     X[k] = w_k * sum_n (x_n w_n) conj(w_{k-n})      -- a linear convolution, done as a cyclic one of length
// M = 2^p >= 2N - 1 with three FFTs.  Only |X[k]| = |conv[k]| is wanted, so the final chirp multiply is skipped.
// Every phase is reduced exactly in integers before sincospi (m^2 mod 2N; twiddles k / Ns with Ns a power of two),
// so the accuracy does not degrade with N.  The direct kernel stays as the checker (tests/test_gpu_parity.py).
#pragma once
#include "common.cuh"

namespace tk {

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// b[m mod M] = conj(w_m) = exp(+i pi m^2 / N) for |m| < N, zero elsewhere (one series; shared by the whole batch)
__global__ void __launch_bounds__(256) bs_chirp_kernel(long long N, long long M, double2 *__restrict__ b)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    long long m = -1;
    if (j < N) m = j;
    else if (j > M - N) m = M - j;  // negative index -m, same m^2
    double2 v = make_double2(0.0, 0.0);
    if (m >= 0) {
        const long long r = (long long)(((unsigned long long)m * (unsigned long long)m) % (unsigned long long)(2 * N));
        double s, c;
        sincospi((double)r / (double)N, &s, &c);
        v = make_double2(c, s);
    }
    b[j] = v;
}

// a[batch][n] = x[n] w_n for n < N, zero up to M
__global__ void __launch_bounds__(256)
bs_load_kernel(const double *__restrict__ series, long long N, long long M, double2 *__restrict__ a)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    double2 v = make_double2(0.0, 0.0);
    if (j < N) {
        const long long r = (long long)(((unsigned long long)j * (unsigned long long)j) % (unsigned long long)(2 * N));
        double s, c;
        sincospi((double)r / (double)N, &s, &c);
        const double x = series[(size_t)blockIdx.y * N + j];
        v = make_double2(x * c, -x * s);
    }
    a[(size_t)blockIdx.y * M + j] = v;
}

// One radix-2 Stockham stage (autosort, out of place): butterflies of half-size Ns.  sign = -1 forward, +1 inverse.
// `mul` (optional, last forward stage of the data / first stage never): multiply the outputs pointwise by mul[.]
__global__ void __launch_bounds__(256)
fft_stage_kernel(const double2 *__restrict__ in, double2 *__restrict__ out, long long M, long long Ns, double sign,
                 const double2 *__restrict__ mul)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M / 2) return;
    const double2 *x = in + (size_t)blockIdx.y * M;
    double2 *y = out + (size_t)blockIdx.y * M;
    const long long k = j & (Ns - 1);
    double s, c;
    sincospi(sign * (double)k / (double)Ns, &s, &c);  // exp(sign * 2 pi i k / (2 Ns)); k / Ns is exact
    const double2 v0 = x[j], v1 = cmul(x[j + M / 2], make_double2(c, s));
    const long long j0 = ((j - k) << 1) + k;
    double2 o0 = make_double2(v0.x + v1.x, v0.y + v1.y), o1 = make_double2(v0.x - v1.x, v0.y - v1.y);
    if (mul != nullptr) {
        o0 = cmul(o0, mul[j0]);
        o1 = cmul(o1, mul[j0 + Ns]);
    }
    y[j0] = o0;
    y[j0 + Ns] = o1;
}

// spectrum[batch][k] = |conv[k]| / M   (the inverse transform is unscaled)
__global__ void __launch_bounds__(256)
bs_abs_kernel(const double2 *__restrict__ conv, long long N, long long M, double *__restrict__ spectrum)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    const double2 v = conv[(size_t)blockIdx.y * M + k];
    spectrum[(size_t)blockIdx.y * N + k] = hypot(v.x, v.y) / (double)M;
}

}  // namespace tk
