// Batched ensembles: one independent small-lattice simulation per CTA, state resident in shared
// memory for all its steps (no HBM state traffic), probe DFT fused into the epilogue.
//
// Replaces one Python process per parameter point (reference examples/naked_singularity_sweep.py:
// 211-212) and the serial loop of tetrakis_sim/evals/dataset.py:81-163; each simulation is the same
// leapfrog update as reference tetrakis_sim/physics.py:104-129 on a lattice.py sheet with a
// defects.py mask, evaluated with the clique-sum form (see kernels_lattice.cuh).
#pragma once
#include "common.cuh"

namespace tk {

struct EnsDev {
    int S, L;          // cells per side, layers
    int n_masks;
    int steps;
    const unsigned char *cell_bits;  // [n_masks][L*S*S] bits 0-3 part(q), 4-7 alive(q)
    const double *attr_im;           // [n_attr_rows][N] dense inv_mass rows for masks that carry tags
    const double *attr_v;            // [n_attr_rows][N] potential
    const int *mask_attr_row;        // [n_masks] row of attr_* or -1
    const int *corr_off;             // [n_masks+1]
    const int *corr_a, *corr_b;      // node indices
    const double *corr_sign;
    const int *sim_mask, *sim_kick;  // [n_sims]
    const double *sim_coeff, *sim_damping, *sim_kick_value;
    double *series;    // [n_sims][steps+1]
    double *spectrum;  // [n_sims][steps+1] or nullptr
    long long *argmax; // [n_sims] or nullptr
    double *features;  // [n_sims][kEnsNFeat] or nullptr (needs spectrum): see ens_features
};

// Spectral + time-series features of one simulation's probe series, in the epilogue of the ensemble kernel
// (reference tetrakis_sim/evals/features.py:11-81, evaluated on the non-negative-frequency half of the spectrum):
//   0 dominant bin   1 centroid [bins]   2 bandwidth [bins]   3 peak amplitude   4 low/high ratio
//   5-9 bins of the five largest peaks (-1 = none)   10-14 their magnitude / (max + 1e-12)
//   15 rms of the series   16 max |series|   17 sum of the half spectrum
// Frequencies are in BINS: f = bin / ((steps+1) dt_eff), which the host applies (the kernel never sees dt).
constexpr int kEnsNFeat = 18;

__device__ __forceinline__ void ens_features(const double *__restrict__ a, const double *__restrict__ v, int nt,
                                             double *__restrict__ out, int lane)
{
    const int nh = (nt + 1) / 2;  // bins with fftfreq >= 0: 0 .. ceil(nt/2) - 1   (features.py:18-21)
    const int half = nh / 2;
    double total = 0.0, ka = 0.0, low = 0.0, high = 0.0, amax = -1.0, sq = 0.0, vpk = 0.0;
    int imax = 0;
    for (int k = lane; k < nh; k += 32) {
        const double x = a[k];
        total += x;
        ka += (double)k * x;
        if (k < half) low += x; else high += x;
        if (x > amax) { amax = x; imax = k; }
    }
    for (int j = lane; j < nt; j += 32) {
        sq += v[j] * v[j];
        vpk = fmax(vpk, fabs(v[j]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        total += __shfl_xor_sync(0xffffffffu, total, o);
        ka += __shfl_xor_sync(0xffffffffu, ka, o);
        low += __shfl_xor_sync(0xffffffffu, low, o);
        high += __shfl_xor_sync(0xffffffffu, high, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        vpk = fmax(vpk, __shfl_xor_sync(0xffffffffu, vpk, o));
        const double oa = __shfl_xor_sync(0xffffffffu, amax, o);
        const int oi = __shfl_xor_sync(0xffffffffu, imax, o);
        if (oa > amax || (oa == amax && oi < imax)) { amax = oa; imax = oi; }  // np.argmax: the first maximum
    }
    double feat[kEnsNFeat];
#pragma unroll
    for (int i = 0; i < kEnsNFeat; ++i) feat[i] = 0.0;
#pragma unroll
    for (int i = 5; i < 10; ++i) feat[i] = -1.0;
    feat[15] = sqrt(sq / (double)nt);
    feat[16] = vpk;
    feat[17] = total;
    if (nh > 0 && total > 0.0) {
        const double kc = ka / total;
        double bw = 0.0;
        for (int k = lane; k < nh; k += 32) bw += ((double)k - kc) * ((double)k - kc) * a[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bw += __shfl_xor_sync(0xffffffffu, bw, o);
        feat[0] = (double)imax;
        feat[1] = kc;
        feat[2] = sqrt(bw / total);
        feat[3] = amax;
        feat[4] = low / (high + 1e-12);
        // five largest peaks, descending (ties: lower bin first)
        int taken[5] = {-1, -1, -1, -1, -1};
        const int kmax = nh < 5 ? nh : 5;
        for (int j = 0; j < kmax; ++j) {
            double best = -1.0;
            int bi = 0x7fffffff;
            for (int k = lane; k < nh; k += 32) {
                bool used = false;
#pragma unroll
                for (int t = 0; t < 5; ++t) used |= taken[t] == k;
                if (!used && (a[k] > best || (a[k] == best && k < bi))) { best = a[k]; bi = k; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            taken[j] = bi;
            feat[5 + j] = (double)bi;
            feat[10 + j] = best / (amax + 1e-12);
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kEnsNFeat; ++i) out[i] = feat[i];
    }
}

constexpr int kEnsMaxCorr = 32;

// Line pair i of the row/column-sum phase: i < nline/2 rows, else columns; pair = (z, x, p) -> lines (z, x, 2p),
// (z, x, 2p+1).  off/stride address the pair's first cell and the walk along the line (in doubles), dst is the
// index in rs (columns: nline + index, cs is laid out right behind rs), m0/m1 the participation masks.
__device__ __forceinline__ void line_pair(int i, int S, int nline, const unsigned long long *lmask, int &off,
                                          int &stride, int &dst, unsigned long long &m0, unsigned long long &m1)
{
    const bool is_col = i >= nline / 2;
    const int lp = is_col ? i - nline / 2 : i;
    const int pr = lp & 1, x = (lp >> 1) % S, z = (lp >> 1) / S;
    const int l0 = ((z * S + x) * 4) + 2 * pr;
    off = (is_col ? (z * S * S + x) * 4 : (z * S + x) * S * 4) + 2 * pr;
    stride = is_col ? S * 4 : 4;
    dst = (is_col ? nline : 0) + l0;
    m0 = lmask[(is_col ? nline : 0) + l0];
    m1 = lmask[(is_col ? nline : 0) + l0 + 1];
}

// Shared-memory layout (per CTA = per simulation in flight):
//   ua, ub        [N]        the two state buffers (u_{t+1} overwrites u_{t-1})
//   rs, cs        [nline]    masked row / column sums of u_t
//   series, twc, tws [nt]    probe series, DFT twiddles
//   lmask         [2*nline]  64-bit participation mask of every row / column line (S <= 64)
//   meta          [ncell]    per cell: {degree of the 4 nodes (u8 x 4), row line | column line << 16,
//                            part|alive|part(z-1)|part(z+1) bits}: no index arithmetic in the step loop
__global__ void __launch_bounds__(448, 2)
ensemble_kernel(const __grid_constant__ EnsDev P, long long n_sims)
{
    extern __shared__ __align__(16) unsigned char tk_smem[];
    const int S = P.S, L = P.L, ncell = L * S * S, N = ncell * 4, nline = L * S * 4;
    const int nt = P.steps + 1;
    double *ua = reinterpret_cast<double *>(tk_smem);
    double *ub = ua + N;
    double *rs = ub + N;
    double *cs = rs + nline;
    double *series = cs + nline;
    double *twc = series + nt;
    double *tws = twc + nt;
    double *csig = tws + nt;                                                       // [kEnsMaxCorr]
    unsigned long long *lmask = reinterpret_cast<unsigned long long *>(csig + kEnsMaxCorr);  // [2*nline]
    // (byte offsets from the start of shared memory: pointer arithmetic that keeps the shared address space)
    const size_t meta_off = ((size_t)(reinterpret_cast<unsigned char *>(lmask + 2 * nline) - tk_smem) + 15) & ~(size_t)15;
    uint4 *meta = reinterpret_cast<uint4 *>(tk_smem + meta_off);                   // [ncell], 16 B aligned
    double *tabA = reinterpret_cast<double *>(meta + ncell);                       // [256]: u coefficient by degree
    int *ca = reinterpret_cast<int *>(tabA + 256);                                 // [kEnsMaxCorr]
    int *cb = ca + kEnsMaxCorr;
    unsigned char *bits = reinterpret_cast<unsigned char *>(cb + kEnsMaxCorr);      // [ncell]

    const int tid = threadIdx.x, nth = blockDim.x;
    for (int j = tid; j < nt; j += nth) {
        double s, c;
        sincospi(2.0 * (double)j / (double)nt, &s, &c);
        twc[j] = c;
        tws[j] = s;
    }

    for (long long sim = blockIdx.x; sim < n_sims; sim += gridDim.x) {
        const int mask = P.sim_mask[sim];
        const int kick = P.sim_kick[sim];
        const double coeff = P.sim_coeff[sim], damping = P.sim_damping[sim];
        const unsigned char *gbits = P.cell_bits + (size_t)mask * ncell;
        const int arow = P.mask_attr_row[mask];
        const bool has_attr = arow >= 0;
        const double *gim = P.attr_im + (size_t)(has_attr ? arow : 0) * N;
        const double *gv = P.attr_v + (size_t)(has_attr ? arow : 0) * N;
        const int c0 = P.corr_off[mask];
        const int ncorr = min(P.corr_off[mask + 1] - c0, kEnsMaxCorr);

        __syncthreads();  // previous simulation fully drained
        for (int i = tid; i < N; i += nth) { ua[i] = 0.0; ub[i] = 0.0; }
        for (int i = tid; i < ncell; i += nth) bits[i] = gbits[i];
        for (int i = tid; i < nt; i += nth) series[i] = 0.0;
        for (int i = tid; i < ncorr; i += nth) {
            ca[i] = P.corr_a[c0 + i];
            cb[i] = P.corr_b[c0 + i];
            csig[i] = P.corr_sign[c0 + i];
        }
        __syncthreads();
        if (tid == 0) {
            ua[kick] = P.sim_kick_value[sim];
            series[0] = ua[kick];
        }
        // participation mask of every row / column line (bit k = k-th cell along the line)
        for (int i = tid; i < 2 * nline; i += nth) {
            const bool is_col = i >= nline;
            const int l = is_col ? i - nline : i;
            const int q = l & 3, x = (l >> 2) % S, z = (l >> 2) / S;
            unsigned long long m = 0ull;
            for (int k = 0; k < S; ++k) {
                const int cell = is_col ? (z * S + k) * S + x : (z * S + x) * S + k;
                m |= (unsigned long long)((bits[cell] >> q) & 1) << k;
            }
            lmask[i] = m;
        }
        __syncthreads();
        // per-cell metadata: degrees (physics.py:60-62 on the clique model), line ids, flag bits
        for (int cell = tid; cell < ncell; cell += nth) {
            const int c = cell % S, r = (cell / S) % S, z = cell / (S * S);
            const unsigned f = bits[cell];
            const unsigned fm = z > 0 ? bits[cell - S * S] : 0u;
            const unsigned fp = z < L - 1 ? bits[cell + S * S] : 0u;
            const int rl = (z * S + r) * 4, cl = (z * S + c) * 4;
            const int ccnt = __popc(f & 0xFu);
            unsigned deg4 = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int dg = 0;
                if ((f >> q) & 1u)
                    dg = (ccnt - 1) + (__popcll(lmask[rl + q]) - 1) + (__popcll(lmask[nline + cl + q]) - 1) +
                         (int)((fm >> q) & 1u) + (int)((fp >> q) & 1u);
                for (int k = 0; k < ncorr; ++k)
                    if (ca[k] == cell * 4 + q) dg += (int)csig[k];
                deg4 |= (unsigned)(dg & 0xFF) << (8 * q);
            }
            // bit h of .w: half h (quadrants 2h, 2h+1) is ORDINARY -- both nodes alive and participating, unit
            // mass, no potential, no edge correction -- and takes the short update path
            unsigned fast = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                bool ok = ((f >> (2 * h)) & 3u) == 3u && ((f >> (4 + 2 * h)) & 3u) == 3u;
                for (int j = 0; j < 2 && ok; ++j) {
                    const int node = cell * 4 + 2 * h + j;
                    if (has_attr && (gim[node] != 1.0 || gv[node] != 0.0)) ok = false;
                    for (int k = 0; k < ncorr; ++k)
                        if (ca[k] == node) ok = false;
                }
                fast |= (unsigned)ok << h;
            }
            meta[cell] = make_uint4(deg4, (unsigned)rl | ((unsigned)cl << 16),
                                    (f & 0xFFu) | ((fm & 0xFu) << 8) | ((fp & 0xFu) << 12), fast);
        }
        // ordinary node of degree d:  u+ = tabA[d] u + (damping - 1) u- + coeff (cell + row + col + z-1 + z+1)
        for (int d = tid; d < 256; d += nth) tabA[d] = (2.0 - damping) - coeff * (3.0 + (double)d);
        const int plane = S * S * 4;
        const double cbm = damping - 1.0;
        for (int t = 0; t < P.steps; ++t) {
            double *const u = (t & 1) ? ub : ua;  // u_t
            double *const w = (t & 1) ? ua : ub;  // u_{t-1}, overwritten by u_{t+1}
            __syncthreads();
            // phase A: masked row / column sums of u_t; one thread per line PAIR (quadrants 2p, 2p+1 of a row or
            // column sit in one 16-byte slot of every cell: LDS.128)
            for (int i = tid; i < nline; i += nth) {
                int off, stride, dst;
                unsigned long long m0, m1;
                line_pair(i, S, nline, lmask, off, stride, dst, m0, m1);
                const double *p = u + off;
                double a0 = 0.0, a1 = 0.0;
                for (int k = 0; k < S; ++k) {
                    const double2 v = *reinterpret_cast<const double2 *>(p + k * stride);
                    if (m0 & 1ull) a0 += v.x;
                    if (m1 & 1ull) a1 += v.y;
                    m0 >>= 1;
                    m1 >>= 1;
                }
                *reinterpret_cast<double2 *>(rs + dst) = make_double2(a0, a1);  // cs follows rs: dst >= nline is a column
            }
            __syncthreads();
            // phase B: update; u_{t+1} overwrites u_{t-1}.  One lane per HALF cell (2 quadrants, one
            // 16-byte LDS.128 per array): the 8 lanes of a quarter warp then read 128 contiguous bytes,
            // which is bank-conflict free -- a whole 32-byte cell per lane makes every access 2-way
            // conflicted (ncu: 41 % of the shared wavefronts were replays).  The cell sum needs the
            // partner lane's pair: one shuffle.
            for (int hc = tid; hc < 2 * ncell; hc += nth) {
                const int cell = hc >> 1, h = hc & 1;
                const uint4 md = meta[cell];
                const unsigned f = md.z;
                const int rl = (int)(md.y & 0xFFFFu), cl = (int)(md.y >> 16);
                const int n0 = cell * 4 + 2 * h;  // first of this lane's two nodes
                const double2 a = *reinterpret_cast<const double2 *>(u + n0);
                const double2 wa = *reinterpret_cast<const double2 *>(w + n0);
                const double2 r0 = *reinterpret_cast<const double2 *>(rs + rl + 2 * h);
                const double2 k0 = *reinterpret_cast<const double2 *>(cs + cl + 2 * h);
                double2 m0 = make_double2(0.0, 0.0), p0 = m0;
                if (f & 0xF00u) m0 = *reinterpret_cast<const double2 *>(u + n0 - plane);
                if (f & 0xF000u) p0 = *reinterpret_cast<const double2 *>(u + n0 + plane);
                const unsigned fh = f >> (2 * h);  // bit 0 / 1: part of this lane's nodes; 8,9: z-1; 12,13: z+1
                const double mine = ((fh & 1u) ? a.x : 0.0) + ((fh & 2u) ? a.y : 0.0);
                const double other = __shfl_xor_sync(__activemask(), mine, 1);
                const double cellsum = h ? other + mine : mine + other;  // same order in both lanes
                double un[2];
                if ((md.w >> h) & 1u) {
                    // ordinary half cell: no masks on the node itself, coefficients from the degree table
                    const unsigned dg = md.x >> (16 * h);
                    const double s0 = (((cellsum + r0.x) + k0.x) + ((fh & 0x100u) ? m0.x : 0.0)) + ((fh & 0x1000u) ? p0.x : 0.0);
                    const double s1 = (((cellsum + r0.y) + k0.y) + ((fh & 0x200u) ? m0.y : 0.0)) + ((fh & 0x2000u) ? p0.y : 0.0);
                    un[0] = fma(tabA[dg & 0xFFu], a.x, fma(cbm, wa.x, coeff * s0));
                    un[1] = fma(tabA[(dg >> 8) & 0xFFu], a.y, fma(cbm, wa.y, coeff * s1));
                    if ((unsigned)(kick - n0) < 2u) series[t + 1] = kick == n0 ? un[0] : un[1];
                } else {
                    const double uq[2] = {a.x, a.y}, up[2] = {wa.x, wa.y};
                    const double RS[2] = {r0.x, r0.y}, CS[2] = {k0.x, k0.y};
                    const double zm[2] = {m0.x, m0.y}, zp[2] = {p0.x, p0.y};
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int q = 2 * h + j, node = n0 + j;
                        const bool part = (f >> q) & 1u, alive = (f >> (4 + q)) & 1u;
                        double lap = 0.0;
                        if (part) {
                            double nsum = (cellsum - uq[j]) + (RS[j] - uq[j]) + (CS[j] - uq[j]);
                            if ((f >> (8 + q)) & 1u) nsum += zm[j];
                            if ((f >> (12 + q)) & 1u) nsum += zp[j];
                            lap = nsum;
                        }
                        for (int k = 0; k < ncorr; ++k)
                            if (ca[k] == node) lap += csig[k] * u[cb[k]];
                        lap -= (double)((md.x >> (8 * q)) & 0xFFu) * uq[j];  // degree incl. corrections
                        double im = 1.0, v = 0.0;
                        if (has_attr) { im = gim[node]; v = gv[node]; }
                        const double rnew = 2.0 * uq[j] - up[j] + (coeff * im) * lap - (coeff * v) * uq[j] -
                                            damping * (uq[j] - up[j]);
                        un[j] = alive ? rnew : 0.0;
                        if (node == kick) series[t + 1] = un[j];
                    }
                }
                *reinterpret_cast<double2 *>(w + n0) = make_double2(un[0], un[1]);
            }
        }
        __syncthreads();
        // epilogue: probe series out, |DFT| (reference physics.py:162-163) and argmax (batch.py:205)
        for (int j = tid; j < nt; j += nth) P.series[(size_t)sim * nt + j] = series[j];
        if (P.spectrum) {
            double *spec = rs;  // row sums are dead now; reuse when large enough
            const bool fits = nt <= 2 * nline;
            for (int k = tid; k < nt; k += nth) {
                double re = 0.0, imag = 0.0;
                int m = 0;  // (k*j) mod nt, updated incrementally
                for (int j = 0; j < nt; ++j) {
                    re += series[j] * twc[m];
                    imag -= series[j] * tws[m];
                    m += k;
                    if (m >= nt) m -= nt;
                }
                const double mag = hypot(re, imag);
                P.spectrum[(size_t)sim * nt + k] = mag;
                if (fits) spec[k] = mag;
            }
            if (P.features) {  // one warp: spectral + time-series features of this simulation (fits: see host check)
                __syncthreads();
                if (tid < 32) ens_features(spec, series, nt, P.features + (size_t)sim * kEnsNFeat, tid);
            }
            if (P.argmax) {
                __syncthreads();
                if (tid == 0) {
                    double best = -1.0;
                    int bi = 0;
                    for (int k = 0; k < nt; ++k) {
                        const double mag = fits ? spec[k] : P.spectrum[(size_t)sim * nt + k];
                        if (mag > best) { best = mag; bi = k; }
                    }
                    P.argmax[sim] = bi;
                }
            }
        }
    }
}

inline size_t ensemble_smem_bytes(int S, int L, int steps)
{
    const size_t ncell = (size_t)L * S * S, N = ncell * 4, nline = (size_t)L * S * 4, nt = steps + 1;
    size_t b = (2 * N + 2 * nline + 3 * nt + kEnsMaxCorr) * sizeof(double);  // ua ub rs cs series twc tws csig
    b += 2 * nline * sizeof(unsigned long long);                             // lmask
    b += ncell * sizeof(uint4) + 256 * sizeof(double);                       // meta, tabA
    b += 2 * kEnsMaxCorr * sizeof(int) + ncell + 16;                         // ca cb bits + alignment slack
    return (b + 15) & ~(size_t)15;
}

}  // namespace tk
