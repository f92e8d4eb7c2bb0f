/*
 * CPU oracle for the tetrakis-sim leapfrog step -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain C restatement of /root/reference/tetrakis_sim/physics.py:104-129.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Build (oracle/build_oracle.py):  gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC
 * -ffp-contract=off matters: the reference is Python float arithmetic (one IEEE-754
 * rounding per operation), so no fused multiply-add may be formed.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* physics.py:120-126, evaluated left to right exactly as Python does. */
static inline double leap(double u, double up, double cm, double lap, double cv, double damping)
{
    double t = 2.0 * u;
    t = t - up;
    t = t + cm * lap;
    t = t - cv * u;
    t = t - damping * (u - up);
    return t;
}

/*
 * Generic graph (any networkx graph): physics.py:106-129.
 *   rowptr/colidx : neighbour lists in adjacency order (physics.py:94)
 *   deg           : G.degree (physics.py:60-62) -- passed separately, self-loops count twice there
 *   u, uprev      : in = state at t0, out = state after `steps`
 *   hist          : NULL or (steps+1)*n doubles; row 0 = initial u (physics.py:104,129)
 *   probe_out     : NULL or (steps+1)*n_probes doubles
 * Sequential and bit-exact with the reference.
 */
int tko_graph_run(int64_t n, const int64_t *rowptr, const int32_t *colidx, const double *deg,
                  const double *inv_mass, const double *potential, double *u, double *uprev,
                  int64_t steps, double coeff, double damping, double *hist,
                  const int64_t *probe_idx, int64_t n_probes, double *probe_out)
{
    double *unew = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (!unew) return -1;
    if (hist) memcpy(hist, u, sizeof(double) * (size_t)n);
    for (int64_t p = 0; p < n_probes; ++p) probe_out[p] = u[probe_idx[p]];
    for (int64_t s = 0; s < steps; ++s) {
        for (int64_t i = 0; i < n; ++i) {
            double nsum = 0.0;                                   /* physics.py:115-117 */
            for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) nsum += u[colidx[k]];
            double lap = nsum - deg[i] * u[i];                   /* physics.py:119 */
            unew[i] = leap(u[i], uprev[i], coeff * inv_mass[i], lap, coeff * potential[i], damping);
        }
        memcpy(uprev, u, sizeof(double) * (size_t)n);            /* physics.py:128 */
        memcpy(u, unew, sizeof(double) * (size_t)n);
        if (hist) memcpy(hist + (size_t)(s + 1) * (size_t)n, u, sizeof(double) * (size_t)n);
        for (int64_t p = 0; p < n_probes; ++p)
            probe_out[(size_t)(s + 1) * (size_t)n_probes + (size_t)p] = u[probe_idx[p]];
    }
    free(unew);
    return 0;
}

/*
 * Rank-structured ("clique-sum") restatement for tetrakis lattices:
 *   neighbour_sum(z,r,c,q) = (CellSum - u) + (RowSum[z,r,q] - u) + (ColSum[z,c,q] - u) + u(z-1) + u(z+1)
 * with a participation mask.  Layout [z][r][c][q], q fastest.
 *   flags : NULL (all nodes alive and participating) or one byte per node: bit0 alive, bit1 part
 *           (part = alive and not edge-pruned; lattice.py cliques + defects.py:110-135,227-232)
 *   inv_mass/potential : NULL (1 and 0) or per node (physics.py:97-102)
 *   corr_* : directed corrections, node a gains sign*u[b] in its neighbour sum and sign in its
 *            degree (sign=-1: an edge the clique formula assumes but the graph lacks, e.g. the
 *            wedge, defects.py:78-81; +1: an extra edge)
 * Threads: OpenMP over (z, r) rows.  Deterministic (fixed summation order per row/column).
 */
int tko_structured_run(int32_t layers, int32_t size, const uint8_t *flags, const double *inv_mass,
                       const double *potential, int64_t n_corr, const int64_t *corr_a,
                       const int64_t *corr_b, const double *corr_sign, double *u, double *uprev,
                       int64_t steps, double coeff, double damping, const int64_t *probe_idx,
                       int64_t n_probes, double *probe_out)
{
    const int64_t S = size, L = layers;
    const int64_t plane = S * S * 4, n = L * plane;
    double *rowsum = (double *)malloc(sizeof(double) * (size_t)(L * S * 4));
    double *colsum = (double *)malloc(sizeof(double) * (size_t)(L * S * 4));
    double *rowcnt = (double *)malloc(sizeof(double) * (size_t)(L * S * 4));
    double *colcnt = (double *)malloc(sizeof(double) * (size_t)(L * S * 4));
    double *cadd = (double *)calloc((size_t)(n_corr > 0 ? n_corr : 1), sizeof(double));
    if (!rowsum || !colsum || !rowcnt || !colcnt || !cadd) return -1;
#define PART(i) (flags ? ((flags[i] >> 1) & 1) : 1)
#define ALIVE(i) (flags ? (flags[i] & 1) : 1)
    /* static participation counts */
#pragma omp parallel for schedule(static)
    for (int64_t zr = 0; zr < L * S; ++zr) {
        for (int q = 0; q < 4; ++q) {
            double cnt = 0.0;
            for (int64_t c = 0; c < S; ++c) cnt += PART((zr * S + c) * 4 + q);
            rowcnt[zr * 4 + q] = cnt;
        }
    }
#pragma omp parallel for schedule(static)
    for (int64_t zc = 0; zc < L * S; ++zc) {
        int64_t z = zc / S, c = zc % S;
        for (int q = 0; q < 4; ++q) {
            double cnt = 0.0;
            for (int64_t r = 0; r < S; ++r) cnt += PART(((z * S + r) * S + c) * 4 + q);
            colcnt[zc * 4 + q] = cnt;
        }
    }
    for (int64_t p = 0; p < n_probes; ++p) probe_out[p] = u[probe_idx[p]];
    for (int64_t s = 0; s < steps; ++s) {
#pragma omp parallel for schedule(static)
        for (int64_t zr = 0; zr < L * S; ++zr) {
            double acc[4] = {0, 0, 0, 0};
            for (int64_t c = 0; c < S; ++c)
                for (int q = 0; q < 4; ++q) {
                    int64_t i = (zr * S + c) * 4 + q;
                    if (PART(i)) acc[q] += u[i];
                }
            for (int q = 0; q < 4; ++q) rowsum[zr * 4 + q] = acc[q];
        }
        /* column sums: threads over (layer, block of 128 column entries); each walks the rows in order,
         * so every column's sum has a fixed order whatever the thread count */
        {
            const int64_t W = S * 4, nblk = (W + 127) / 128;
#pragma omp parallel for collapse(2) schedule(static)
            for (int64_t z = 0; z < L; ++z)
                for (int64_t blk = 0; blk < nblk; ++blk) {
                    const int64_t lo = blk * 128, hi = lo + 128 < W ? lo + 128 : W;
                    double *cs = colsum + z * W;
                    for (int64_t cq = lo; cq < hi; ++cq) cs[cq] = 0.0;
                    for (int64_t r = 0; r < S; ++r) {
                        const int64_t base = (z * S + r) * W;
                        for (int64_t cq = lo; cq < hi; ++cq)
                            if (PART(base + cq)) cs[cq] += u[base + cq];
                    }
                }
        }
        /* corrections read u_t, so gather them before u is overwritten (uprev <- unew in place) */
        for (int64_t k = 0; k < n_corr; ++k) cadd[k] = u[corr_b[k]];
#pragma omp parallel for schedule(static)
        for (int64_t zr = 0; zr < L * S; ++zr) {
            int64_t z = zr / S;
            for (int64_t c = 0; c < S; ++c) {
                int64_t base = (zr * S + c) * 4;
                double cell = 0.0, ccnt = 0.0;
                for (int q = 0; q < 4; ++q)
                    if (PART(base + q)) { cell += u[base + q]; ccnt += 1.0; }
                for (int q = 0; q < 4; ++q) {
                    int64_t i = base + q;
                    if (!ALIVE(i)) { uprev[i] = 0.0; continue; }
                    double ui = u[i], lap = 0.0;
                    if (PART(i)) {
                        double nsum = (cell - ui) + (rowsum[zr * 4 + q] - ui) +
                                      (colsum[(z * S + c) * 4 + q] - ui);
                        double deg = (ccnt - 1.0) + (rowcnt[zr * 4 + q] - 1.0) +
                                     (colcnt[(z * S + c) * 4 + q] - 1.0);
                        if (z > 0 && PART(i - plane)) { nsum += u[i - plane]; deg += 1.0; }
                        if (z < L - 1 && PART(i + plane)) { nsum += u[i + plane]; deg += 1.0; }
                        lap = nsum - deg * ui;
                    }
                    /* stash lap in place of the neighbour sum; finish after corrections */
                    double im = inv_mass ? inv_mass[i] : 1.0, v = potential ? potential[i] : 0.0;
                    if (n_corr == 0)
                        uprev[i] = leap(ui, uprev[i], coeff * im, lap, coeff * v, damping);
                    else {
                        /* rare path: apply corrections for this node by linear scan */
                        for (int64_t k = 0; k < n_corr; ++k)
                            if (corr_a[k] == i) lap += corr_sign[k] * (cadd[k] - ui);
                        uprev[i] = leap(ui, uprev[i], coeff * im, lap, coeff * v, damping);
                    }
                }
            }
        }
        double *t = u; u = uprev; uprev = t; /* physics.py:128 */
        for (int64_t p = 0; p < n_probes; ++p)
            probe_out[(size_t)(s + 1) * (size_t)n_probes + (size_t)p] = u[probe_idx[p]];
    }
    /* an odd number of steps leaves the caller's buffers swapped: report it */
    free(rowsum); free(colsum); free(rowcnt); free(colcnt); free(cadd);
    (void)n;
    return (int)(steps & 1);
}
