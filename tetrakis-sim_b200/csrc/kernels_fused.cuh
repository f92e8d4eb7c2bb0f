// K leapfrog steps in ONE pass over the state, for defect-free 2-D sheets.
//
// On a sheet without defects every row/column clique is complete and all nodes share one degree
// (2S+1), mass and potential, so the clique sums obey their own closed recurrences
// (reference operator: tetrakis_sim/lattice.py:30-73, update: tetrakis_sim/physics.py:119-126):
//
//   lap(r,c,q)  = cell(r,c) + RS[r,q] + CS[c,q] - (2S+4) u
//   RS+[r,q]    = 2 RS - RS- + coeff (RT[r] + Tot[q] - (S+4) RS[r,q]) - g (RS - RS-),  RT[r]  = sum_q RS[r,q]
//   CS+[c,q]    = 2 CS - CS- + coeff (CT[c] + Tot[q] - (S+4) CS[c,q]) - g (CS - CS-),  CT[c]  = sum_q CS[c,q]
//   Tot+[q]     = 2 Tot - Tot- + coeff (TT - 4 Tot[q]) - g (Tot - Tot-),               TT     = sum_q Tot[q]
//
// i.e. the sums of steps t+1 .. t+K-1 are known without touching the state.  Given them, a node's
// update only needs its own cell:  u+ = ca u + cb u- + coeff cell + coeff RS[r] + coeff CS[c],
// ca = (2-g) - coeff (2S+4), cb = -(1-g).  The update is linear, so the pass splits it:
//   * every lane steps its own cell K times WITHOUT the line terms (registers only), and
//   * the contribution of the line terms -- the response y of a cell at rest to the forcing sequence
//     coeff RS_k[r] (resp. coeff CS_k[c]), stepped with the same cell recurrence -- is computed once
//     per row / column line by lat_tables_kernel and added at the end of the pass.
// The lane steps its cell in the cell's eigenbasis (mean + three deviations: four independent scalar
// recurrences, 4 FP64 instructions per cell and step when damping = 0, 8 otherwise).
// A pass reads u_t, u_{t-1} and writes u_{t+K}, u_{t+K-1} in place: 32/K bytes per node update
// instead of 24.  It also produces the exact partial row/column sums of both written states, so the
// next pass restarts every recurrence from sums of the real data.
#pragma once
#include "common.cuh"

namespace tk {

// Probes of a fused run: every CTA filters them down to the ones inside its own tile once (shared memory), so the
// per-row check costs nothing where there is no probe and the list can be long.
constexpr int kFusedMaxProbes = 64;
struct FusedProbes {
    int n;
    int q[kFusedMaxProbes];
    long long cell[kFusedMaxProbes];  // r*S + c of the probe's cell
};

// Row/column sums of a sheet that is zero except for n <= 64 distinct kicks (tk_plan_set_state_sparse):
// one thread per (line, quadrant) adds the kicks of its line in kick order.  idx = node index (r*S+c)*4+q.
// info_prev (optional) gets zeros: the previous level of a freshly kicked state is empty.
__global__ void __launch_bounds__(256)
lat_sparse_sums_kernel(int S, int n, const long long *__restrict__ idx, const double *__restrict__ val,
                       int state_is_f32, double *__restrict__ rowinfo, double *__restrict__ colinfo)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * S * 4) return;
    const int q = t & 3, line = (t >> 2) % S;
    const bool is_col = (t >> 2) >= S;
    double a = 0.0;
    for (int i = 0; i < n; ++i) {
        const long long cell = idx[i] >> 2;
        const int kq = (int)(idx[i] & 3), r = (int)(cell / S), c = (int)(cell % S);
        if (kq == q && (is_col ? c : r) == line) a += state_is_f32 ? (double)(float)val[i] : val[i];  // as stored
    }
    (is_col ? colinfo : rowinfo)[(size_t)line * 8 + q] = a;
}

__global__ void __launch_bounds__(256) lat_zero_sums_kernel(int S, double *__restrict__ rowinfo, double *__restrict__ colinfo)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * S * 4) return;
    ((t >> 2) >= S ? colinfo : rowinfo)[(size_t)((t >> 2) % S) * 8 + (t & 3)] = 0.0;
}

// Per-block quadrant totals (8 consecutive rows per block, fixed order) of both time levels from the
// finalised row sums: totpart[lvl][blk][q].  Only needed when a run enters the fused mode; afterwards
// lat_finalise2_kernel produces them as a by-product.
__global__ void __launch_bounds__(256)
lat_totparts_kernel(int nrows, int nblk, const double *__restrict__ rowinfo, const double *__restrict__ rowinfo_prev,
                    double *__restrict__ totpart)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * nblk * 4) return;
    const int q = t & 3, blk = (t >> 2) % nblk, lvl = (t >> 2) / nblk;
    const double *src = lvl ? rowinfo_prev : rowinfo;
    double a = 0.0;
    for (int r = blk * 8; r < min(nrows, blk * 8 + 8); ++r) a += src[(size_t)r * 8 + q];
    totpart[((size_t)lvl * nblk + blk) * 4 + q] = a;
}

// Finalise BOTH states a pass has written, in one launch: fixed-order sums of the row partials (one warp per
// row line) and of the column partials (8 chunk slices per CTA), as lat_finalise_kernel does for one state,
// plus the per-block quadrant totals the next lat_tables_kernel starts from.
// grid = 2 * (nrowblk + ncolblk); first half = state a (u_{t+K}), second half = state b (u_{t+K-1}).
__global__ void __launch_bounds__(256)
lat_finalise2_kernel(int S, int nrows, int nsegk, int nchunk, int nrowblk, int ncolblk, const double *__restrict__ prow_a,
                     const double *__restrict__ pcol_a, const double *__restrict__ prow_b,
                     const double *__restrict__ pcol_b, double *__restrict__ rowinfo_a, double *__restrict__ colinfo_a,
                     double *__restrict__ rowinfo_b, double *__restrict__ colinfo_b, double *__restrict__ totpart)
{
    __shared__ double part[8][32];
    const int half = nrowblk + ncolblk;
    const int lvl = (int)blockIdx.x >= half ? 1 : 0;
    const int blk = (int)blockIdx.x - lvl * half;
    const double *prow = lvl ? prow_b : prow_a, *pcol = lvl ? pcol_b : pcol_a;
    double *rowinfo = lvl ? rowinfo_b : rowinfo_a, *colinfo = lvl ? colinfo_b : colinfo_a;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (blk < nrowblk) {
        const int line = blk * 8 + w;
        double acc = 0.0;
        if (line < nrows) {
            const double *p = prow + (size_t)line * nsegk * 4;
            const int len = nsegk * 4;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int i = lane;
            for (; i + 96 < len; i += 128) {
                a0 += p[i];
                a1 += p[i + 32];
                a2 += p[i + 64];
                a3 += p[i + 96];
            }
            for (; i < len; i += 32) a0 += p[i];
            acc = (a0 + a1) + (a2 + a3);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            acc += __shfl_xor_sync(0xffffffffu, acc, 8);
            acc += __shfl_xor_sync(0xffffffffu, acc, 16);
            if (lane < 4) rowinfo[(size_t)line * 8 + lane] = acc;
        }
        if (lane < 4) part[w][lane] = acc;
        __syncthreads();
        if (threadIdx.x < 4) {
            const int q = threadIdx.x;
            totpart[((size_t)lvl * nrowblk + blk) * 4 + q] =
                ((part[0][q] + part[1][q]) + (part[2][q] + part[3][q])) + ((part[4][q] + part[5][q]) + (part[6][q] + part[7][q]));
        }
    } else {
        const long long t = (long long)(blk - nrowblk) * 32 + lane;
        double acc = 0.0;
        if (t < (long long)S * 4) {
            const double *p = pcol + t;
            const size_t stride = (size_t)S * 4;
            double a0 = 0.0, a1 = 0.0;
            int k = w;
            for (; k + 8 < nchunk; k += 16) {
                a0 += p[(size_t)k * stride];
                a1 += p[(size_t)(k + 8) * stride];
            }
            if (k < nchunk) a0 += p[(size_t)k * stride];
            acc = a0 + a1;
        }
        part[w][lane] = acc;
        __syncthreads();
        if (w == 0 && t < (long long)S * 4) {
            const double s01 = part[0][lane] + part[1][lane], s23 = part[2][lane] + part[3][lane];
            const double s45 = part[4][lane] + part[5][lane], s67 = part[6][lane] + part[7][lane];
            colinfo[(t >> 2) * 8 + (t & 3)] = (s01 + s23) + (s45 + s67);
        }
    }
}

// Row-band plans (one sheet sharded over GPUs by rows): the column sums and quadrant totals of a band are
// partial; they travel in one exchange buffer that the host all-reduces between passes.
//   ex[0 .. 4S)      column sums of the current state      ex[4S .. 8S)      of the previous state
//   ex[8S .. 8S+4)   quadrant totals of the current state  ex[8S+4 .. 8S+8)  of the previous state
__global__ void __launch_bounds__(256)
sheet_pack_kernel(int S, int nblk, const double *__restrict__ colinfo, const double *__restrict__ colinfo_prev,
                  const double *__restrict__ totpart, double *__restrict__ ex)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 8 * S) {
        const int lvl = t >= 4 * S, j = t - lvl * 4 * S;
        ex[t] = (lvl ? colinfo_prev : colinfo)[(size_t)(j >> 2) * 8 + (j & 3)];
    } else if (t < 8 * S + 8) {
        const int j = t - 8 * S, lvl = j >> 2, q = j & 3;
        double a = 0.0;
        for (int blk = 0; blk < nblk; ++blk) a += totpart[((size_t)lvl * nblk + blk) * 4 + q];
        ex[t] = a;
    }
}

__global__ void __launch_bounds__(256)
sheet_unpack_kernel(int S, const double *__restrict__ ex, double *__restrict__ colinfo,
                    double *__restrict__ colinfo_prev, double *__restrict__ tot_global)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 8 * S) {
        const int lvl = t >= 4 * S, j = t - lvl * 4 * S;
        (lvl ? colinfo_prev : colinfo)[(size_t)(j >> 2) * 8 + (j & 3)] = ex[t];
    } else if (t < 8 * S + 8) {
        tot_global[t - 8 * S] = ex[t];
    }
}

// One thread per row or column line.  Advances the line's four quadrant sums X_k (k = 0..K-1) with the
// recurrences above and, alongside, the response y of a cell at rest to the forcing coeff * X_k:
//   y_{k+1} = ca y_k + cb y_{k-1} + coeff sum_q y_k + coeff X_k,   y_0 = y_{-1} = 0.
// resp[k][line][q] = y_{k+1}[q]; the pass adds y_K to u_{t+K} and y_{K-1} to u_{t+K-1} (and y_{k+1}
// to a probe's value at step t+k+1).
__global__ void __launch_bounds__(128)
lat_tables_kernel(int S, int nrows, int K, double coeff, double damping, double ca, double cb,
                  const double *__restrict__ rowinfo, const double *__restrict__ rowinfo_prev,
                  const double *__restrict__ colinfo, const double *__restrict__ colinfo_prev,
                  const double *__restrict__ totpart, int nblk, const double *__restrict__ tot_global,
                  double *__restrict__ respR, double *__restrict__ respC)
{
    // quadrant totals of both time levels: every CTA sums the per-block totals in the same fixed order
    // (row-band plans of a sheet sharded over GPUs pass the all-reduced totals in tot_global instead)
    __shared__ double red[2][128];
    if (tot_global != nullptr) {
        if (threadIdx.x < 4) {
            red[0][threadIdx.x] = tot_global[threadIdx.x];
            red[1][threadIdx.x] = tot_global[4 + threadIdx.x];
        }
        __syncthreads();
    } else {
        const int t = threadIdx.x, q = t & 3;
        double a = 0.0, b = 0.0;
        for (int blk = t >> 2; blk < nblk; blk += 32) {
            a += totpart[(size_t)blk * 4 + q];
            b += totpart[((size_t)nblk + blk) * 4 + q];
        }
        red[0][t] = a;
        red[1][t] = b;
        __syncthreads();
        for (int o = 64; o >= 4; o >>= 1) {
            if (t < o) {
                red[0][t] += red[0][t + o];
                red[1][t] += red[1][t + o];
            }
            __syncthreads();
        }
    }
    const int line = blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= nrows + S) return;
    const bool is_col = line >= nrows;
    const int idx = is_col ? line - nrows : line;
    const double *src = (is_col ? colinfo : rowinfo) + (size_t)idx * 8;
    const double *srcp = (is_col ? colinfo_prev : rowinfo_prev) + (size_t)idx * 8;
    // layout [k][line][q]: the threads of a warp (consecutive lines) store contiguously
    const size_t nlines = is_col ? (size_t)S : (size_t)nrows;
    double *dst = (is_col ? respC : respR) + (size_t)idx * 4;
    double X[4], Xp[4], T[4], Tp[4], y[4] = {0, 0, 0, 0}, yp[4] = {0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        X[q] = src[q];
        Xp[q] = srcp[q];
        T[q] = red[0][q];
        Tp[q] = red[1][q];
    }
    const double s4 = (double)S + 4.0;
    for (int k = 0; k < K; ++k) {
        const double yc = coeff * ((y[0] + y[1]) + (y[2] + y[3]));
        const double XT = (X[0] + X[1]) + (X[2] + X[3]);
        const double TT = (T[0] + T[1]) + (T[2] + T[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double yn = fma(ca, y[q], fma(cb, yp[q], yc)) + coeff * X[q];
            yp[q] = y[q];
            y[q] = yn;
            dst[(size_t)k * nlines * 4 + q] = yn;
            const double xn = 2.0 * X[q] - Xp[q] + coeff * ((XT + T[q]) - s4 * X[q]) - damping * (X[q] - Xp[q]);
            const double tn = 2.0 * T[q] - Tp[q] + coeff * (TT - 4.0 * T[q]) - damping * (T[q] - Tp[q]);
            Xp[q] = X[q];
            X[q] = xn;
            Tp[q] = T[q];
            T[q] = tn;
        }
    }
}

// The pass.  CTA = WC = blockDim/32 adjacent 32-cell segments marching down `rchunk` rows, one cell
// per lane (256-bit loads/stores), like lat_step_kernel.  Shared memory holds the two response vectors
// (y_K, y_{K-1}) of the CTA's rows and columns: (rows + columns) x 64 bytes, whatever K is.
//   ua: in u_t,     out u_{t+K}      ub: in u_{t-1}, out u_{t+K-1}
//   T = double, or float for the optional fp32 state mode (arithmetic, sums and responses stay fp64; the state is
//   rounded to fp32 once per pass, not once per step)
//   CROSS (sheets with defects, kernels_cross.cuh): cells of the cross -- rows with rowflag, columns with colflag --
//   are left alone (the cross stage has already advanced them in place) and do not enter the partial sums, which
//   then are the BULK parts of the line sums; the responses come from fx_lines_kernel as two plain [line][4] arrays
//   (respR = y_K, respC = y_K of the columns, respRb / respCb = y_{K-1}; element stride 2); ua / ub are the buffers u_t / u_{t-1} are
//   read from, oa / ob the ones u_{t+K} / u_{t+K-1} are written to (the same two, swapped when K is odd); probes are
//   cross cells, so the probe path is off.
template <int MINB, bool UNDAMPED, typename T = double, bool CROSS = false>
__global__ void __launch_bounds__(256, MINB)
lat_fused2d_kernel(int S, int nrows, int K, T *ua, T *ub, const double *__restrict__ respR,
                   const double *__restrict__ respC, double *__restrict__ prow_a, double *__restrict__ pcol_a,
                   double *__restrict__ prow_b, double *__restrict__ pcol_b, double ca, double cb, double coeff,
                   int rchunk, int nchunk, int nsegk, int nct, int pf_rows,
                   const __grid_constant__ FusedProbes probes, double *__restrict__ probe_series,
                   const long long *__restrict__ ctr, long long row0,
                   const unsigned char *__restrict__ rowflag = nullptr, const unsigned char *__restrict__ colflag = nullptr,
                   const double *__restrict__ respRb = nullptr, const double *__restrict__ respCb = nullptr,
                   T *oa = nullptr, T *ob = nullptr)
{
    extern __shared__ __align__(16) unsigned char tk_smem[];
    const int WC = blockDim.x >> 5;
    const int ncol = WC * 32;
    double *colresp = reinterpret_cast<double *>(tk_smem);  // [4 parts: a01 a23 b01 b23][ncol][2]: lanes 16 B apart
    double *rowresp = colresp + (size_t)ncol * 8;            // [row][a0..a3 b0..b3]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ct = blockIdx.x % nct;
    const int ch = blockIdx.x / nct;
    const int r_begin = ch * rchunk;
    const int r_end = min(nrows, r_begin + rchunk);
    const int col0 = ct * ncol;
    // j = 0..3: y_K[q] (added to u_{t+K}); j = 4..7: y_{K-1}[q] (added to u_{t+K-1}; zero when K = 1)
    for (int i = threadIdx.x; i < ncol * 8; i += blockDim.x) {
        const int cl_ = i >> 3, j = i & 7, q = j & 3, k = j < 4 ? K - 1 : K - 2;
        const int c = col0 + cl_;
        double v;
        if (CROSS) v = c < S ? (j < 4 ? respC : respCb)[((size_t)c * 4 + q) * 2] : 0.0;  // levels interleaved
        else v = (c < S && k >= 0) ? respC[((size_t)k * S + c) * 4 + q] : 0.0;
        colresp[((size_t)(j >> 1) * ncol + cl_) * 2 + (j & 1)] = v;
    }
    for (int i = threadIdx.x; i < (r_end - r_begin) * 8; i += blockDim.x) {
        const int rl = i >> 3, j = i & 7, q = j & 3, k = j < 4 ? K - 1 : K - 2;
        if (CROSS) rowresp[i] = (j < 4 ? respR : respRb)[((size_t)(r_begin + rl) * 4 + q) * 2];
        else rowresp[i] = k >= 0 ? respR[((size_t)k * nrows + r_begin + rl) * 4 + q] : 0.0;
    }
    // probes inside this CTA's tile (rows [r_begin, r_end) x columns [col0, col0 + ncol)): usually none
    __shared__ int sp_n;
    __shared__ int sp_idx[kFusedMaxProbes];
    if (threadIdx.x == 0) {
        int n = 0;
        if (!CROSS) {
            for (int p = 0; p < probes.n; ++p) {
                const long long pr = probes.cell[p] / S, pc = probes.cell[p] % S;
                if (pr >= r_begin && pr < r_end && pc >= col0 && pc < col0 + ncol) sp_idx[n++] = p;
            }
        }
        sp_n = n;
    }
    __syncthreads();
    if (!CROSS) {
        oa = ua;
        ob = ub;
    }

    const int segk = ct * WC + warp;
    if (segk >= nsegk) return;
    const int cl = warp * 32 + lane;
    const int c0 = col0 + cl;
    const bool valid = c0 < S && !(CROSS && colflag[min(c0, S - 1)]);
    long long base_row = row0;
    if (ctr != nullptr) base_row = ctr[0];
    const double lam_m = ca + 4.0 * coeff;
    double cacc_a[4] = {0, 0, 0, 0}, cacc_b[4] = {0, 0, 0, 0};
    for (int r = r_begin; r < r_end; ++r) {
        const size_t off = ((size_t)r * S + c0) * 4;
        double u[4] = {0, 0, 0, 0}, up[4] = {0, 0, 0, 0};
        if (CROSS && rowflag[r]) {  // a row of the cross: nothing to do, its bulk part is empty
            warp_sum4_store(u, lane, prow_a + ((size_t)r * nsegk + segk) * 4);
            warp_sum4_store(up, lane, prow_b + ((size_t)r * nsegk + segk) * 4);
            continue;
        }
        if (valid) {
            if constexpr (sizeof(T) == 8) {
                ld256_cs(reinterpret_cast<const double *>(ua) + off, u);
                ld256_cs(reinterpret_cast<const double *>(ub) + off, up);
            } else {
                const float4 fa = ld128_cs(reinterpret_cast<const float *>(ua) + off);
                const float4 fb = ld128_cs(reinterpret_cast<const float *>(ub) + off);
                u[0] = fa.x; u[1] = fa.y; u[2] = fa.z; u[3] = fa.w;
                up[0] = fb.x; up[1] = fb.y; up[2] = fb.z; up[3] = fb.w;
            }
            // one L2 prefetch request per 128-byte line: 4 lanes of 32-byte cells, 8 lanes of 16-byte cells
            if (pf_rows > 0 && r + pf_rows < r_end && (lane & (sizeof(T) == 8 ? 3 : 7)) == 0) {
                prefetch_l2(ua + off + (size_t)pf_rows * S * 4);
                prefetch_l2(ub + off + (size_t)pf_rows * S * 4);
            }
        }
        bool hit = false;
        if (!CROSS && sp_n > 0) {
#pragma unroll 1
            for (int i = 0; i < sp_n; ++i) hit |= valid && probes.cell[sp_idx[i]] == (long long)r * S + c0;
        }
        // Cell eigenbasis: the cell operator ca*I + coeff*J (J = all-ones 4x4) acts as (ca + 4 coeff) on the
        // cell mean m and as ca on the deviations d_q = u_q - m (d_3 = -(d_0+d_1+d_2)), so K steps are four
        // independent scalar recurrences y+ = lam*y + cb*y-: one FMA per recurrence and step when damping = 0
        // (cb = -1), an FMA and a multiply otherwise.
        double m = 0.25 * ((u[0] + u[1]) + (u[2] + u[3])), mp = 0.25 * ((up[0] + up[1]) + (up[2] + up[3]));
        double d0 = u[0] - m, d1 = u[1] - m, d2 = u[2] - m;
        double e0 = up[0] - mp, e1 = up[1] - mp, e2 = up[2] - mp;
#define TK_FUSED_STEP()                          \
    do {                                         \
        double mn, n0, n1, n2;                   \
        if (UNDAMPED) {                          \
            mn = fma(lam_m, m, -mp);             \
            n0 = fma(ca, d0, -e0);               \
            n1 = fma(ca, d1, -e1);               \
            n2 = fma(ca, d2, -e2);               \
        } else {                                 \
            mn = fma(lam_m, m, cb * mp);         \
            n0 = fma(ca, d0, cb * e0);           \
            n1 = fma(ca, d1, cb * e1);           \
            n2 = fma(ca, d2, cb * e2);           \
        }                                        \
        mp = m; m = mn;                          \
        e0 = d0; d0 = n0;                        \
        e1 = d1; d1 = n1;                        \
        e2 = d2; d2 = n2;                        \
    } while (0)
        if (!__any_sync(0xffffffffu, hit)) {
#pragma unroll 8
            for (int k = 0; k < K; ++k) TK_FUSED_STEP();
        } else {  // a probe's cell is in this warp's segment of the row (rare): record it after every step
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
                TK_FUSED_STEP();
                if (hit) {
#pragma unroll 1
                    for (int i = 0; i < sp_n; ++i) {
                        const int p = sp_idx[i];
                        if (probes.cell[p] == (long long)r * S + c0) {
                            const int pq = probes.q[p];
                            const double dq = pq == 0 ? d0 : (pq == 1 ? d1 : (pq == 2 ? d2 : -((d0 + d1) + d2)));
                            probe_series[(size_t)(base_row + k) * probes.n + p] =
                                (m + dq) + (respR[((size_t)k * nrows + r) * 4 + pq] + respC[((size_t)k * S + c0) * 4 + pq]);
                        }
                    }
                }
            }
        }
#undef TK_FUSED_STEP
        u[0] = m + d0; u[1] = m + d1; u[2] = m + d2; u[3] = m - ((d0 + d1) + d2);
        up[0] = mp + e0; up[1] = mp + e1; up[2] = mp + e2; up[3] = mp - ((e0 + e1) + e2);
        {
            const double *rr = rowresp + (size_t)(r - r_begin) * 8;
            const double2 ra01 = *reinterpret_cast<const double2 *>(rr);
            const double2 ra23 = *reinterpret_cast<const double2 *>(rr + 2);
            const double2 rb01 = *reinterpret_cast<const double2 *>(rr + 4);
            const double2 rb23 = *reinterpret_cast<const double2 *>(rr + 6);
            const double2 ca01 = *reinterpret_cast<const double2 *>(colresp + ((size_t)0 * ncol + cl) * 2);
            const double2 ca23 = *reinterpret_cast<const double2 *>(colresp + ((size_t)1 * ncol + cl) * 2);
            const double2 cb01 = *reinterpret_cast<const double2 *>(colresp + ((size_t)2 * ncol + cl) * 2);
            const double2 cb23 = *reinterpret_cast<const double2 *>(colresp + ((size_t)3 * ncol + cl) * 2);
            u[0] += ra01.x + ca01.x;  u[1] += ra01.y + ca01.y;  u[2] += ra23.x + ca23.x;  u[3] += ra23.y + ca23.y;
            up[0] += rb01.x + cb01.x; up[1] += rb01.y + cb01.y; up[2] += rb23.x + cb23.x; up[3] += rb23.y + cb23.y;
        }
        if (valid) {
            if constexpr (sizeof(T) == 8) {
                st256_cs(reinterpret_cast<double *>(oa) + off, u);
                st256_cs(reinterpret_cast<double *>(ob) + off, up);
            } else {
                const float4 fa = make_float4((float)u[0], (float)u[1], (float)u[2], (float)u[3]);
                const float4 fb = make_float4((float)up[0], (float)up[1], (float)up[2], (float)up[3]);
                st128_cs(reinterpret_cast<float *>(oa) + off, fa);
                st128_cs(reinterpret_cast<float *>(ob) + off, fb);
                // the sums below must be those of the stored state
                u[0] = fa.x; u[1] = fa.y; u[2] = fa.z; u[3] = fa.w;
                up[0] = fb.x; up[1] = fb.y; up[2] = fb.z; up[3] = fb.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                u[q] = 0.0;
                up[q] = 0.0;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            cacc_a[q] += u[q];
            cacc_b[q] += up[q];
        }
        warp_sum4_store(u, lane, prow_a + ((size_t)r * nsegk + segk) * 4);
        warp_sum4_store(up, lane, prow_b + ((size_t)r * nsegk + segk) * 4);
    }
    if (c0 < S) {
        st256(pcol_a + ((size_t)ch * S + c0) * 4, cacc_a);
        st256(pcol_b + ((size_t)ch * S + c0) * 4, cacc_b);
    }
}

}  // namespace tk
