// K leapfrog steps per pass on lattices WITH defects and on 3-D lattices: the "cross" scheme.
//
// Reference operator: tetrakis_sim/lattice.py:30-73 (cell / row / column cliques + vertical links) with the
// defects of tetrakis_sim/defects.py:55-107 (wedge), 110-135 (black hole), 194-247 (singularity); update:
// tetrakis_sim/physics.py:119-126.
//
// Let R_d / C_d be the rows / columns that contain a special cell in ANY layer (removed, pruned or tagged
// node, an edge correction, or a probe -- probes are made part of the cross so that their value exists at
// every step).  The "cross" = all cells with r in R_d or c in C_d, every layer; the rest is the "bulk".
// Every bulk cell is plain (degree 2S+1+nz, unit mass, no potential) and its row and column are complete
// cliques, so with RS / CS the total row / column sums its update is
//     u+ = ca_z u + cb u- + coeff (cell + RS[z,r,q] + CS[z,c,q] + u(z-1) + u(z+1)),
//     ca_z = (2-g) - coeff (2S+4+nz),  cb = -(1-g).
// A pass of K steps has two stages:
//   1. the CROSS STAGE advances, one step at a time, the cross cells (general update, in place in the state
//      buffers: fx_cross_step_kernel) together with every line sum (fx_lines_kernel).  A line sum splits into the
//      part over cross cells, summed from the data, and the part B over bulk cells, which obeys a closed
//      recurrence (sum the bulk update over the bulk cells of the line):
//          B+[z,r,q] = ca_z B + cb B- + coeff (sum_q' B[q'] + n_bulk_cols RS[z,r,q] + sum_{c bulk} CS[z,c,q]
//                                              + B[z-1,r,q] + B[z+1,r,q])          (same for columns, D);
//      alongside, the response y of a bulk z-column at rest to the forcing coeff * RS_k (resp. coeff * CS_k):
//          y+[z] = ca_z y + cb y- + coeff (sum_q y + X_k[z,line,q] + y[z-1] + y[z+1]),   y_0 = y_{-1} = 0;
//   2. the BULK PASS: every bulk (r, c) z-column evolves on its own (cell coupling + z stencil) for K steps
//      without the line terms, then adds the row and column responses y_K / y_{K-1} (the update is linear):
//      2-D: lat_fused2d_kernel<CROSS> (kernels_fused.cuh); 3-D: fx_bulk3d_kernel below, which streams along z with
//      a K-level register pipeline (no halo, no redundant work, in place).
// Traffic per K steps: 32 B per bulk node + 24 B x K per cross node, instead of 24 B x K per node.
// Derivation restated in NumPy against the oracle: tests/test_cpu_host.py::test_fused_pass_with_defects_algebra_*.
#pragma once
#include "common.cuh"
#include "kernels_lattice.cuh"

namespace tk {

struct FxDev {
    int S, L;
    int nR, nC;      // rows / columns of the cross
    int nbr, nbc;    // bulk rows / columns (S - nR, S - nC)
    const int *rd_list, *cd_list, *br_list;      // cross rows, cross columns, bulk rows (ascending)
    const unsigned char *rowflag, *colflag;      // [S] 1 = in the cross
    const int *rowslot, *colslot;                // [S] index into rd_list / br_list, cd_list / (unused)
    // partial sums of the participation-masked cross cells written by fx_cross_step_kernel
    int nsegR, nchR, rcR;   // row band: 32-cell segments per row, row chunks, rows per chunk
    int nsegC, nchC, rcC;   // column band: segments of cd_list, chunks of bulk rows, rows per chunk
    int wshC;               // column band: log2 of the lanes per row (W = 2^wshC >= min(nC, 32) columns per warp)
    double *prowX;  // [L][nR][nsegR][4]
    double *pcolX;  // [L][nchR][S][4]
    double *prowC;  // [L][nbr][nsegC][4]
    double *pcolC;  // [L][nchC][nC][4]
};

// ---------------------------------------------------------------------------------------------------------
// Cross stage, lines.  Per step k two pieces of work on every (kind, layer, line):
//   advance (fx_lines_advance, runs INSIDE the launch of fx_cross_step_kernel, beside the cells, because it only needs
//            the line sums X_k): the bulk parts B (rows) / D (columns) and the responses y one step further;
//   sum     (fx_lines_kernel, after the cells): finishes the sums of the cross parts, writes the total line sums
//            X_{k+1} = bulk part + cross part into rowinfo / colinfo and leaves per-block totals of X over the bulk
//            lines for the next advance; block 0 also gathers the probes from `state` (the level X describes).
struct FxLines {
    int first;        // advance: 1 = first step of a pass, y and y- are zero (whatever the buffers hold);
                      // 2 = second step, y- is zero
    double coeff, damping;
    double *rowinfo, *colinfo;        // [L][S][8] total line sums (0-3) + static counts (4-7)
    double *Bcur[2], *Bprev[2];       // [kind][L][S][8] bulk parts (stride 8, entries 0-3), two time levels
    double *ycur[2], *yprev[2];       // [kind] responses, [L][S][4][2]: the two time levels of a value are adjacent
                                      // (the pointers differ by one double), so the bulk pass fetches both at once
    const double *tot;                // [2][L][nblk][4] per-block totals over bulk lines of X (written by the sum)
    double *tot_out;
    const void *state;                // probes
    const long long *probe_idx;
    int n_probes;
    double *probe_row;                // row to write (or the base of the series when ctr != null)
    const long long *ctr;
};

// One block of 256 lines of one (kind, layer): B_{k+1}, y_{k+1} from B_k, B_{k-1}, y_k, y_{k-1}, X_k and the totals.
__device__ __forceinline__ void fx_lines_advance(const FxDev &X, const FxLines &A, int ablk)
{
    __shared__ double tot[4];
    const int S = X.S, L = X.L, nblk = (S + 127) / 128, nb256 = (S + 255) / 256;
    const int blk = ablk % nb256;
    ablk /= nb256;
    const int zl = ablk % L, kind = ablk / L;
    const int tid = threadIdx.x;
    if (tid < 4) {
        const double *tp = A.tot + ((size_t)(1 - kind) * L + zl) * nblk * 4 + tid;
        double a = 0.0;
        for (int k = 0; k < nblk; ++k) a += tp[(size_t)k * 4];
        tot[tid] = a;
    }
    __syncthreads();
    const int x = blk * 256 + tid;
    if (x >= S || (kind ? X.colflag : X.rowflag)[x]) return;
    const double *info = (kind ? A.colinfo : A.rowinfo) + ((size_t)zl * S + x) * 8;
    const size_t li = (size_t)zl * S + x;
    const double *Bc = A.Bcur[kind], *Bp = A.Bprev[kind];
    const double nz = (zl > 0 ? 1.0 : 0.0) + (zl < L - 1 ? 1.0 : 0.0);
    const double caz = (2.0 - A.damping) - A.coeff * (2.0 * S + 4.0 + nz), cb = -(1.0 - A.damping);
    const double nb = (double)(kind ? X.nbr : X.nbc);
    double bc[4], bp[4], zs[4] = {0, 0, 0, 0}, xp[4];
    ld256_nc(Bc + li * 8, bc);
    ld256_nc(Bp + li * 8, bp);
    ld256_nc(info, xp);
    if (zl > 0) {
        double t[4];
        ld256_nc(Bc + (li - S) * 8, t);
#pragma unroll
        for (int q = 0; q < 4; ++q) zs[q] += t[q];
    }
    if (zl < L - 1) {
        double t[4];
        ld256_nc(Bc + (li + S) * 8, t);
#pragma unroll
        for (int q = 0; q < 4; ++q) zs[q] += t[q];
    }
    const double sb = (bc[0] + bc[1]) + (bc[2] + bc[3]);
    double Bn[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) Bn[q] = caz * bc[q] + cb * bp[q] + A.coeff * (((sb + nb * xp[q]) + tot[q]) + zs[q]);
    st256(A.Bprev[kind] + li * 8, Bn);  // becomes "cur" after the host's swap
    // response of a bulk z-column at rest to the forcing coeff * X_k of this line
    double yc[4] = {0, 0, 0, 0}, yp[4] = {0, 0, 0, 0}, ys[4] = {0, 0, 0, 0};
    if (A.first != 1) {
        const double *Yc = A.ycur[kind], *Yp = A.yprev[kind];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            yc[q] = Yc[(li * 4 + q) * 2];
            yp[q] = A.first == 2 ? 0.0 : Yp[(li * 4 + q) * 2];
            if (zl > 0) ys[q] += Yc[((li - S) * 4 + q) * 2];
            if (zl < L - 1) ys[q] += Yc[((li + S) * 4 + q) * 2];
        }
    }
    const double sy = (yc[0] + yc[1]) + (yc[2] + yc[3]);
#pragma unroll
    for (int q = 0; q < 4; ++q)
        A.yprev[kind][(li * 4 + q) * 2] = caz * yc[q] + cb * yp[q] + A.coeff * ((sy + xp[q]) + ys[q]);
}

// ---------------------------------------------------------------------------------------------------------
// Cross stage, cells: one leapfrog step (MODE 0) or the sums only (MODE 1) of the cross cells.
// A warp task = (layer, column segment, row chunk) of the ROW BAND (rows of R_d, all columns, 32 columns per warp)
// or of the COLUMN BAND (bulk rows, columns of C_d).  One cell per lane (256-bit accesses); a warp covers W columns
// x 32/W rows per iteration -- W = 32 in the row band, the smallest power of two >= |C_d| (at most 32) in the column
// band, so a narrow C_d does not leave most lanes idle -- and marches down its chunk like lat_step_kernel: column
// sums in registers, row sums by a butterfly over the W lanes of a row.
// rowinfo / colinfo hold the TOTAL line sums of the state being read (fx_lines_kernel) + the static counts.
// The first `nadv` blocks of the grid do not step cells: they advance the bulk parts and the responses of the lines
// (fx_lines_advance), which depends on the same line sums and nothing else -- ~200 MB of streaming that hides behind
// the latency-bound cell update instead of running as a launch of its own.
template <int MODE, int MINB = 2>
__global__ void __launch_bounds__(256, MINB)
fx_cross_step_kernel(const __grid_constant__ LatDev P, const __grid_constant__ FxDev X,
                     const __grid_constant__ FxLines A, const double *__restrict__ u, double *__restrict__ w,
                     double coeff, double damping, long long ntaskR, long long ntaskC, int nadv)
{
    constexpr bool STEP = MODE == 0;
    if ((int)blockIdx.x < nadv) {
        fx_lines_advance(X, A, (int)blockIdx.x);
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long task = (long long)(blockIdx.x - nadv) * 8 + warp;
    const bool band = task >= ntaskR;  // false: row band, true: column band
    if (band) task -= ntaskR;
    if (task >= (band ? ntaskC : ntaskR)) return;  // whole warp; no block-level sync below
    const int S = P.S, L = P.Lloc;
    const int zl = (int)(task % L);
    task /= L;
    const int nseg = band ? X.nsegC : X.nsegR;
    const int seg = (int)(task % nseg), ch = (int)(task / nseg);
    const int rc = band ? X.rcC : X.rcR, nrows = band ? X.nbr : X.nR, nch = band ? X.nchC : X.nchR;
    const int ncolb = band ? X.nC : S;
    const int wsh = band ? X.wshC : 5, W = 1 << wsh, rpw = 32 >> wsh;  // lanes per row, rows per iteration
    const int *__restrict__ rlist = band ? X.br_list : X.rd_list;
    const int jl = lane & (W - 1), rs = lane >> wsh;
    const int j = seg * W + jl;
    const bool vcol = j < ncolb;
    const int c = band ? (vcol ? X.cd_list[j] : 0) : min(j, S - 1);

    const size_t plane = (size_t)S * S * 4;
    const size_t zbase = (size_t)(zl + P.halo) * plane;
    const unsigned fplain = plain_flags(P, zl);
    const bool zm_on = P.halo && (fplain & 0x100u), zp_on = P.halo && (fplain & 0x1000u);
    const double degbase = 4.0 + (zm_on ? 1.0 : 0.0) + (zp_on ? 1.0 : 0.0);

    double CS[4] = {0, 0, 0, 0}, CC[4] = {0, 0, 0, 0}, cacc[4] = {0, 0, 0, 0};
    if (STEP && vcol) {
        ld256_nc(P.colinfo + ((size_t)zl * S + c) * 8, CS);
        ld256_nc(P.colinfo + ((size_t)zl * S + c) * 8 + 4, CC);
    }
    const int i_begin = ch * rc, i_end = min(nrows, i_begin + rc);
    const int tab0 = c / kTabSeg;
    double *__restrict__ prow = band ? X.prowC : X.prowX;
    for (int ib = i_begin; ib < i_end; ib += rpw) {
        const int i = ib + rs;
        const bool valid = vcol && i < i_end;
        const int r = rlist[min(i, nrows - 1)];
        const size_t off = zbase + ((size_t)r * S + c) * 4;
        if (STEP && valid && i + rpw < i_end && (jl & 3) == 0) {  // next iteration's rows towards L2, one request per 128 B
            const size_t o2 = zbase + ((size_t)rlist[i + rpw] * S + c) * 4;
            prefetch_l2(u + o2);
            prefetch_l2(w + o2);
        }
        double uq[4] = {0, 0, 0, 0}, upq[4] = {0, 0, 0, 0}, zm[4] = {0, 0, 0, 0}, zp[4] = {0, 0, 0, 0};
        int det = -1;
        if (valid) {
            ld256_nc(u + off, uq);
            if (STEP) {
                ld256_cs(w + off, upq);
                if (zm_on) ld256_nc(u + off - plane, zm);
                if (zp_on) ld256_nc(u + off + plane, zp);
            }
            det = P.seg_flag[((size_t)zl * S + r) * P.ntab + tab0];
        }
        const bool all_plain = __all_sync(0xffffffffu, det < 0);
        double un[4], pun[4];
        if (!STEP) {
            unsigned f = fplain;
            if (det >= 0) f = P.det_flags[(size_t)det * kTabSeg + (c & (kTabSeg - 1))];
            if (!valid) f = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) pun[q] = ((f >> q) & 1u) ? uq[q] : 0.0;
        } else {
            double RS[4] = {0, 0, 0, 0}, RC[4] = {0, 0, 0, 0};
            if (valid) {
                ld256_nc(P.rowinfo + ((size_t)zl * S + r) * 8, RS);
                ld256_nc(P.rowinfo + ((size_t)zl * S + r) * 8 + 4, RC);
            }
            if (all_plain) {
                const double cell = (uq[0] + uq[1]) + (uq[2] + uq[3]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double nb = (cell + RS[q]) + (CS[q] + (zm[q] + zp[q]));
                    const double dg = (RC[q] + degbase) + CC[q];
                    const double lap = nb - dg * uq[q];
                    un[q] = 2.0 * uq[q] - upq[q] + coeff * lap - damping * (uq[q] - upq[q]);
                    pun[q] = valid ? un[q] : 0.0;
                }
            } else {
                unsigned f = fplain;
                if (det >= 0) f = P.det_flags[(size_t)det * kTabSeg + (c & (kTabSeg - 1))];
                if (!valid) f = 0;
                CellIO io;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    io.uq[q] = uq[q]; io.upq[q] = upq[q]; io.zm[q] = zm[q]; io.zp[q] = zp[q];
                    io.RS[q] = RS[q]; io.RC[q] = RC[q]; io.CS[q] = CS[q]; io.CC[q] = CC[q];
                }
                cell_update_detail(P, u, det, c & (kTabSeg - 1), (long long)off, f, coeff, damping, 0.0, io);
#pragma unroll
                for (int q = 0; q < 4; ++q) { un[q] = io.un[q]; pun[q] = io.pun[q]; }
            }
            if (valid) st256(w + off, un);  // the cross is re-read next step: keep it in L2 (no evict-first hint)
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) cacc[q] += pun[q];
        if (wsh == 5) {
            if (i < i_end) warp_sum4_store(pun, lane, prow + (((size_t)zl * nrows + i) * nseg + seg) * 4);
        } else {  // butterfly over the W lanes of each row; lane jl == 0 of a row stores its four sums
            for (int o = W >> 1; o > 0; o >>= 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) pun[q] += __shfl_xor_sync(0xffffffffu, pun[q], o);
            }
            if (jl == 0 && i < i_end) st256(prow + (((size_t)zl * nrows + i) * nseg + seg) * 4, pun);
        }
    }
    // column sums of the chunk: add the 32/W row lanes of a column (fixed tree), lanes of row 0 store
    for (int o = 16; o >= W; o >>= 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) cacc[q] += __shfl_xor_sync(0xffffffffu, cacc[q], o);
    }
    if (vcol && rs == 0) st256((band ? X.pcolC : X.pcolX) + (((size_t)zl * nch + ch) * ncolb + j) * 4, cacc);
}

// Cross stage, lines: the sum after the cells (see above).
__global__ void __launch_bounds__(128)
fx_lines_kernel(const __grid_constant__ FxDev X, const __grid_constant__ FxLines A)
{
    __shared__ double red[4][128];
    const int S = X.S, L = X.L, nblk = (S + 127) / 128;
    int b = blockIdx.x;
    const int blk = b % nblk;
    b /= nblk;
    const int zl = b % L, kind = b / L;
    const int tid = threadIdx.x;
    if (blockIdx.x == 0 && A.n_probes > 0) {
        double *row = A.probe_row;
        if (A.ctr != nullptr) row += (size_t)A.ctr[0] * A.n_probes;
        for (int t = tid; t < A.n_probes; t += blockDim.x) row[t] = static_cast<const double *>(A.state)[A.probe_idx[t]];
    }
    const int x = blk * 128 + tid;
    double Xn[4] = {0, 0, 0, 0};
    bool bulk_line = false;
    if (x < S) {
        const bool flag = (kind ? X.colflag : X.rowflag)[x] != 0;
        const int slot = (kind ? X.colslot : X.rowslot)[x];
        double *info = (kind ? A.colinfo : A.rowinfo) + ((size_t)zl * S + x) * 8;
        // cross part of the line sum
        double cr[4] = {0, 0, 0, 0};
        if (kind == 0) {
            const double *pp = flag ? X.prowX + ((size_t)zl * X.nR + slot) * X.nsegR * 4
                                    : X.prowC + ((size_t)zl * X.nbr + slot) * X.nsegC * 4;
            const int n = flag ? X.nsegR : X.nsegC;
            for (int s = 0; s < n; ++s)
#pragma unroll
                for (int q = 0; q < 4; ++q) cr[q] += pp[(size_t)s * 4 + q];
        } else {
            for (int k = 0; k < X.nchR; ++k)
#pragma unroll
                for (int q = 0; q < 4; ++q) cr[q] += X.pcolX[(((size_t)zl * X.nchR + k) * S + x) * 4 + q];
            if (flag)
                for (int k = 0; k < X.nchC; ++k)
#pragma unroll
                    for (int q = 0; q < 4; ++q) cr[q] += X.pcolC[(((size_t)zl * X.nchC + k) * X.nC + slot) * 4 + q];
        }
        if (!flag) {
            bulk_line = true;
#pragma unroll
            for (int q = 0; q < 4; ++q) Xn[q] = A.Bcur[kind][((size_t)zl * S + x) * 8 + q] + cr[q];
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) Xn[q] = cr[q];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) info[q] = Xn[q];
    }
    // totals of X over the BULK lines of this block (fixed tree)
#pragma unroll
    for (int q = 0; q < 4; ++q) red[q][tid] = bulk_line ? Xn[q] : 0.0;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (tid < o) {
#pragma unroll
            for (int q = 0; q < 4; ++q) red[q][tid] += red[q][tid + o];
        }
        __syncthreads();
    }
    if (tid < 4) A.tot_out[(((size_t)kind * L + zl) * nblk + blk) * 4 + tid] = red[tid][0];
}

// ---------------------------------------------------------------------------------------------------------
// Bulk parts of the line sums straight from a state buffer (entering the fused mode, periodic re-anchoring in
// 3-D): warp = (layer, 32-cell segment, row chunk) marching down the chunk, cross rows / columns masked out.
// Partials go to prow [L][S][nseg32][4] / pcol [L][nchunk][S][4] and are summed by lat_finalise_kernel.
__global__ void __launch_bounds__(256)
fx_bulk_sums_kernel(const __grid_constant__ FxDev X, int halo, const double *__restrict__ u, double *__restrict__ prow,
                    double *__restrict__ pcol, int rchunk, int nchunk, int nseg, long long ntask)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long task = (long long)blockIdx.x * 8 + warp;
    if (task >= ntask) return;
    const int S = X.S, L = X.L;
    const int zl = (int)(task % L);
    task /= L;
    const int seg = (int)(task % nseg), ch = (int)(task / nseg);
    const int c = seg * 32 + lane;
    const bool use = c < S && !X.colflag[min(c, S - 1)];
    const size_t zbase = (size_t)(zl + halo) * S * S * 4;
    double cacc[4] = {0, 0, 0, 0};
    const int r_end = min(S, (ch + 1) * rchunk);
    for (int r = ch * rchunk; r < r_end; ++r) {
        double v[4] = {0, 0, 0, 0};
        if (use && !X.rowflag[r]) ld256_nc(u + zbase + ((size_t)r * S + c) * 4, v);
#pragma unroll
        for (int q = 0; q < 4; ++q) cacc[q] += v[q];
        warp_sum4_store(v, lane, prow + (((size_t)zl * S + r) * nseg + seg) * 4);
    }
    if (c < S) st256(pcol + (((size_t)zl * nchunk + ch) * S + c) * 4, cacc);
}

// ---------------------------------------------------------------------------------------------------------
// 3-D bulk pass: K steps of every bulk (r, c) z-column in ONE streaming sweep along z.
//
// Thread = one Hadamard mode of one cell: lanes 4j..4j+3 of a warp hold the four quadrant values of cell j and
// turn them into the sum mode + three difference modes with two xor-shuffle rounds; the cell operator
// ca I + coeff J acts as ca + 4 coeff on the sum mode and as ca on the others, and the vertical links do not mix
// quadrants, so each mode is an independent 1-D recurrence along z:
//     A_k(z) = lam A_{k-1}(z) + cb A_{k-2}(z) + coeff (A_{k-1}(z-1) + A_{k-1}(z+1)),
// with lam taken for two vertical neighbours and the end layers mirrored (a missing neighbour contributes
// "itself": -coeff u + coeff u = 0, which is the end layer's degree).  The thread sweeps z once with a pipeline of
// K levels in registers: at tick t it loads layer t of u_t / u_{t-1} and level k advances to layer t-k, so level K
// of layer t-K and level K-1 of the same layer leave the pipeline and are stored in place -- every one of the K
// updates of every node is computed, no intermediate state is stored, there is no halo and no redundant work.
// Window of level j: slot (t mod 3) receives layer t-j, slots (t+2) mod 3 / (t+1) mod 3 hold layers t-j-1 / t-j-2.
// The responses of the row and the column (fx_lines_kernel) are added as the values leave the pipeline.
//   in_a / in_b : u_t / u_{t-1};  out_a / out_b : where u_{t+K} / u_{t+K-1} go (the same two buffers, swapped when
//   K is odd).  A thread only ever touches its own (r, c) column, layer t-K after reading layer t: in place is safe.
template <int K, bool UNDAMPED>
struct FxPipe {
    double w[K][3];
    double mprev, mcur;  // level -1 at layers t-1, t
};

// GUARD: bit 0 = mirror below the first layer (needed while t <= K), bit 1 = mirror above the last (needed from t >= L)
template <int K, bool UNDAMPED, int R, int GUARD>
__device__ __forceinline__ void fx_tick(FxPipe<K, UNDAMPED> &p, int t, int L, double lam, double cb, double coeff,
                                        double in0, double inm, double &outK, double &outK1)
{
    constexpr int SN = R % 3, SB = (R + 2) % 3, SA = (R + 1) % 3;  // newest, middle, oldest slot of this tick
    p.mprev = p.mcur;
    p.mcur = inm;
    p.w[0][SN] = in0;
#pragma unroll
    for (int k = 1; k <= K; ++k) {
        // level k at layer x = t - k from level k-1 at x-1, x, x+1 and level k-2 at x
        double a = p.w[k - 1][SA], b = p.w[k - 1][SB], c = p.w[k - 1][SN];
        const double prev = k >= 2 ? p.w[k - 2][SA] : p.mprev;
        if (GUARD & 1) {
            if (t - k <= 0) a = b;        // mirror below the first layer (uniform over the grid)
        }
        if (GUARD & 2) {
            if (t - k >= L - 1) c = b;    // mirror above the last layer
        }
        double n;
        if (UNDAMPED) n = fma(lam, b, -prev);
        else n = fma(lam, b, cb * prev);
        n = fma(coeff, a, n);
        n = fma(coeff, c, n);
        if (k < K) p.w[k][SN] = n;
        else outK = n;
    }
    outK1 = p.w[K - 1][SB];  // level K-1 at layer t-K (computed one tick ago)
}

// Sum / difference modes of the four quadrant values held by lanes 4j..4j+3: two butterfly rounds; m1 / m2 are sign
// masks for the high word (0x80000000 on odd lanes / on lanes with bit 1 set).  Applying it twice gives 4 x identity.
__device__ __forceinline__ double fx_flip(double v, int m) { return __hiloint2double(__double2hiint(v) ^ m, __double2loint(v)); }
__device__ __forceinline__ double hadamard4(double v, int m1, int m2)
{
    v = __shfl_xor_sync(0xffffffffu, v, 1) + fx_flip(v, m1);
    v = __shfl_xor_sync(0xffffffffu, v, 2) + fx_flip(v, m2);
    return v;
}

// Loads: every thread keeps a ring of PD ticks of its own inputs (u_t, u_{t-1} of layer t: 8 bytes each; the row and
// the column response of layer t-K, both time levels: 16 bytes each) in shared memory, filled with cp.async copies PD
// ticks ahead and completed with cp.async.wait_group -- a thread only ever reads what it copied itself, so there is no
// block-level barrier and no register is spent on the prefetch depth (the K-level pipeline needs them all).
constexpr int kFxPD = 6;                                    // ticks of loads in flight per thread
constexpr int kFxSlotBytes = 256 * (8 + 8 + 16 + 16);       // one tick of one CTA

template <int K, bool UNDAMPED>
struct FxCtx {
    FxPipe<K, UNDAMPED> p;
    // kernel arguments (uniform) and per-thread running offsets: every address is base + offset
    const double *in_a, *in_b, *yR, *yC;
    double *out_a, *out_b;
    size_t o;              // element offset of layer t of this thread's value (advances by `plane` per tick)
    size_t ol;             // element offset of the row response pair of layer t (advances by `lstride2` per tick)
    long long dcr;         // column response pair = row response pair + dcr
    unsigned char *r8, *r16;  // this thread's 8-byte / 16-byte places inside slot 0 of the ring
    double lam, cb, coeff;
    size_t plane, lstride2;
    int L, m1, m2;
    bool act;
};

// One tick.  R3 = t mod 3 is the rotation of the pipeline windows (compile time: it indexes registers); slot = t mod 6 is
// the ring slot (a constant in the unrolled steady loop, a run-time value elsewhere).
// MODE 0 = STEADY: K < t < L - PD, i.e. every stage is away from the end layers, the tick loads, requests and stores.
// MODE 1 / 2 / 3: guarded tick with the lower / upper / both end-layer mirrors (fx_tick GUARD bits); MODE 4: guarded tick
// that needs no mirror (K < t < L, only the loads and stores are conditional).  The mirrors cost two selects per level
// and side, so a tick pays only for the side it can reach.
template <int K, bool UNDAMPED, int R3, int MODE>
__device__ __forceinline__ void fx_bulk_tick(FxCtx<K, UNDAMPED> &c, int t, int slot)
{
    constexpr bool STEADY = MODE == 0;
    cp_async_wait<kFxPD - 1>();
    unsigned char *s8 = c.r8 + slot * kFxSlotBytes, *s16 = c.r16 + slot * kFxSlotBytes;
    double na = *reinterpret_cast<const double *>(s8);
    double nb = *reinterpret_cast<const double *>(s8 + 256 * 8);
    const double2 yr = *reinterpret_cast<const double2 *>(s16);
    const double2 yc = *reinterpret_cast<const double2 *>(s16 + 256 * 16);
    if (!STEADY && t >= c.L) {
        na = 0.0;
        nb = 0.0;
    }
    // request the inputs of tick t + PD into the slot just read (same thread: program order)
    if (STEADY || t + kFxPD < c.L) {
        cp_async8(s8, c.in_a + c.o + kFxPD * c.plane);
        cp_async8(s8 + 256 * 8, c.in_b + c.o + kFxPD * c.plane);
    }
    if (STEADY || (t + kFxPD >= K && t + kFxPD < c.L + K)) {  // responses of layer t + PD - K
        const double *pr = c.yR + c.ol + (long long)(kFxPD - K) * (long long)c.lstride2;
        cp_async16(s16, pr, true);
        cp_async16(s16 + 256 * 16, pr + c.dcr, true);
    }
    cp_async_commit();
    const double h0 = hadamard4(na, c.m1, c.m2), hm = hadamard4(nb, c.m1, c.m2);
    double oK, oK1;
    fx_tick<K, UNDAMPED, R3, (MODE == 0 || MODE == 4) ? 0 : MODE>(c.p, t, c.L, c.lam, c.cb, c.coeff, h0, hm, oK, oK1);
    if (STEADY || t >= K) {
        const double va = hadamard4(oK, c.m1, c.m2), vb = hadamard4(oK1, c.m1, c.m2);
        if (c.act) {  // .x = level K, .y = level K-1 (the host starts a pass on the slot that makes it so)
            c.out_a[c.o - K * c.plane] = fma(0.25, va, yr.x + yc.x);
            c.out_b[c.o - K * c.plane] = fma(0.25, vb, yr.y + yc.y);
        }
    }
    c.o += c.plane;
    c.ol += c.lstride2;
}

template <int K, bool UNDAMPED, int MINB>
__global__ void __launch_bounds__(256, MINB)
fx_bulk3d_kernel(const __grid_constant__ FxDev X, int halo, const double *in_a, const double *in_b, double *out_a,
                 double *out_b, const double *__restrict__ yR, const double *__restrict__ yC, double ca, double cb,
                 double coeff, int nct)
{
    extern __shared__ __align__(16) unsigned char tk_smem[];
    const int S = X.S, L = X.L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ct = blockIdx.x % nct;
    const int bi = blockIdx.x / nct;
    const int r = X.br_list[bi];  // bulk rows only
    const int c = ct * 64 + warp * 8 + (lane >> 2), m = lane & 3;
    const int cc = min(c, S - 1);  // lanes beyond the row / on cross columns compute on valid memory and store nothing
    FxCtx<K, UNDAMPED> x;
    x.act = c < S && !X.colflag[cc];
    // a warp whose 8 cells are all cross cells / beyond the row has nothing to do (no block-level sync below)
    if (!__any_sync(0xffffffffu, x.act)) return;
    x.in_a = in_a; x.in_b = in_b; x.out_a = out_a; x.out_b = out_b; x.yR = yR; x.yC = yC;
    x.plane = (size_t)S * S * 4;
    x.lstride2 = (size_t)S * 8;
    x.L = L;
    x.o = (size_t)halo * x.plane + ((size_t)r * S + cc) * 4 + m;
    x.ol = ((size_t)r * 4 + m) * 2;
    x.dcr = (yC - yR) + ((long long)cc - r) * 8;
    x.r8 = tk_smem + threadIdx.x * 8;
    x.r16 = tk_smem + 256 * 16 + threadIdx.x * 16;
    x.lam = m == 0 ? ca + 4.0 * coeff : ca;
    x.cb = cb;
    x.coeff = coeff;
    x.m1 = (lane & 1) ? (int)0x80000000 : 0;
    x.m2 = (lane & 2) ? (int)0x80000000 : 0;
#pragma unroll
    for (int k = 0; k < K; ++k) x.p.w[k][0] = x.p.w[k][1] = x.p.w[k][2] = 0.0;
    x.p.mprev = x.p.mcur = 0.0;
    // prime the ring: ticks 0 .. PD-1
#pragma unroll
    for (int d = 0; d < kFxPD; ++d) {
        if (d < L) {
            cp_async8(x.r8 + d * kFxSlotBytes, in_a + x.o + d * x.plane);
            cp_async8(x.r8 + d * kFxSlotBytes + 256 * 8, in_b + x.o + d * x.plane);
        }
        if (d >= K && d < L + K) {
            const double *pr = yR + x.ol + (long long)(d - K) * (long long)x.lstride2;
            cp_async16(x.r16 + d * kFxSlotBytes, pr, true);
            cp_async16(x.r16 + d * kFxSlotBytes + 256 * 16, pr + x.dcr, true);
        }
        cp_async_commit();
    }
    const int nt = L + K;
    auto guarded = [&](int t) {  // any tick: rotation and the mirrors it can reach picked at run time
        const int slot = t % 6;
        const int mode = (t <= K ? 1 : 0) | (t >= L ? 2 : 0);  // level k sits at layer t - k, k = 1 .. K
        switch (mode * 3 + t % 3) {
        case 0: fx_bulk_tick<K, UNDAMPED, 0, 4>(x, t, slot); break;
        case 1: fx_bulk_tick<K, UNDAMPED, 1, 4>(x, t, slot); break;
        case 2: fx_bulk_tick<K, UNDAMPED, 2, 4>(x, t, slot); break;
        case 3: fx_bulk_tick<K, UNDAMPED, 0, 1>(x, t, slot); break;
        case 4: fx_bulk_tick<K, UNDAMPED, 1, 1>(x, t, slot); break;
        case 5: fx_bulk_tick<K, UNDAMPED, 2, 1>(x, t, slot); break;
        case 6: fx_bulk_tick<K, UNDAMPED, 0, 2>(x, t, slot); break;
        case 7: fx_bulk_tick<K, UNDAMPED, 1, 2>(x, t, slot); break;
        case 8: fx_bulk_tick<K, UNDAMPED, 2, 2>(x, t, slot); break;
        case 9: fx_bulk_tick<K, UNDAMPED, 0, 3>(x, t, slot); break;
        case 10: fx_bulk_tick<K, UNDAMPED, 1, 3>(x, t, slot); break;
        default: fx_bulk_tick<K, UNDAMPED, 2, 3>(x, t, slot); break;
        }
    };
    int t = 0;
    const int t_steady = (K + 1 + 5) / 6 * 6;  // first multiple of 6 above K
    // prologue (guarded), steady loop, epilogue (guarded): ONE call site for the guarded ticks -- the twelve specialised
    // bodies are inlined once, not once per loop (with two call sites the compiler stops inlining and the pipeline
    // registers turn into a stack frame)
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {
        const int t_end = phase == 0 ? min(nt, t_steady) : nt;
#pragma unroll 1
        for (; t < t_end; ++t) guarded(t);
        if (phase == 0) {
#pragma unroll 1
            for (; t + 6 <= L - kFxPD; t += 6) {
                fx_bulk_tick<K, UNDAMPED, 0, 0>(x, t, 0);
                fx_bulk_tick<K, UNDAMPED, 1, 0>(x, t + 1, 1);
                fx_bulk_tick<K, UNDAMPED, 2, 0>(x, t + 2, 2);
                fx_bulk_tick<K, UNDAMPED, 0, 0>(x, t + 3, 3);
                fx_bulk_tick<K, UNDAMPED, 1, 0>(x, t + 4, 4);
                fx_bulk_tick<K, UNDAMPED, 2, 0>(x, t + 5, 5);
            }
        }
    }
    cp_async_wait<0>();
}

}  // namespace tk
