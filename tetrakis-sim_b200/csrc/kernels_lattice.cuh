// Structured ("clique-sum") leapfrog step for tetrakis lattices -- the hot kernel.
//
// Replaces the triple Python loop of reference tetrakis_sim/physics.py:106-129 on graphs built by
// lattice.py:15-76 (+ defects.py).  Every row and column of same-quadrant nodes is a clique
// (lattice.py:37-44, 61-66), so the neighbour sum collapses to
//     (CellSum - u) + (RowSum[z,r,q] - u) + (ColSum[z,c,q] - u) + u(z-1) + u(z+1)
// and one fused, memory-bound pass per step suffices:
//     read u_t, read u_{t-1}, write u_{t+1} in place of u_{t-1}     = 24 B / node-update (fp64)
// while producing the partial row / column sums of u_{t+1} for the next step.
//
// Layout: state[plane][r][c][q], q fastest (32 B per cell); plane = local layer + halo.
// Work decomposition: a warp owns a segment of 32*CPL consecutive cells of one (layer, row) --
// CPL*1 KB contiguous per array -- and marches down `rchunk` rows, so column sums accumulate in
// registers and row sums need one transposing warp reduction per row.  A CTA is 8 warps: WC
// adjacent segments x WZ = 8/WC consecutive layers (the z+-1 neighbours of the interior warps are
// the same lines their sibling warps stream in, so they hit L1/L2 instead of HBM).
// Defects are sparse: a per-(layer,row,64-cell) table says "plain" (-1) or points at a detail
// record (participation bits, mass/potential, edge corrections) used by a slower path.
#pragma once
#include "common.cuh"

namespace tk {

constexpr int kTabSeg = 64;  // cells per defect-table segment (independent of CPL)

struct LatDev {
    int S;          // cells per side
    int Lloc;       // owned layers
    int halo;       // 1 if halo planes exist (3-D), 0 for a 2-D sheet
    int ntab;       // table segments per row = ceil(S / 64)
    int has_below;  // a layer exists below local layer 0 (z_begin > 0)
    int has_above;  // a layer exists above local layer Lloc-1
    const int *seg_flag;            // [Lloc][S][ntab]   -1 plain, else detail index
    const unsigned short *det_flags;  // [ndet][64] bits 0-3 part(q) 4-7 alive(q) 8-11 part(z-1,q) 12-15 part(z+1,q)
    const int *det_attr;            // [ndet] -1 or row of attr_im / attr_v
    const int2 *det_corr;           // [ndet] [begin, end) into corr_*
    const double *attr_im;          // [nattr][256] inv_mass per node of the segment
    const double *attr_v;           // [nattr][256] potential
    const long long *corr_a;        // padded-local node index that receives the correction
    const long long *corr_b;        // padded-local node index whose u is added
    const double *corr_sign;
    const double *rowinfo;  // [Lloc][S][8]: 0-3 RowSum(q) of the current state, 4-7 row participation count
    const double *colinfo;  // [Lloc][S][8]: same for columns
    // Fused halo push (multi-GPU slabs): when set, the warps that update the first / last owned layer
    // also store u_{t+1} straight into the z-neighbour's halo plane through an NVLink-mapped peer
    // pointer (cudaIpc), so the boundary update and the halo exchange are one kernel.
    double *peer_lo;        // lower neighbour's upper-halo plane of the buffer being written, or null
    double *peer_hi;        // upper neighbour's lower-halo plane, or null
    // Early arrival signal (optional, lat_step_kernel only): the warps that update the first / last owned layer are
    // the only ones that read this rank's halos and push into the neighbours'; the last of them to finish writes the
    // step's flag value into the neighbour's arrival flag from inside the kernel (fence + counter), so the neighbour's
    // stream-ordered wait is satisfied while the interior is still being updated.  The CTAs of the two boundary layer
    // groups are scheduled first (bnd_first) so that this happens early in the launch.
    unsigned long long *sig_lo, *sig_hi;  // neighbours' arrival flags (peer-mapped), or null
    unsigned long long sig_value;
    unsigned int *sig_cnt;                // local arrival counters: [0] lower side, [32] upper side; zero between launches
    unsigned int sig_target;              // warps per boundary layer in this launch
    int bnd_first;
};

__device__ __forceinline__ unsigned plain_flags(const LatDev &P, int zl)
{
    unsigned f = 0xFFu;
    if (zl > 0 || P.has_below) f |= 0xF00u;
    if (zl < P.Lloc - 1 || P.has_above) f |= 0xF000u;
    return f;
}

// Inputs / outputs of one cell (4 quadrant nodes) on the defect-aware path.  Passed by reference
// to an out-of-line function so that the rarely taken path costs the streaming path no registers.
struct CellIO {
    double uq[4], upq[4], zm[4], zp[4], RS[4], RC[4], CS[4], CC[4];
    double un[4], pun[4];
    double e;  // sum over the cell of m (u+ - u)^2 / coeff + u+ (V u - lap): twice the staggered energy
};

//   io.un[q] <- u_{t+1};  io.pun[q] <- participation-masked u_{t+1} (what row/col sums accumulate)
__device__ __noinline__ void cell_update_detail(const LatDev &P, const double *__restrict__ ustate,
                                                int det, int cell_in_tab, long long node0,
                                                unsigned f, double coeff, double damping,
                                                double inv_coeff, CellIO &io)
{
    double cell = 0.0;
    io.e = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if ((f >> q) & 1u) cell += io.uq[q];
    const double ccnt = (double)__popc(f & 0xFu);
    int attr = -1;
    int2 cr = make_int2(0, 0);
    if (det >= 0) {
        attr = P.det_attr[det];
        cr = P.det_corr[det];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const bool part = (f >> q) & 1u, alive = (f >> (4 + q)) & 1u;
        const double uq = io.uq[q], upq = io.upq[q];
        double lap = 0.0;
        if (part) {
            const bool pm = (f >> (8 + q)) & 1u, pp = (f >> (12 + q)) & 1u;
            double nsum = (cell - uq) + (io.RS[q] - uq) + (io.CS[q] - uq);
            double deg = (ccnt - 1.0) + (io.RC[q] - 1.0) + (io.CC[q] - 1.0);
            if (pm) { nsum += io.zm[q]; deg += 1.0; }
            if (pp) { nsum += io.zp[q]; deg += 1.0; }
            lap = nsum - deg * uq;
        }
        for (int k = cr.x; k < cr.y; ++k)
            if (P.corr_a[k] == node0 + q) lap += P.corr_sign[k] * (ustate[P.corr_b[k]] - uq);
        double im = 1.0, v = 0.0;
        if (attr >= 0) {
            im = P.attr_im[(size_t)attr * 256 + cell_in_tab * 4 + q];
            v = P.attr_v[(size_t)attr * 256 + cell_in_tab * 4 + q];
        }
        const double r = 2.0 * uq - upq + (coeff * im) * lap - (coeff * v) * uq - damping * (uq - upq);
        io.un[q] = alive ? r : 0.0;
        io.pun[q] = part ? io.un[q] : 0.0;
        if (alive) io.e += (r - uq) * (r - uq) * inv_coeff / im + r * (v * uq - lap);
    }
}

// Same for the fp32 state layout: identical arithmetic in fp64, the correction partners are read as floats.
__device__ __noinline__ void cell_update_detail_f32(const LatDev &P, const float *__restrict__ ustate,
                                                    int det, int cell_in_tab, long long node0,
                                                    unsigned f, double coeff, double damping, CellIO &io)
{
    double cell = 0.0;
    io.e = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if ((f >> q) & 1u) cell += io.uq[q];
    const double ccnt = (double)__popc(f & 0xFu);
    int attr = -1;
    int2 cr = make_int2(0, 0);
    if (det >= 0) {
        attr = P.det_attr[det];
        cr = P.det_corr[det];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const bool part = (f >> q) & 1u, alive = (f >> (4 + q)) & 1u;
        const double uq = io.uq[q], upq = io.upq[q];
        double lap = 0.0;
        if (part) {
            const bool pm = (f >> (8 + q)) & 1u, pp = (f >> (12 + q)) & 1u;
            double nsum = (cell - uq) + (io.RS[q] - uq) + (io.CS[q] - uq);
            double deg = (ccnt - 1.0) + (io.RC[q] - 1.0) + (io.CC[q] - 1.0);
            if (pm) { nsum += io.zm[q]; deg += 1.0; }
            if (pp) { nsum += io.zp[q]; deg += 1.0; }
            lap = nsum - deg * uq;
        }
        for (int k = cr.x; k < cr.y; ++k)
            if (P.corr_a[k] == node0 + q) lap += P.corr_sign[k] * ((double)ustate[P.corr_b[k]] - uq);
        double im = 1.0, v = 0.0;
        if (attr >= 0) {
            im = P.attr_im[(size_t)attr * 256 + cell_in_tab * 4 + q];
            v = P.attr_v[(size_t)attr * 256 + cell_in_tab * 4 + q];
        }
        const double r = 2.0 * uq - upq + (coeff * im) * lap - (coeff * v) * uq - damping * (uq - upq);
        io.un[q] = alive ? r : 0.0;
        io.pun[q] = part ? io.un[q] : 0.0;
    }
}

// MODE 0: leapfrog step.  MODE 1: sums only (row/col sums of the current state, no update).
template <int CPL, int WC, bool K3D, int MODE, int MINB = 1>
__global__ void __launch_bounds__(256, MINB)
lat_step_kernel(const __grid_constant__ LatDev P, const double *__restrict__ u,
                double *__restrict__ w, double *__restrict__ prow, double *__restrict__ pcol,
                double coeff, double damping, int zl_begin, int zl_end, int rchunk, int nchunk,
                int nsegk, int nct, int nzg, int pf_rows, double *__restrict__ pen, double inv_coeff)
{
    constexpr int WZ = 8 / WC;
    constexpr bool STEP = MODE != 1;  // MODE 0: step, 1: sums only, 2: step + energy partials
    constexpr bool KEN = MODE == 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int b = blockIdx.x;
    int zg;
    if (K3D && P.bnd_first && nzg >= 3) {  // the two boundary layer groups first, then the interior groups interleaved
        const int nb2 = 2 * nct * nchunk;
        if (b < nb2) {
            zg = (b & 1) ? nzg - 1 : 0;
            b >>= 1;
        } else {
            b -= nb2;
            zg = 1 + b % (nzg - 2);
            b /= nzg - 2;
        }
    } else {
        zg = b % nzg;
        b /= nzg;
    }
    const int ct = b % nct;
    const int ch = b / nct;
    const int zl = zl_begin + zg * WZ + warp / WC;
    const int segk = ct * WC + warp % WC;
    if (zl >= zl_end || segk >= nsegk) {  // whole warp leaves; no block-level sync below
        if (KEN && lane == 0) pen[(size_t)blockIdx.x * 8 + warp] = 0.0;
        return;
    }

    const int S = P.S;
    const int c0 = (segk * 32 + lane) * CPL;
    bool valid[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) valid[j] = (c0 + j) < S;
    const size_t plane = (size_t)S * S * 4;
    const size_t zbase = (size_t)(zl + P.halo) * plane;
    const unsigned fplain = plain_flags(P, zl);
    const bool zm_on = K3D && (fplain & 0x100u), zp_on = K3D && (fplain & 0x1000u);
    const double degbase = 4.0 + (zm_on ? 1.0 : 0.0) + (zp_on ? 1.0 : 0.0);

    // column context: constant while marching down the rows
    double CS[CPL][4], CC[CPL][4], cacc[CPL][4];
#pragma unroll
    for (int j = 0; j < CPL; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            CS[j][q] = 0.0;
            CC[j][q] = 0.0;
            cacc[j][q] = 0.0;
        }
    if (STEP) {
#pragma unroll
        for (int j = 0; j < CPL; ++j)
            if (valid[j]) {
                const double2 *ci = reinterpret_cast<const double2 *>(
                    P.colinfo + ((size_t)zl * S + (c0 + j)) * 8);
                const double2 a = ci[0], bb = ci[1], c = ci[2], d = ci[3];
                CS[j][0] = a.x; CS[j][1] = a.y; CS[j][2] = bb.x; CS[j][3] = bb.y;
                CC[j][0] = c.x; CC[j][1] = c.y; CC[j][2] = d.x; CC[j][3] = d.y;
            }
    }

    double eacc = 0.0;  // energy partial of this lane (MODE 2)
    const int r_begin = ch * rchunk;
    const int r_end = min(S, r_begin + rchunk);
    const int tab0 = c0 / kTabSeg;  // all CPL cells of a lane share a table segment (CPL | 64)

    // L2 prefetch `pf_rows` rows ahead (one request per 128-byte line): turns the HBM latency of the
    // streamed rows into an L2 hit without spending registers on software pipelining.
    const bool pf_lane = STEP && pf_rows > 0 && valid[0] && ((lane * CPL) & 3) == 0;
    if (pf_lane) {
        for (int r = r_begin; r < min(r_end, r_begin + pf_rows); ++r) {
            const size_t o = zbase + ((size_t)r * S + c0) * 4;
            prefetch_l2(u + o);
            prefetch_l2(w + o);
        }
    }

    for (int r = r_begin; r < r_end; ++r) {
        const size_t rowoff = zbase + ((size_t)r * S + c0) * 4;
        if (pf_lane && r + pf_rows < r_end) {
            const size_t o = rowoff + (size_t)pf_rows * S * 4;
            prefetch_l2(u + o);
            prefetch_l2(w + o);
        }
        double uq[CPL][4], upq[CPL][4], zm[CPL][4], zp[CPL][4];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            if (valid[j]) {
                ld256_nc(u + rowoff + j * 4, uq[j]);
                if (STEP) ld256_cs(w + rowoff + j * 4, upq[j]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) { uq[j][q] = 0.0; upq[j][q] = 0.0; }
            }
            if (!STEP) {
#pragma unroll
                for (int q = 0; q < 4; ++q) upq[j][q] = 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) { zm[j][q] = 0.0; zp[j][q] = 0.0; }
            if (STEP && K3D && valid[j]) {
                if (zm_on) ld256_nc(u + rowoff - plane + j * 4, zm[j]);
                if (zp_on) ld256_nc(u + rowoff + plane + j * 4, zp[j]);
            }
        }
        int det = -1;
        if (c0 < S) det = P.seg_flag[((size_t)zl * S + r) * P.ntab + tab0];
        const bool all_plain = __all_sync(0xffffffffu, det < 0);

        double RS[4] = {0, 0, 0, 0}, RC[4] = {0, 0, 0, 0};
        if (STEP) {
            const double2 *ri = reinterpret_cast<const double2 *>(P.rowinfo + ((size_t)zl * S + r) * 8);
            const double2 a = ri[0], bb = ri[1], c = ri[2], d = ri[3];
            RS[0] = a.x; RS[1] = a.y; RS[2] = bb.x; RS[3] = bb.y;
            RC[0] = c.x; RC[1] = c.y; RC[2] = d.x; RC[3] = d.y;
        }

        double rsum[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            double un[4], pun[4];
            if (MODE == 1) {
                unsigned f = fplain;
                if (det >= 0) f = P.det_flags[(size_t)det * kTabSeg + ((c0 + j) & (kTabSeg - 1))];
                if (!valid[j]) f = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) pun[q] = ((f >> q) & 1u) ? uq[j][q] : 0.0;
            } else if (all_plain) {
                const double cell = (uq[j][0] + uq[j][1]) + (uq[j][2] + uq[j][3]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double nb = (cell + RS[q]) + (CS[j][q] + (zm[j][q] + zp[j][q]));
                    const double dg = (RC[q] + degbase) + CC[j][q];
                    const double lap = nb - dg * uq[j][q];
                    un[q] = 2.0 * uq[j][q] - upq[j][q] + coeff * lap -
                            damping * (uq[j][q] - upq[j][q]);
                    pun[q] = valid[j] ? un[q] : 0.0;
                    if (KEN && valid[j]) {
                        const double dv = un[q] - uq[j][q];
                        eacc += dv * dv * inv_coeff - un[q] * lap;
                    }
                }
            } else {
                unsigned f = fplain;
                if (det >= 0) f = P.det_flags[(size_t)det * kTabSeg + ((c0 + j) & (kTabSeg - 1))];
                if (!valid[j]) f = 0;
                CellIO io;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    io.uq[q] = uq[j][q]; io.upq[q] = upq[j][q]; io.zm[q] = zm[j][q]; io.zp[q] = zp[j][q];
                    io.RS[q] = RS[q]; io.RC[q] = RC[q]; io.CS[q] = CS[j][q]; io.CC[q] = CC[j][q];
                }
                cell_update_detail(P, u, det, (c0 + j) & (kTabSeg - 1), (long long)(rowoff + j * 4), f,
                                   coeff, damping, inv_coeff, io);
#pragma unroll
                for (int q = 0; q < 4; ++q) { un[q] = io.un[q]; pun[q] = io.pun[q]; }
                if (KEN) eacc += io.e;
            }
            if (STEP && valid[j]) {
                st256_cs(w + rowoff + j * 4, un);
                if (K3D) {  // warp-uniform: zl is per warp
                    if (zl == 0 && P.peer_lo) st256(P.peer_lo + (rowoff - zbase) + j * 4, un);
                    if (zl == P.Lloc - 1 && P.peer_hi) st256(P.peer_hi + (rowoff - zbase) + j * 4, un);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                cacc[j][q] += pun[q];
                rsum[q] += pun[q];
            }
        }
        warp_sum4_store(rsum, lane, prow + (((size_t)zl * S + r) * nsegk + segk) * 4);
    }
#pragma unroll
    for (int j = 0; j < CPL; ++j)
        if (valid[j]) st256(pcol + (((size_t)zl * nchunk + ch) * S + (c0 + j)) * 4, cacc[j]);
    if (KEN) {
        eacc = warp_sum(eacc);
        if (lane == 0) pen[(size_t)blockIdx.x * 8 + warp] = eacc;
    }
    if (K3D && STEP) {  // early arrival signal (see LatDev): warp-uniform conditions
        const bool lo = zl == 0 && P.sig_lo != nullptr, hi = zl == P.Lloc - 1 && P.sig_hi != nullptr;
        if (lo || hi) {
            __threadfence_system();  // this lane's pushes are visible to the neighbour before the count moves
            __syncwarp();
            if (lane == 0) {
                if (lo && atomicAdd(P.sig_cnt, 1u) == P.sig_target - 1) {
                    __threadfence_system();
                    atomicExch(P.sig_cnt, 0u);
                    *reinterpret_cast<volatile unsigned long long *>(P.sig_lo) = P.sig_value;
                }
                if (hi && atomicAdd(P.sig_cnt + 32, 1u) == P.sig_target - 1) {
                    __threadfence_system();
                    atomicExch(P.sig_cnt + 32, 0u);
                    *reinterpret_cast<volatile unsigned long long *>(P.sig_hi) = P.sig_value;
                }
            }
        }
    }
}

// TMA variant of lat_step_kernel (fp64, one cell per lane): every warp is its own producer and consumer.
//
// Lane 0 issues cp.async.bulk copies (SASS UBLKCP) of the warp's next row segments -- u, u_prev, in 3-D also
// the z-1 / z+1 segments, plus the row's 64-byte {RowSum, rowcnt} record and its defect-table entry -- into a
// per-warp ring of STAGES shared-memory stages, each completing on its own mbarrier; all lanes wait on the
// stage's phase, copy their 32 bytes to registers, and lane 0 re-arms the stage for row r+STAGES.  No CTA-wide
// barrier (the cp.async version's problem), no registers spent on prefetching, STAGES rows in flight per warp.
template <int WC, bool K3D, int STAGES, int MINB>
__global__ void __launch_bounds__(256, MINB)
lat_step_bulk_kernel(const __grid_constant__ LatDev P, const double *__restrict__ u, double *__restrict__ w,
                     double *__restrict__ prow, double *__restrict__ pcol, double coeff, double damping,
                     int zl_begin, int zl_end, int rchunk, int nchunk, int nsegk, int nct, int nzg)
{
    constexpr int WZ = 8 / WC;
    constexpr int NARR = K3D ? 4 : 2;
    constexpr int STAGE_DOUBLES = NARR * 128 + 16;  // + 8 doubles rowinfo + 16 B table entry (padded to 128 B)
    extern __shared__ __align__(16) unsigned char tk_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *ring = reinterpret_cast<double *>(tk_smem) + (size_t)warp * STAGES * STAGE_DOUBLES;
    unsigned long long *bars =
        reinterpret_cast<unsigned long long *>(tk_smem + (size_t)8 * STAGES * STAGE_DOUBLES * sizeof(double)) + warp * STAGES;
    int b = blockIdx.x;
    const int zg = b % nzg;
    b /= nzg;
    const int ct = b % nct;
    const int ch = b / nct;
    const int zl = zl_begin + zg * WZ + warp / WC;
    const int segk = ct * WC + warp % WC;
    if (zl >= zl_end || segk >= nsegk) return;  // per-warp barriers only: a whole warp may leave

    const int S = P.S;
    const int cw = segk * 32;                   // first cell of the warp's segment
    const int c0 = cw + lane;
    const bool valid = c0 < S;
    const unsigned seg_bytes = (unsigned)min(32, S - cw) * 32u;
    const size_t plane = (size_t)S * S * 4;
    const size_t zbase = (size_t)(zl + P.halo) * plane;
    const unsigned fplain = plain_flags(P, zl);
    const bool zm_on = K3D && (fplain & 0x100u), zp_on = K3D && (fplain & 0x1000u);
    const double degbase = 4.0 + (zm_on ? 1.0 : 0.0) + (zp_on ? 1.0 : 0.0);
    const int tab0 = cw / kTabSeg;

    double CS[4] = {0, 0, 0, 0}, CC[4] = {0, 0, 0, 0}, cacc[4] = {0, 0, 0, 0};
    if (valid) {
        ld256_nc(P.colinfo + ((size_t)zl * S + c0) * 8, CS);
        ld256_nc(P.colinfo + ((size_t)zl * S + c0) * 8 + 4, CC);
    }
    const int r_begin = ch * rchunk;
    const int r_end = min(S, r_begin + rchunk);

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
        fence_proxy_async();  // make the initialised barriers visible to the async proxy
    }
    __syncwarp();
    const unsigned tx = seg_bytes * (2u + (zm_on ? 1u : 0u) + (zp_on ? 1u : 0u)) + 64u + 16u;
    auto issue = [&](int r, int s) {  // lane 0
        double *st = ring + (size_t)s * STAGE_DOUBLES;
        const size_t off = zbase + ((size_t)r * S + cw) * 4;
        mbar_expect_tx(&bars[s], tx);
        bulk_g2s(st, u + off, seg_bytes, &bars[s]);
        bulk_g2s(st + 128, w + off, seg_bytes, &bars[s]);
        if (zm_on) bulk_g2s(st + 256, u + off - plane, seg_bytes, &bars[s]);
        if (zp_on) bulk_g2s(st + 384, u + off + plane, seg_bytes, &bars[s]);
        bulk_g2s(st + NARR * 128, P.rowinfo + ((size_t)zl * S + r) * 8, 64u, &bars[s]);
        // the 4-byte table entry travels inside its 16-byte aligned quad (the table is padded at the end)
        const size_t fi = ((size_t)zl * S + r) * P.ntab + tab0;
        bulk_g2s(st + NARR * 128 + 8, P.seg_flag + (fi & ~(size_t)3), 16u, &bars[s]);
    };
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; ++i)
            if (r_begin + i < r_end) issue(r_begin + i, i);
    }
    int s = 0;
    unsigned parity = 0;
    for (int r = r_begin; r < r_end; ++r) {
        const double *st = ring + (size_t)s * STAGE_DOUBLES;
        mbar_wait(&bars[s], parity);
        double uq[4], upq[4], zm[4] = {0, 0, 0, 0}, zp[4] = {0, 0, 0, 0}, RS[4], RC[4];
        {
            const double2 a = *reinterpret_cast<const double2 *>(st + lane * 4);
            const double2 bb = *reinterpret_cast<const double2 *>(st + lane * 4 + 2);
            uq[0] = a.x; uq[1] = a.y; uq[2] = bb.x; uq[3] = bb.y;
            const double2 c = *reinterpret_cast<const double2 *>(st + 128 + lane * 4);
            const double2 d = *reinterpret_cast<const double2 *>(st + 128 + lane * 4 + 2);
            upq[0] = c.x; upq[1] = c.y; upq[2] = d.x; upq[3] = d.y;
            if (zm_on) {
                const double2 e = *reinterpret_cast<const double2 *>(st + 256 + lane * 4);
                const double2 f = *reinterpret_cast<const double2 *>(st + 256 + lane * 4 + 2);
                zm[0] = e.x; zm[1] = e.y; zm[2] = f.x; zm[3] = f.y;
            }
            if (zp_on) {
                const double2 e = *reinterpret_cast<const double2 *>(st + 384 + lane * 4);
                const double2 f = *reinterpret_cast<const double2 *>(st + 384 + lane * 4 + 2);
                zp[0] = e.x; zp[1] = e.y; zp[2] = f.x; zp[3] = f.y;
            }
            const double2 r0 = *reinterpret_cast<const double2 *>(st + NARR * 128);
            const double2 r1 = *reinterpret_cast<const double2 *>(st + NARR * 128 + 2);
            const double2 r2 = *reinterpret_cast<const double2 *>(st + NARR * 128 + 4);
            const double2 r3 = *reinterpret_cast<const double2 *>(st + NARR * 128 + 6);
            RS[0] = r0.x; RS[1] = r0.y; RS[2] = r1.x; RS[3] = r1.y;
            RC[0] = r2.x; RC[1] = r2.y; RC[2] = r3.x; RC[3] = r3.y;
        }
        const size_t fi = ((size_t)zl * S + r) * P.ntab + tab0;
        const int det_w = reinterpret_cast<const int *>(st + NARR * 128 + 8)[fi & 3];
        if (!valid) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { uq[q] = 0.0; upq[q] = 0.0; zm[q] = 0.0; zp[q] = 0.0; }
        }
        __syncwarp();  // every lane has taken its values: the stage may be refilled
        if (lane == 0 && r + STAGES < r_end) {
            fence_proxy_async();
            issue(r + STAGES, s);
        }
        const int det = valid ? det_w : -1;
        const bool all_plain = det_w < 0;  // one table entry per 64 cells: warp-uniform
        const size_t rowoff = zbase + ((size_t)r * S + c0) * 4;
        double un[4], pun[4];
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
            if (det >= 0) f = P.det_flags[(size_t)det * kTabSeg + (c0 & (kTabSeg - 1))];
            if (!valid) f = 0;
            CellIO io;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                io.uq[q] = uq[q]; io.upq[q] = upq[q]; io.zm[q] = zm[q]; io.zp[q] = zp[q];
                io.RS[q] = RS[q]; io.RC[q] = RC[q]; io.CS[q] = CS[q]; io.CC[q] = CC[q];
            }
            cell_update_detail(P, u, det, c0 & (kTabSeg - 1), (long long)rowoff, f, coeff, damping, 0.0, io);
#pragma unroll
            for (int q = 0; q < 4; ++q) { un[q] = io.un[q]; pun[q] = io.pun[q]; }
        }
        if (valid) {
            st256_cs(w + rowoff, un);
            if (K3D) {
                if (zl == 0 && P.peer_lo) st256(P.peer_lo + (rowoff - zbase), un);
                if (zl == P.Lloc - 1 && P.peer_hi) st256(P.peer_hi + (rowoff - zbase), un);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) cacc[q] += pun[q];
        warp_sum4_store(pun, lane, prow + (((size_t)zl * S + r) * nsegk + segk) * 4);
        if (++s == STAGES) { s = 0; parity ^= 1u; }
    }
    if (valid) st256(pcol + (((size_t)zl * nchunk + ch) * S + c0) * 4, cacc);
}

// fp32 state variant of lat_step_kernel (optional mode; 12 B/update).
//
// Storage and the update expression are fp32; the Laplacian -- a difference of two sums of up to
// ~16 000 terms -- is formed in fp64 from the fp64 row/column sums (which are accumulated and kept in
// fp64 exactly as in the fp64 path) and rounded once, so the only fp32 errors are the per-step
// rounding of the state itself.  One cell per lane (16-byte LDG/STG.128), same tiling, same partial
// sum layout, same defect tables, same peer halo push as the fp64 kernel.  Tolerance: see
// tests/test_gpu_fp32.py (documented in DESIGN.md).
template <int CPL, int WC, bool K3D, int MODE, int MINB>
__global__ void __launch_bounds__(256, MINB)
lat_step_f32_kernel(const __grid_constant__ LatDev P, const float *__restrict__ u,
                    float *__restrict__ w, double *__restrict__ prow, double *__restrict__ pcol,
                    double coeff, double damping, int zl_begin, int zl_end, int rchunk, int nchunk,
                    int nsegk, int nct, int nzg, int pf_rows)
{
    static_assert(CPL == 1 || CPL == 2, "fp32 kernel: 1 or 2 cells per lane");
    constexpr int WZ = 8 / WC;
    constexpr bool STEP = MODE != 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int b = blockIdx.x;
    const int zg = b % nzg;
    b /= nzg;
    const int ct = b % nct;
    const int ch = b / nct;
    const int zl = zl_begin + zg * WZ + warp / WC;
    const int segk = ct * WC + warp % WC;
    if (zl >= zl_end || segk >= nsegk) return;

    const int S = P.S;
    const int c0 = (segk * 32 + lane) * CPL;
    bool valid[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) valid[j] = (c0 + j) < S;
    const bool both = CPL == 2 && valid[CPL - 1];  // the lane's cells are contiguous: one 256-bit access
    const size_t plane = (size_t)S * S * 4;
    const size_t zbase = (size_t)(zl + P.halo) * plane;
    const unsigned fplain = plain_flags(P, zl);
    const bool zm_on = K3D && (fplain & 0x100u), zp_on = K3D && (fplain & 0x1000u);
    const double degbase = 4.0 + (zm_on ? 1.0 : 0.0) + (zp_on ? 1.0 : 0.0);
    const float coeff_f = (float)coeff, damp_f = (float)damping;
    float *peer_lo = reinterpret_cast<float *>(P.peer_lo), *peer_hi = reinterpret_cast<float *>(P.peer_hi);

    double CS[CPL][4], CC[CPL][4], cacc[CPL][4];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { CS[j][q] = 0.0; CC[j][q] = 0.0; cacc[j][q] = 0.0; }
        if (STEP && valid[j]) {
            ld256_nc(P.colinfo + ((size_t)zl * S + c0 + j) * 8, CS[j]);
            ld256_nc(P.colinfo + ((size_t)zl * S + c0 + j) * 8 + 4, CC[j]);
        }
    }
    const int r_begin = ch * rchunk;
    const int r_end = min(S, r_begin + rchunk);
    const int tab0 = c0 / kTabSeg;
    const bool pf_lane = STEP && pf_rows > 0 && valid[0] && ((lane * CPL) & 7) == 0;  // one per 128 B
    if (pf_lane) {
        for (int r = r_begin; r < min(r_end, r_begin + pf_rows); ++r) {
            const size_t o = zbase + ((size_t)r * S + c0) * 4;
            prefetch_l2(u + o);
            prefetch_l2(w + o);
        }
    }
    // rows of odd length can leave a lane's second cell misaligned for a 256-bit access
    const bool wide_ok = CPL == 2 && (S & 1) == 0;
    auto load = [&](const float *base, float v[CPL * 4], bool streaming) {
        if (CPL == 2 && both && wide_ok) {
            if (streaming) ld256f_cs(base, v); else ld256f_nc(base, v);
            return;
        }
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid[j]) t = streaming ? ld128_cs(base + j * 4) : ld128_nc(base + j * 4);
            v[j * 4 + 0] = t.x; v[j * 4 + 1] = t.y; v[j * 4 + 2] = t.z; v[j * 4 + 3] = t.w;
        }
    };
    for (int r = r_begin; r < r_end; ++r) {
        const size_t rowoff = zbase + ((size_t)r * S + c0) * 4;
        if (pf_lane && r + pf_rows < r_end) {
            const size_t o = rowoff + (size_t)pf_rows * S * 4;
            prefetch_l2(u + o);
            prefetch_l2(w + o);
        }
        float uf[CPL * 4], upf[CPL * 4], zmf[CPL * 4], zpf[CPL * 4];
#pragma unroll
        for (int i = 0; i < CPL * 4; ++i) { upf[i] = 0.f; zmf[i] = 0.f; zpf[i] = 0.f; }
        load(u + rowoff, uf, false);
        if (STEP) {
            load(w + rowoff, upf, true);
            if (zm_on) load(u + rowoff - plane, zmf, false);
            if (zp_on) load(u + rowoff + plane, zpf, false);
        }
        int det = -1;
        if (valid[0]) det = P.seg_flag[((size_t)zl * S + r) * P.ntab + tab0];
        const bool all_plain = __all_sync(0xffffffffu, det < 0);
        double RS[4] = {0, 0, 0, 0}, RC[4] = {0, 0, 0, 0};
        if (STEP) {
            ld256_nc(P.rowinfo + ((size_t)zl * S + r) * 8, RS);
            ld256_nc(P.rowinfo + ((size_t)zl * S + r) * 8 + 4, RC);
        }
        float un[CPL * 4];
        double rsum[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            double pun[4];
            if (MODE == 1) {
                unsigned f = fplain;
                if (det >= 0) f = P.det_flags[(size_t)det * kTabSeg + ((c0 + j) & (kTabSeg - 1))];
                if (!valid[j]) f = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) pun[q] = ((f >> q) & 1u) ? (double)uf[j * 4 + q] : 0.0;
            } else if (all_plain) {
                const double cell = (double)((uf[j * 4] + uf[j * 4 + 1]) + (uf[j * 4 + 2] + uf[j * 4 + 3]));
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float x = uf[j * 4 + q], xp = upf[j * 4 + q];
                    const double zz = K3D ? (double)(zmf[j * 4 + q] + zpf[j * 4 + q]) : 0.0;
                    const double nb = (cell + RS[q]) + (CS[j][q] + zz);
                    const double dg = (RC[q] + degbase) + CC[j][q];
                    const float lap = (float)(nb - dg * (double)x);
                    un[j * 4 + q] = 2.0f * x - xp + coeff_f * lap - damp_f * (x - xp);
                    pun[q] = valid[j] ? (double)un[j * 4 + q] : 0.0;
                }
            } else {
                unsigned f = fplain;
                if (det >= 0) f = P.det_flags[(size_t)det * kTabSeg + ((c0 + j) & (kTabSeg - 1))];
                if (!valid[j]) f = 0;
                CellIO io;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    io.uq[q] = uf[j * 4 + q]; io.upq[q] = upf[j * 4 + q];
                    io.zm[q] = zmf[j * 4 + q]; io.zp[q] = zpf[j * 4 + q];
                    io.RS[q] = RS[q]; io.RC[q] = RC[q]; io.CS[q] = CS[j][q]; io.CC[q] = CC[j][q];
                }
                cell_update_detail_f32(P, u, det, (c0 + j) & (kTabSeg - 1), (long long)(rowoff + j * 4), f,
                                       coeff, damping, io);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    un[j * 4 + q] = (float)io.un[q];
                    pun[q] = ((f >> q) & 1u) ? (double)un[j * 4 + q] : 0.0;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) { cacc[j][q] += pun[q]; rsum[q] += pun[q]; }
        }
        if (STEP) {
            if (CPL == 2 && both && wide_ok) {
                st256f_cs(w + rowoff, un);
                if (K3D) {
                    if (zl == 0 && peer_lo) st256f(peer_lo + (rowoff - zbase), un);
                    if (zl == P.Lloc - 1 && peer_hi) st256f(peer_hi + (rowoff - zbase), un);
                }
            } else {
#pragma unroll
                for (int j = 0; j < CPL; ++j)
                    if (valid[j]) {
                        const float4 v = make_float4(un[j * 4], un[j * 4 + 1], un[j * 4 + 2], un[j * 4 + 3]);
                        st128_cs(w + rowoff + j * 4, v);
                        if (K3D) {
                            if (zl == 0 && peer_lo) st128(peer_lo + (rowoff - zbase) + j * 4, v);
                            if (zl == P.Lloc - 1 && peer_hi) st128(peer_hi + (rowoff - zbase) + j * 4, v);
                        }
                    }
            }
        }
        warp_sum4_store(rsum, lane, prow + (((size_t)zl * S + r) * nsegk + segk) * 4);
    }
#pragma unroll
    for (int j = 0; j < CPL; ++j)
        if (valid[j]) st256(pcol + (((size_t)zl * nchunk + ch) * S + c0 + j) * 4, cacc[j]);
}

// 3-D step with explicit on-chip sharing of the z neighbours (cp.async multi-stage pipeline).
//
// In 3-D every node also reads u(z-1) and u(z+1): 40 B/update through L2 if left to the caches
// (measured: 61 % of the HBM roofline).  Here a CTA owns WZ consecutive layers x 32 cells and
// marches down the rows; each row iteration stages the WZ (+2 halo) row segments of u_t and the WZ
// segments of u_{t-1} in shared memory with 16-byte LDGSTS copies, STAGES deep, one barrier per row.
// Warp wz then reads its own values from slot wz+1 and its z-neighbours from slots wz and wz+2, so
// each u value enters the SM once per CTA: (WZ+2)/WZ * 8 + 8 + 8 B/update of L2->SM traffic (26 B
// for WZ = 8), and HBM traffic stays ~24 B because the two halo rows are the rows the neighbouring
// z-group CTA streams at the same time (z-groups are the fastest block index).
// Slots of layers that do not exist are zero-filled (src-size 0), matching deg via degbase/flags.
template <int WZ, int STAGES, int MINB>
__global__ void __launch_bounds__(WZ * 32, MINB)
lat_step3d_kernel(const __grid_constant__ LatDev P, const double *__restrict__ u,
                  double *__restrict__ w, double *__restrict__ prow, double *__restrict__ pcol,
                  double coeff, double damping, int zl_begin, int zl_end, int rchunk, int nchunk,
                  int nsegk, int nzg)
{
    extern __shared__ __align__(16) unsigned char tk_smem[];
    // sU[stage][slot 0..WZ+1][half][lane], sW[stage][wz][half][lane]   (double2 = 16 B each)
    // sR[stage][wz][4] = the row's {RowSum(4), rowcnt(4)}, sF[stage][wz] = its defect-table entry: the
    // small per-row loads are staged too, otherwise they queue behind the bulk prefetch traffic in the
    // in-order L1 pipeline and every row pays a full loaded-memory latency after the barrier.
    double2 *sU = reinterpret_cast<double2 *>(tk_smem);
    double2 *sW = sU + STAGES * (WZ + 2) * 64;
    double2 *sR = sW + STAGES * WZ * 64;
    int *sF = reinterpret_cast<int *>(sR + STAGES * WZ * 4);

    const int lane = threadIdx.x & 31, wz = threadIdx.x >> 5;
    int b = blockIdx.x;
    const int zg = b % nzg;
    b /= nzg;
    const int segk = b % nsegk;
    const int ch = b / nsegk;
    const int zl = zl_begin + zg * WZ + wz;
    const bool active = zl < zl_end;  // warp-uniform; inactive warps only keep the barriers
    const bool last_active = active && (wz == WZ - 1 || zl + 1 >= zl_end);

    const int S = P.S;
    const int c = segk * 32 + lane;
    const bool valid = active && c < S;
    const size_t plane = (size_t)S * S * 4;
    const size_t zbase = (size_t)(zl + P.halo) * plane;
    const unsigned fplain = active ? plain_flags(P, zl) : 0u;
    const bool zm_on = (fplain & 0x100u) != 0, zp_on = (fplain & 0x1000u) != 0;
    const double degbase = 4.0 + (zm_on ? 1.0 : 0.0) + (zp_on ? 1.0 : 0.0);

    double CS[4] = {0, 0, 0, 0}, CC[4] = {0, 0, 0, 0}, cacc[4] = {0, 0, 0, 0};
    if (valid) {
        ld256_nc(P.colinfo + ((size_t)zl * S + c) * 8, CS);
        ld256_nc(P.colinfo + ((size_t)zl * S + c) * 8 + 4, CC);
    }
    const int r_begin = ch * rchunk;
    const int r_end = min(S, r_begin + rchunk);

    auto issue = [&](int r, int s) {
        if (active) {
            const double *gu = u + zbase + ((size_t)r * S + c) * 4;
            const double *gw = w + zbase + ((size_t)r * S + c) * 4;
            double2 *du = sU + ((s * (WZ + 2) + wz + 1) * 2) * 32 + lane;
            double2 *dw = sW + ((s * WZ + wz) * 2) * 32 + lane;
            cp_async16(du, valid ? gu : u, valid);
            cp_async16(du + 32, valid ? gu + 2 : u, valid);
            cp_async16(dw, valid ? gw : u, valid);
            cp_async16(dw + 32, valid ? gw + 2 : u, valid);
            if (lane < 4)
                cp_async16(sR + (s * WZ + wz) * 4 + lane, P.rowinfo + ((size_t)zl * S + r) * 8 + lane * 2, true);
            if (lane == 4)
                cp_async4(sF + s * WZ + wz, P.seg_flag + ((size_t)zl * S + r) * P.ntab + (segk * 32) / kTabSeg);
            if (wz == 0) {  // halo row below the CTA's first layer
                const bool ok = valid && zm_on;
                double2 *dh = sU + ((s * (WZ + 2)) * 2) * 32 + lane;
                cp_async16(dh, ok ? gu - plane : u, ok);
                cp_async16(dh + 32, ok ? gu - plane + 2 : u, ok);
            }
            if (last_active) {  // halo row above the CTA's last active layer
                const bool ok = valid && zp_on;
                double2 *dh = sU + ((s * (WZ + 2) + wz + 2) * 2) * 32 + lane;
                cp_async16(dh, ok ? gu + plane : u, ok);
                cp_async16(dh + 32, ok ? gu + plane + 2 : u, ok);
            }
        }
        cp_async_commit();
    };

#pragma unroll
    for (int i = 0; i < STAGES - 1; ++i) {
        if (r_begin + i < r_end) issue(r_begin + i, i);
        else cp_async_commit();
    }
    int s = 0;
    for (int r = r_begin; r < r_end; ++r) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int sn = s + STAGES - 1;
            if (sn >= STAGES) sn -= STAGES;
            if (r + STAGES - 1 < r_end) issue(r + STAGES - 1, sn);
            else cp_async_commit();
        }
        if (active) {
            double uq[4], upq[4], zm[4], zp[4];
            {
                const double2 *pu = sU + ((s * (WZ + 2) + wz + 1) * 2) * 32 + lane;
                const double2 a = pu[0], bb = pu[32];
                uq[0] = a.x; uq[1] = a.y; uq[2] = bb.x; uq[3] = bb.y;
                const double2 *pm = sU + ((s * (WZ + 2) + wz) * 2) * 32 + lane;
                const double2 m0 = pm[0], m1 = pm[32];
                zm[0] = m0.x; zm[1] = m0.y; zm[2] = m1.x; zm[3] = m1.y;
                const double2 *pp = sU + ((s * (WZ + 2) + wz + 2) * 2) * 32 + lane;
                const double2 p0 = pp[0], p1 = pp[32];
                zp[0] = p0.x; zp[1] = p0.y; zp[2] = p1.x; zp[3] = p1.y;
                const double2 *pw = sW + ((s * WZ + wz) * 2) * 32 + lane;
                const double2 w0 = pw[0], w1 = pw[32];
                upq[0] = w0.x; upq[1] = w0.y; upq[2] = w1.x; upq[3] = w1.y;
            }
            const int det = valid ? sF[s * WZ + wz] : -1;
            const bool all_plain = sF[s * WZ + wz] < 0;  // warp-uniform (one table entry per 64 cells)
            double RS[4], RC[4];
            {
                const double2 *pr = sR + (s * WZ + wz) * 4;
                const double2 r0 = pr[0], r1 = pr[1], r2 = pr[2], r3 = pr[3];
                RS[0] = r0.x; RS[1] = r0.y; RS[2] = r1.x; RS[3] = r1.y;
                RC[0] = r2.x; RC[1] = r2.y; RC[2] = r3.x; RC[3] = r3.y;
            }
            double un[4], pun[4];
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
                cell_update_detail(P, u, det, c & (kTabSeg - 1),
                                   (long long)(zbase + ((size_t)r * S + c) * 4), f, coeff, damping, 0.0, io);
#pragma unroll
                for (int q = 0; q < 4; ++q) { un[q] = io.un[q]; pun[q] = io.pun[q]; }
            }
            if (valid) st256_cs(w + zbase + ((size_t)r * S + c) * 4, un);
#pragma unroll
            for (int q = 0; q < 4; ++q) cacc[q] += pun[q];
            warp_sum4_store(pun, lane, prow + (((size_t)zl * S + r) * nsegk + segk) * 4);
        }
        if (++s == STAGES) s = 0;
    }
    cp_async_wait<0>();
    if (valid) st256(pcol + (((size_t)zl * nchunk + ch) * S + c) * 4, cacc);
}

// Deterministic finalise: fixed-order sums of the partials written by lat_step_kernel, plus the
// probe gather (reference physics.py:162 reads one node per recorded state).
//   blocks [0, nrowblk)   : one WARP per (layer,row) line.  The line's partials are nsegk*4
//                           contiguous doubles; lane l strides by 32 so it always meets quadrant
//                           l&3, then lanes of equal quadrant are combined by a fixed butterfly.
//   blocks [nrowblk, ...) : one THREAD per (layer, column, quadrant), summing over row chunks
//                           (coalesced across columns).
//   block 0 also gathers the probes from the NEW state.
__global__ void __launch_bounds__(256)
lat_finalise_kernel(int S, int Lloc, int nsegk, int nchunk, int nrowblk,
                    const double *__restrict__ prow, const double *__restrict__ pcol,
                    double *__restrict__ rowinfo, double *__restrict__ colinfo,
                    const void *__restrict__ state, int state_is_f32,
                    const long long *__restrict__ probe_idx, int n_probes, double *__restrict__ probe_row,
                    const double *__restrict__ pen, long long n_pen, double *__restrict__ energy_out,
                    const long long *__restrict__ ctr)
{
    // ctr != null (CUDA-graph replay): probe_row / energy_out are the BASES of the series and the row
    // to write comes from device counters (ctr[0] probes, ctr[2] energy), advanced by advance_kernel
    if (ctr != nullptr) {
        probe_row += (size_t)ctr[0] * n_probes;
        if (energy_out != nullptr) energy_out += ctr[2];
    }
    if (blockIdx.x == 0)
        for (int t = threadIdx.x; t < n_probes; t += blockDim.x)
            probe_row[t] = state_is_f32 ? (double)static_cast<const float *>(state)[probe_idx[t]]
                                        : static_cast<const double *>(state)[probe_idx[t]];
    if (blockIdx.x == gridDim.x - 1 && energy_out != nullptr) {
        // last CTA: staggered energy E_{t+1/2} = 1/2 * sum of the per-warp partials, fixed order
        __shared__ double red[256];
        double a = 0.0;
        for (long long i = threadIdx.x; i < n_pen; i += 256) a += pen[i];
        red[threadIdx.x] = a;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) *energy_out = 0.5 * red[0];
        return;
    }
    const long long nline = (long long)Lloc * S;
    if ((int)blockIdx.x < nrowblk) {
        const int lane = threadIdx.x & 31;
        const long long line = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
        if (line >= nline) return;
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
        double acc = (a0 + a1) + (a2 + a3);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        if (lane < 4) rowinfo[line * 8 + lane] = acc;
    } else {
        // 32 (column, quadrant) values x 8 chunk slices per CTA; slices combined in fixed order
        __shared__ double part[8][32];
        const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
        const long long t = (long long)(blockIdx.x - nrowblk) * 32 + lane;
        double acc = 0.0;
        if (t < nline * 4) {
            const long long line = t >> 2;  // zl*S + c
            const int q = (int)(t & 3);
            const long long zl = line / S, c = line % S;
            const double *p = pcol + ((size_t)zl * nchunk * S + c) * 4 + q;
            const size_t stride = (size_t)S * 4;
            double a0 = 0.0, a1 = 0.0;
            int k = slice;
            for (; k + 8 < nchunk; k += 16) {
                a0 += p[(size_t)k * stride];
                a1 += p[(size_t)(k + 8) * stride];
            }
            if (k < nchunk) a0 += p[(size_t)k * stride];
            acc = a0 + a1;
        }
        part[slice][lane] = acc;
        __syncthreads();
        if (slice == 0 && t < nline * 4) {
            const double s01 = part[0][lane] + part[1][lane], s23 = part[2][lane] + part[3][lane];
            const double s45 = part[4][lane] + part[5][lane], s67 = part[6][lane] + part[7][lane];
            colinfo[(t >> 2) * 8 + (t & 3)] = (s01 + s23) + (s45 + s67);
        }
    }
}

// Arrival flags of the halo exchange as kernels (one thread each), the default since round 2: the stream-ordered
// driver memory operations (cuStreamWriteValue64 / cuStreamWaitValue64) dead-locked under random host-side delays
// between the ranks (tests/test_gpu_multi.py::test_halo_flag_protocol_survives_random_delays: both GPUs parked in
// their waits with every operation enqueued) although they never did in lock-step runs.  The wait polls this rank's
// own flag words with acquire loads -- the neighbour runs on ANOTHER GPU, so nothing on this device has to make
// progress for the flag to arrive -- and traps after ~20 s instead of hanging the device.
__global__ void flag_signal_kernel(unsigned long long *f0, unsigned long long *f1, unsigned long long value)
{
    __threadfence_system();  // everything this stream did before (the pushes) is visible before the flag moves
    if (f0) asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(f0), "l"(value) : "memory");
    if (f1) asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(f1), "l"(value) : "memory");
}

__global__ void flag_wait_kernel(const unsigned long long *f0, const unsigned long long *f1, unsigned long long value)
{
    const long long t0 = clock64();
    for (int side = 0; side < 2; ++side) {
        const unsigned long long *f = side ? f1 : f0;
        if (!f) continue;
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= value) break;
            __nanosleep(100);
            if (clock64() - t0 > 40000000000LL) __trap();  // ~20 s at 2 GHz: fail loudly, never hang the device
        }
    }
}

// max over existing nodes of the graph degree (reference physics.py:60-64), from the tables.
__global__ void __launch_bounds__(256)
lat_maxdeg_kernel(const __grid_constant__ LatDev P, int *__restrict__ out)
{
    const long long ncell = (long long)P.Lloc * P.S * P.S;
    int best = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ncell;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % P.S);
        const long long zr = i / P.S;
        const int r = (int)(zr % P.S), zl = (int)(zr / P.S);
        const int det = P.seg_flag[(size_t)zr * P.ntab + c / kTabSeg];
        unsigned f = plain_flags(P, zl);
        int2 cr = make_int2(0, 0);
        if (det >= 0) {
            f = P.det_flags[(size_t)det * kTabSeg + (c & (kTabSeg - 1))];
            cr = P.det_corr[det];
        }
        const int ccnt = __popc(f & 0xFu);
        const long long node0 = (((long long)(zl + P.halo) * P.S + r) * P.S + c) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (!((f >> (4 + q)) & 1u)) continue;
            int deg = 0;
            if ((f >> q) & 1u) {
                deg = (ccnt - 1) + ((int)P.rowinfo[((size_t)zl * P.S + r) * 8 + 4 + q] - 1) +
                      ((int)P.colinfo[((size_t)zl * P.S + c) * 8 + 4 + q] - 1) +
                      (int)((f >> (8 + q)) & 1u) + (int)((f >> (12 + q)) & 1u);
            }
            for (int k = cr.x; k < cr.y; ++k)
                if (P.corr_a[k] == node0 + q) deg += (int)P.corr_sign[k];
            best = max(best, deg);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best > 0) atomicMax(out, best);
}

__global__ void scatter_kernel(double *__restrict__ dst, const long long *__restrict__ idx,
                               const double *__restrict__ val, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[idx[i]] = val[i];
}

__global__ void scatter_f32_kernel(float *__restrict__ dst, const long long *__restrict__ idx,
                                   const double *__restrict__ val, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[idx[i]] = (float)val[i];
}

// fp32 state <-> fp64 staging (set_state / get_state / history of an fp32 plan)
__global__ void f32_to_f64_kernel(const float *__restrict__ src, double *__restrict__ dst, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = (double)src[i];
}

__global__ void f64_to_f32_kernel(const double *__restrict__ src, float *__restrict__ dst, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = (float)src[i];
}

__global__ void gather_kernel(const double *__restrict__ src, const long long *__restrict__ idx,
                              double *__restrict__ out, int n, const long long *__restrict__ ctr)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (ctr != nullptr) out += (size_t)ctr[0] * n;  // graph replay: row from the device counter
    if (i < n) out[i] = src[idx[i]];
}

// One state row into the history staging buffer at the row given by a device counter (graph replay).
__global__ void history_row_kernel(const void *__restrict__ src, int src_is_f32, double *__restrict__ hist,
                                   long long n, const long long *__restrict__ ctr)
{
    double *dst = hist + (size_t)ctr[1] * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = src_is_f32 ? (double)static_cast<const float *>(src)[i] : static_cast<const double *>(src)[i];
}

// End of a replayed step: advance the row counters (probes, history, energy).
__global__ void advance_kernel(long long *ctr, int inc)
{
    if (threadIdx.x < 3) ctr[threadIdx.x] += inc;
}

}  // namespace tk
