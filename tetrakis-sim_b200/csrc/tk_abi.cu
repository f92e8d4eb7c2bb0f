// C ABI of the tetrakis-b200 hot path (see include/tetrakis_b200.h for the contract).
// Host side: plan construction (graph -> sliced ELL; lattice descriptor -> defect tables),
// the stepping loops and all CUDA resource management.  No CPU compute fallback anywhere.
#include "../../include/tetrakis_b200.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "kernels_ensemble.cuh"
#include "kernels_fused.cuh"
#include "kernels_graph.cuh"
#include "kernels_lattice.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define TK_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(e_ == cudaErrorMemoryAllocation ? TK_ERR_NOMEM : TK_ERR_CUDA,          \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define TK_REQUIRE(cond, ...)                                    \
    do {                                                         \
        if (!(cond)) return fail(TK_ERR_INVALID, __VA_ARGS__);   \
    } while (0)

int check_device(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(TK_ERR_NO_DEVICE,
                    "no CUDA device available (%s); tetrakis-b200 has no CPU path",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(TK_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    TK_CUDA(cudaSetDevice(device));
    return TK_OK;
}

template <typename T>
int dev_alloc(T **p, size_t count)
{
    *p = nullptr;
    if (count == 0) count = 1;
    TK_CUDA(cudaMalloc((void **)p, count * sizeof(T)));
    return TK_OK;
}

template <typename T>
int dev_upload(T **p, const std::vector<T> &v)
{
    int rc = dev_alloc(p, v.size());
    if (rc) return rc;
    if (!v.empty()) TK_CUDA(cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return TK_OK;
}

enum PlanKind { kGraph = 1, kLattice = 2, kSheetBand = 3 };

constexpr size_t kHistoryChunkBytes = (size_t)1 << 30;

}  // namespace

struct tk_plan {
    int kind = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    long long n = 0;          // local nodes visible through the ABI
    long long alloc = 0;      // doubles per state buffer
    long long state_off = 0;  // ABI index 0 lives at this offset in a state buffer
    double *buf[2] = {nullptr, nullptr};  // raw state storage; fp32 plans reinterpret it as float
    int esz = 8;                          // bytes per state element: 8 (fp64) or 4 (optional fp32 mode)
    int cur = 0;  // buf[cur] = u_t, buf[cur^1] = u_{t-1} (overwritten by u_{t+1})
    long long launches = 0;
    long long max_degree = -1;
    std::vector<void *> owned;  // device allocations to free

    // probes
    long long *probe_idx = nullptr;  // device, buffer-relative indices
    int n_probes = 0;
    double *probe_series = nullptr;  // device, rows x n_probes
    long long probe_rows_cap = 0;
    long long probe_row = 0;

    // graph
    long long *slice_off = nullptr;
    int *ell_idx = nullptr;
    double *deg = nullptr, *inv_mass = nullptr, *potential = nullptr;

    // lattice
    tk::LatDev lat{};
    double *rowinfo = nullptr, *colinfo = nullptr, *prow = nullptr, *pcol = nullptr;
    int cpl = 2, wc = 8, rchunk = 64, nchunk = 1, nsegk = 1;
    int bulk_stages = 0;                     // > 0: fp64 steps use lat_step_bulk_kernel with this many stages
    int wz3d = 0, stages3d = 0, minb3d = 2;  // wz3d > 0: 3-D steps use lat_step3d_kernel<wz3d, stages3d, minb3d>
    int pf_rows = 0;                         // register path: L2 prefetch distance in rows (0 = off)
    // energy series (optional): per-warp / per-CTA partials + one value per step
    double *pen = nullptr;
    long long pen_cap = 0, pen_used = 0;
    double *energy_series = nullptr;
    long long energy_cap = 0, energy_row = 0;
    bool energy_on = false;
    cudaStream_t launch_stream = nullptr;    // stream of the next layer launch (defaults to `stream`)
    // fused halo push: neighbours' state buffers (peer-mapped) and flag words
    double *peer_buf[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [side][buffer]
    int peer_layers[2] = {0, 0};
    unsigned long long *flags = nullptr;     // this rank's arrival flags: [0] from below, [16] from above
    bool push_in_kernel = true;              // step kernel stores boundary planes into the peers' halos
    // CUDA-graph replay of the stepping loop: output rows come from device counters
    long long *ctr_dev = nullptr;            // [0] probe row, [1] history row, [2] energy row
    bool ctr_mode = false;
    int minb = 1;                            // register path: min resident CTAs per SM asked of ptxas
    int layers_global = 1, z_begin = 0;
    double coeff = 0.0, damping = 0.0;
    // K-step fused passes (defect-free 2-D sheets, kernels_fused.cuh)
    bool plain2d = false;                    // eligible: one layer, no defect records, fp64
    int fuse_k = 0, fuse_minb = 3;           // largest K the response tables were sized for (0 = not set up)
    int fuse_k_run = 0;                      // steps per pass of the last tk_plan_run
    // which finalised sums are current (plain 2-D sheets only; lets a run skip the sums-only passes)
    bool sums_cur_valid = false;             // rowinfo/colinfo hold the sums of buf[cur]
    int sums_prev = 0;                       // rowinfo_prev/colinfo_prev: 0 unknown, 1 valid, 2 buf[cur^1] is all zero
    // rows this plan holds: S for a whole sheet, fewer for one row band of a sheet sharded over GPUs
    // (kSheetBand: column sums and totals are partial and travel through an all-reduced exchange buffer)
    int nrows = 0, row_begin = 0;
    double *ftot_global = nullptr;           // kSheetBand: all-reduced quadrant totals, 2 levels
    std::vector<cudaEvent_t> band_ev;        // kSheetBand: ring of event pairs around the pass kernels
    long long band_ev_next = 0, band_ev_read = 0;
    int fuse_wc = 8, fuse_rchunk = 0, fuse_nchunk = 0, fuse_nsegk = 0, fuse_occ = 0;
    double *tabR = nullptr, *tabC = nullptr, *rowinfo_prev = nullptr, *colinfo_prev = nullptr;
    double *prow_b = nullptr, *fpcol_a = nullptr, *fpcol_b = nullptr, *ftot = nullptr;
    std::vector<long long> probe_idx_host;   // buffer-relative probe indices (fused passes write probes)
    long long fused_steps_last = 0;          // steps the last tk_plan_run advanced with fused passes
    std::vector<long long> dead_local;  // ABI-local indices of removed nodes (for dense set_state)
};

namespace {

inline bool is_f32(const tk_plan *p) { return p->esz == 4; }
// address of element `off` of state buffer b (buffers are typed by p->esz)
inline char *bptr(const tk_plan *p, int b, long long off) { return reinterpret_cast<char *>(p->buf[b]) + (size_t)off * p->esz; }
inline char *pptr(const tk_plan *p, int side, int b, long long off)
{
    return reinterpret_cast<char *>(p->peer_buf[side][b]) + (size_t)off * p->esz;
}

int plan_common_init(tk_plan *p, int device)
{
    p->device = device;
    TK_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    p->own_stream = true;
    cudaDeviceProp prop;
    TK_CUDA(cudaGetDeviceProperties(&prop, device));
    p->sm_count = prop.multiProcessorCount;
    return TK_OK;
}

template <typename T>
int plan_upload(tk_plan *p, T **dst, const std::vector<T> &v)
{
    int rc = dev_upload(dst, v);
    if (rc == TK_OK) p->owned.push_back((void *)*dst);
    return rc;
}

template <typename T>
int plan_alloc(tk_plan *p, T **dst, size_t count, bool zero)
{
    int rc = dev_alloc(dst, count);
    if (rc) return rc;
    p->owned.push_back((void *)*dst);
    if (zero) TK_CUDA(cudaMemset(*dst, 0, std::max<size_t>(count, 1) * sizeof(T)));
    return TK_OK;
}

int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// ------------------------------------------------------------------ lattice launches
constexpr size_t lat_bulk_smem(bool k3d, int stages)
{
    return (size_t)8 * stages * ((k3d ? 4 : 2) * 128 + 16) * sizeof(double) + (size_t)8 * stages * 8;
}

#define TK_BULK_STAGES(X) X(2) X(3) X(4) X(6)

template <int WC, bool K3D, int MINB>
int launch_lat_bulk(tk_plan *p, long long grid, cudaStream_t st, int zl_begin, int zl_end, int nsegk, int nct, int nzg)
{
#define X(ST)                                                                                               \
    if (p->bulk_stages == ST) {                                                                             \
        auto kern = tk::lat_step_bulk_kernel<WC, K3D, ST, MINB>;                                            \
        TK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,                     \
                                     (int)lat_bulk_smem(K3D, ST)));                                         \
        kern<<<(unsigned)grid, 256, lat_bulk_smem(K3D, ST), st>>>(                                          \
            p->lat, p->buf[p->cur], p->buf[p->cur ^ 1], p->prow, p->pcol, p->coeff, p->damping, zl_begin,   \
            zl_end, p->rchunk, p->nchunk, nsegk, nct, nzg);                                                 \
        p->launches++;                                                                                      \
        TK_CUDA(cudaGetLastError());                                                                        \
        return TK_OK;                                                                                       \
    }
    TK_BULK_STAGES(X)
#undef X
    return fail(TK_ERR_INVALID, "unsupported TK_LAT_BULK stage count %d (2, 3, 4 or 6)", p->bulk_stages);
}

template <int WC, bool K3D, int MINB>
int lat_bulk_occ(int stages)
{
    int n = 0;
#define X(ST)                                                                                               \
    if (stages == ST) {                                                                                     \
        auto kern = tk::lat_step_bulk_kernel<WC, K3D, ST, MINB>;                                            \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lat_bulk_smem(K3D, ST)); \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 256, lat_bulk_smem(K3D, ST)) != cudaSuccess) { \
            cudaGetLastError();                                                                             \
            n = 1;                                                                                          \
        }                                                                                                   \
    }
    TK_BULK_STAGES(X)
#undef X
    return std::max(n, 1);
}

template <int CPL, int WC, bool K3D, int MODE, int MINB = 1>
int launch_lat(tk_plan *p, int zl_begin, int zl_end)
{
    constexpr int WZ = 8 / WC;
    const int nsegk = p->nsegk;
    const int nct = (nsegk + WC - 1) / WC;
    const int nzg = (zl_end - zl_begin + WZ - 1) / WZ;
    const long long grid = (long long)p->nchunk * nct * nzg;
    if (grid <= 0) return TK_OK;
    if (grid > 0x7fffffffLL) return fail(TK_ERR_INVALID, "lattice grid too large");
    double *pen = nullptr;
    if (MODE == 2) {
        if (p->pen_used + grid * 8 > p->pen_cap) return fail(TK_ERR_INVALID, "energy partial buffer overflow");
        pen = p->pen + p->pen_used;
        p->pen_used += grid * 8;
    }
    cudaStream_t st = p->launch_stream ? p->launch_stream : p->stream;
    if (K3D && MODE != 1) {
        const long long plane = (long long)p->lat.S * p->lat.S * 4;
        const int nb = p->cur ^ 1;  // every rank swaps in lockstep, so buffer parities agree
        p->lat.peer_lo = (p->push_in_kernel && p->peer_buf[0][nb])
                             ? reinterpret_cast<double *>(pptr(p, 0, nb, (p->peer_layers[0] + 1) * plane)) : nullptr;
        p->lat.peer_hi = (p->push_in_kernel && p->peer_buf[1][nb]) ? p->peer_buf[1][nb] : nullptr;
    }
    if constexpr (MODE == 0 && CPL == 1 && ((WC == 8 && (MINB == 3 || (K3D && MINB == 2))) || (WC == 1 && K3D && MINB == 2))) {
        if (p->bulk_stages && !is_f32(p))
            return launch_lat_bulk<WC, K3D, MINB>(p, grid, st, zl_begin, zl_end, nsegk, nct, nzg);
    }
    if (is_f32(p)) {
        if (CPL > 2 || MODE == 2)
            return fail(TK_ERR_INVALID, "fp32 plans support 1 or 2 cells per lane and no energy series");
        constexpr int FC = CPL <= 2 ? CPL : 1;
        constexpr int FM = MODE != 2 ? MODE : 0;
        constexpr int FB = (MINB >= 1 && MINB <= 4) ? MINB : 3;
        tk::lat_step_f32_kernel<FC, WC, K3D, FM, FB><<<(unsigned)grid, 256, 0, st>>>(
            p->lat, reinterpret_cast<const float *>(p->buf[p->cur]), reinterpret_cast<float *>(p->buf[p->cur ^ 1]),
            p->prow, p->pcol, p->coeff, p->damping, zl_begin, zl_end, p->rchunk, p->nchunk, nsegk, nct, nzg,
            p->pf_rows);
    } else {
        tk::lat_step_kernel<CPL, WC, K3D, MODE, MINB><<<(unsigned)grid, 256, 0, st>>>(
            p->lat, p->buf[p->cur], p->buf[p->cur ^ 1], p->prow, p->pcol, p->coeff, p->damping, zl_begin,
            zl_end, p->rchunk, p->nchunk, nsegk, nct, nzg, p->pf_rows, pen,
            p->coeff > 0.0 ? 1.0 / p->coeff : 0.0);
    }
    p->launches++;
    TK_CUDA(cudaGetLastError());
    return TK_OK;
}

// (cells per lane, segments per CTA, min resident CTAs per SM asked of ptxas)
#define TK_LAT_VARIANTS(X)                                                                        \
    X(1, 8, 2) X(1, 8, 3) X(1, 8, 4) X(1, 1, 2) X(1, 1, 3) X(1, 1, 4) X(1, 2, 2) X(1, 2, 3) \
    X(2, 8, 1) X(2, 8, 2) X(2, 1, 1) X(2, 1, 2) X(2, 2, 1) X(4, 8, 1) X(4, 1, 1)

bool lat_supported(int cpl, int wc, int mb)
{
#define X(CPL, WC, MB) if (cpl == CPL && wc == WC && mb == MB) return true;
    TK_LAT_VARIANTS(X)
#undef X
    return false;
}

template <int MODE>
int launch_lat_dispatch(tk_plan *p, int zl_begin, int zl_end)
{
    const bool k3d = p->lat.halo != 0;
    if (MODE == 2) {  // step + energy partials: instantiated for the default shapes only
        if (p->cpl == 1 && p->wc == 8)
            return k3d ? launch_lat<1, 8, true, 2, 2>(p, zl_begin, zl_end) : launch_lat<1, 8, false, 2, 3>(p, zl_begin, zl_end);
        if (p->cpl == 1 && p->wc == 1)
            return k3d ? launch_lat<1, 1, true, 2, 2>(p, zl_begin, zl_end) : launch_lat<1, 1, false, 2, 3>(p, zl_begin, zl_end);
        return fail(TK_ERR_INVALID, "the energy series needs the default kernel tiling (cpl=1, wc=1 or 8)");
    }
    if (MODE != 0) {  // sums-only pass: register pressure is irrelevant, one instantiation per shape
#define X(CPL, WC, MB)                                                        \
    if (p->cpl == CPL && p->wc == WC && p->minb == MB)                        \
        return k3d ? launch_lat<CPL, WC, true, 1, 1>(p, zl_begin, zl_end)     \
                   : launch_lat<CPL, WC, false, 1, 1>(p, zl_begin, zl_end);
        TK_LAT_VARIANTS(X)
#undef X
    } else {
#define X(CPL, WC, MB)                                                        \
    if (p->cpl == CPL && p->wc == WC && p->minb == MB)                        \
        return k3d ? launch_lat<CPL, WC, true, 0, MB>(p, zl_begin, zl_end)    \
                   : launch_lat<CPL, WC, false, 0, MB>(p, zl_begin, zl_end);
        TK_LAT_VARIANTS(X)
#undef X
    }
    return fail(TK_ERR_INVALID, "unsupported lattice kernel variant cpl=%d wc=%d minb=%d", p->cpl, p->wc,
                p->minb);
}

template <int WZ, int STAGES>
constexpr size_t lat3d_smem()
{
    return (size_t)STAGES * ((2 * WZ + 2) * 1024 + WZ * 64 + ((WZ * 4 + 15) / 16) * 16);
}

template <int WZ, int STAGES, int MINB>
int launch_lat3d(tk_plan *p, int zl_begin, int zl_end)
{
    auto kern = tk::lat_step3d_kernel<WZ, STAGES, MINB>;
    // per device and cheap: set on every launch rather than caching a per-process flag
    TK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)lat3d_smem<WZ, STAGES>()));
    const int nzg = (zl_end - zl_begin + WZ - 1) / WZ;
    const long long grid = (long long)p->nchunk * p->nsegk * nzg;
    if (grid <= 0) return TK_OK;
    if (grid > 0x7fffffffLL) return fail(TK_ERR_INVALID, "lattice grid too large");
    kern<<<(unsigned)grid, WZ * 32, lat3d_smem<WZ, STAGES>(), p->stream>>>(
        p->lat, p->buf[p->cur], p->buf[p->cur ^ 1], p->prow, p->pcol, p->coeff, p->damping, zl_begin,
        zl_end, p->rchunk, p->nchunk, p->nsegk, nzg);
    p->launches++;
    TK_CUDA(cudaGetLastError());
    return TK_OK;
}

template <int WZ, int STAGES, int MINB>
int lat3d_occ_one()
{
    int n = 0;
    auto kern = tk::lat_step3d_kernel<WZ, STAGES, MINB>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lat3d_smem<WZ, STAGES>());
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, WZ * 32, lat3d_smem<WZ, STAGES>()) != cudaSuccess) {
        cudaGetLastError();
        n = 1;
    }
    return std::max(n, 1);
}

// (layers per CTA, pipeline stages, min resident CTAs per SM)
#define TK_3D_VARIANTS(X) X(4, 4, 4) X(8, 3, 2) X(8, 3, 3) X(8, 4, 2) X(8, 4, 3) X(8, 6, 2) X(16, 3, 1) X(16, 4, 1)

bool lat3d_supported(int wz, int st, int mb)
{
#define X(WZ, ST, MB) if (wz == WZ && st == ST && mb == MB) return true;
    TK_3D_VARIANTS(X)
#undef X
    return false;
}

int lat3d_occupancy(int wz, int st, int mb)
{
#define X(WZ, ST, MB) if (wz == WZ && st == ST && mb == MB) return lat3d_occ_one<WZ, ST, MB>();
    TK_3D_VARIANTS(X)
#undef X
    return 1;
}

int launch_lat3d_dispatch(tk_plan *p, int zl_begin, int zl_end)
{
#define X(WZ, ST, MB) \
    if (p->wz3d == WZ && p->stages3d == ST && p->minb3d == MB) return launch_lat3d<WZ, ST, MB>(p, zl_begin, zl_end);
    TK_3D_VARIANTS(X)
#undef X
    return fail(TK_ERR_INVALID, "unsupported 3-D kernel variant wz=%d stages=%d minb=%d", p->wz3d,
                p->stages3d, p->minb3d);
}

template <int CPL, int WC, bool K3D, int MB>
int lat_occ_one()
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, tk::lat_step_kernel<CPL, WC, K3D, 0, MB>, 256, 0) != cudaSuccess) {
        cudaGetLastError();
        n = 2;
    }
    return std::max(n, 1);
}

// resident CTAs per SM of the step kernel variant in use
int lat_occupancy(int cpl, int wc, int mb, bool k3d)
{
#define X(CPL, WC, MB) \
    if (cpl == CPL && wc == WC && mb == MB) return k3d ? lat_occ_one<CPL, WC, true, MB>() : lat_occ_one<CPL, WC, false, MB>();
    TK_LAT_VARIANTS(X)
#undef X
    return 2;
}

int lat_finalise_sums(tk_plan *p, const double *prow, const double *pcol, int nsegk, int nchunk, double *rowinfo,
                      double *colinfo)
{
    const long long nline = (long long)p->lat.Lloc * p->lat.S;
    const int nrowblk = (int)((nline + 7) / 8);
    const int ncolblk = (int)((nline * 4 + 31) / 32);
    tk::lat_finalise_kernel<<<(unsigned)(nrowblk + ncolblk), 256, 0, p->stream>>>(
        p->lat.S, p->lat.Lloc, nsegk, nchunk, nrowblk, prow, pcol, rowinfo, colinfo, nullptr, 0, nullptr, 0,
        nullptr, nullptr, 0, nullptr, nullptr);
    p->launches++;
    TK_CUDA(cudaGetLastError());
    return TK_OK;
}

int lat_finalise(tk_plan *p, const double *state, double *probe_row)
{
    const long long nline = (long long)p->lat.Lloc * p->lat.S;
    const int nrowblk = (int)((nline + 7) / 8);
    const int ncolblk = (int)((nline * 4 + 31) / 32);
    double *eout = nullptr;
    if (p->ctr_mode) {
        if (p->energy_on && p->pen_used > 0) eout = p->energy_series;  // base; the row comes from ctr[2]
    } else if (p->energy_on && p->pen_used > 0 && p->energy_row < p->energy_cap) {
        eout = p->energy_series + p->energy_row++;
    }
    tk::lat_finalise_kernel<<<(unsigned)(nrowblk + ncolblk + (eout ? 1 : 0)), 256, 0, p->stream>>>(
        p->lat.S, p->lat.Lloc, p->nsegk, p->nchunk, nrowblk, p->prow, p->pcol, p->rowinfo, p->colinfo,
        state, is_f32(p) ? 1 : 0, p->probe_idx, probe_row ? p->n_probes : 0, probe_row, p->pen, p->pen_used,
        eout, p->ctr_mode ? p->ctr_dev : nullptr);
    p->pen_used = 0;
    p->launches++;
    TK_CUDA(cudaGetLastError());
    return TK_OK;
}

int set_energy(tk_plan *p, bool on, long long steps)
{
    p->energy_on = false;
    p->energy_row = 0;
    p->pen_used = 0;
    if (p->energy_series) { cudaFree(p->energy_series); p->energy_series = nullptr; }
    if (!on) return TK_OK;
    long long need;
    if (p->kind == kGraph) {
        need = (p->n + 255) / 256;
    } else {
        TK_REQUIRE(p->wz3d == 0, "the energy series is not available with TK_LAT3D=1");
        const int wz = 8 / p->wc;
        need = (long long)p->nchunk * ((p->nsegk + p->wc - 1) / p->wc) * ((p->lat.Lloc + wz - 1) / wz + 2) * 8;
    }
    if (need > p->pen_cap) {
        if (p->pen) cudaFree(p->pen);
        p->pen = nullptr;
        TK_CUDA(cudaMalloc((void **)&p->pen, std::max<long long>(need, 1) * sizeof(double)));
        p->pen_cap = need;
    }
    TK_CUDA(cudaMalloc((void **)&p->energy_series, std::max<long long>(steps, 1) * sizeof(double)));
    p->energy_cap = steps;
    p->energy_on = true;
    return TK_OK;
}

double *next_probe_row(tk_plan *p)
{
    if (p->ctr_mode) return p->n_probes > 0 ? p->probe_series : nullptr;  // base; row from ctr[0]
    if (p->n_probes == 0 || p->probe_row >= p->probe_rows_cap) return nullptr;
    return p->probe_series + (size_t)(p->probe_row++) * p->n_probes;
}

int set_probes(tk_plan *p, long long n_probes, const int64_t *idx, long long rows)
{
    p->n_probes = 0;
    p->probe_row = 0;
    p->probe_idx_host.clear();
    if (p->probe_idx) { cudaFree(p->probe_idx); p->probe_idx = nullptr; }
    if (p->probe_series) { cudaFree(p->probe_series); p->probe_series = nullptr; }
    if (n_probes <= 0) return TK_OK;
    std::vector<long long> h(n_probes);
    for (long long i = 0; i < n_probes; ++i) {
        TK_REQUIRE(idx[i] >= 0 && idx[i] < p->n, "probe index %lld out of range [0,%lld)",
                   (long long)idx[i], p->n);
        h[i] = idx[i] + p->state_off;
    }
    TK_CUDA(cudaMalloc((void **)&p->probe_idx, n_probes * sizeof(long long)));
    TK_CUDA(cudaMemcpy(p->probe_idx, h.data(), n_probes * sizeof(long long), cudaMemcpyHostToDevice));
    TK_CUDA(cudaMalloc((void **)&p->probe_series, (size_t)rows * n_probes * sizeof(double)));
    p->n_probes = (int)n_probes;
    p->probe_rows_cap = rows;
    p->probe_idx_host = h;
    return TK_OK;
}

// ------------------------------------------------------------------ K-step fused passes (2-D, no defects)
#define TK_CHK(x)          \
    do {                   \
        const int rc_ = (x); \
        if (rc_) return rc_; \
    } while (0)

size_t fuse_smem(int wc, int rows) { return ((size_t)wc * 32 + rows) * 8 * sizeof(double); }

template <int MB>
int fuse_occ_one(int wc, int rows)
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, tk::lat_fused2d_kernel<MB, false>, wc * 32, fuse_smem(wc, rows)) !=
        cudaSuccess) {
        cudaGetLastError();
        n = 1;
    }
    return std::max(n, 1);
}

int fuse_occupancy(int mb, int wc, int rows)
{
    return mb == 2 ? fuse_occ_one<2>(wc, rows) : (mb == 4 ? fuse_occ_one<4>(wc, rows) : fuse_occ_one<3>(wc, rows));
}

// Steps per pass for this run (0 = step by step).  TK_LAT_FUSE: 0/1 off, else steps per pass.
int fuse_choice(const tk_plan *p, const tk_run_args *a)
{
    if (p->kind != kLattice || !p->plain2d || a->history_out || a->energy_out) return 0;
    if (a->n_probes > tk::kFusedMaxProbes) return 0;
    // default: the largest K at which the pass is still HBM-bound on a B200 (FP64 pipe ~60 % busy): the cell
    // recurrences cost 4 FP64 instructions per cell and step without damping, 8 with
    const int kmax = env_int("TK_LAT_FUSE", a->damping == 0.0 ? 48 : 24);
    if (kmax < 2 || a->steps < 2) return 0;
    // passes of (nearly) equal length: every pass moves the whole state, so 200 steps are 5 x 40, not 4 x 48 + 8
    const int64_t npass = (a->steps + kmax - 1) / kmax;
    int64_t k = (a->steps + npass - 1) / npass;
    // the pass kernel's step loop is unrolled by 8 and its remainder is slow (K = 54 costs what K = 64 does):
    // round the pass length up to a multiple of 8 where K allows
    if (k > 8 && (k + 7) / 8 * 8 <= kmax) k = (k + 7) / 8 * 8;
    return (int)std::min<int64_t>(k, a->steps);
}

int fuse_setup(tk_plan *p, int k)
{
    if (p->fuse_k >= k) return TK_OK;  // tables are sized for the largest K seen; shorter passes reuse them
    k = std::max(k, 64);               // (cudaMalloc is slow: do not grow them run by run)
    const int S = p->lat.S, R = p->nrows;
    int mb = env_int("TK_LAT_FUSE_MINB", 4);
    if (mb < 2 || mb > 4) mb = 3;
    const int nseg32 = (S + 31) / 32;
    const int wc = nseg32 >= 8 ? 8 : (nseg32 >= 2 ? 2 : 1);
    const int nct = (nseg32 + wc - 1) / wc;
    // rows per CTA: the same trade as the step kernel -- per-CTA response/partial traffic against whole
    // waves of (SMs x resident CTAs); two partial-sum sets are written per pass
    const int rc_cap = 128;
    const int occ = fuse_occupancy(mb, wc, rc_cap);
    const long long slots = (long long)p->sm_count * occ;
    int best = 16;
    double best_eff = -1.0;
    if ((long long)((R + 15) / 16) * nct < slots) {
        const long long want = std::max<long long>(1, slots / std::max(1, nct));
        best = (int)std::max<long long>(4, std::min<long long>((R + want - 1) / want, 16));
    } else {
        for (int rc_ = 16; rc_ <= rc_cap && rc_ <= std::max(16, R); ++rc_) {
            const long long tiles = (long long)((R + rc_ - 1) / rc_) * nct;
            const long long waves = (tiles + slots - 1) / slots;
            const double eff = (double)tiles / (double)(waves * slots) * (1.0 / (1.0 + 2.0 / rc_));
            if (eff > best_eff + 1e-9) { best_eff = eff; best = rc_; }
        }
    }
    int rchunk = env_int("TK_LAT_FUSE_RCHUNK", best);
    if (rchunk < 1 || rchunk > rc_cap) rchunk = best;
    const int nchunk = (R + rchunk - 1) / rchunk;
    TK_CHK(plan_alloc(p, &p->tabR, (size_t)R * k * 4, true));
    TK_CHK(plan_alloc(p, &p->tabC, (size_t)S * k * 4, true));
    if (!p->rowinfo_prev) {
        TK_CHK(plan_alloc(p, &p->rowinfo_prev, (size_t)R * 8, true));
        TK_CHK(plan_alloc(p, &p->colinfo_prev, (size_t)S * 8, true));
        TK_CHK(plan_alloc(p, &p->prow_b, (size_t)R * nseg32 * 4, true));
        TK_CHK(plan_alloc(p, &p->ftot, (size_t)2 * ((R + 7) / 8) * 4, true));  // per-block quadrant totals, 2 levels
        TK_CHK(plan_alloc(p, &p->fpcol_a, (size_t)nchunk * S * 4, true));
        TK_CHK(plan_alloc(p, &p->fpcol_b, (size_t)nchunk * S * 4, true));
        p->fuse_minb = mb; p->fuse_wc = wc; p->fuse_rchunk = rchunk; p->fuse_nchunk = nchunk;
        p->fuse_nsegk = nseg32; p->fuse_occ = occ;
    }
    p->fuse_k = k;
    return TK_OK;
}

// Row/column sums of u_{t-1} (the recurrences need two time levels); the sums of u_t are current.
int fuse_prepare(tk_plan *p)
{
    if (p->sums_prev == 2) {
        tk::lat_zero_sums_kernel<<<(unsigned)((2 * p->lat.S * 4 + 255) / 256), 256, 0, p->stream>>>(
            p->lat.S, p->rowinfo_prev, p->colinfo_prev);
        p->launches++;
    } else if (p->sums_prev != 1) {
        p->cur ^= 1;
        const int rc = launch_lat_dispatch<1>(p, 0, p->lat.Lloc);
        p->cur ^= 1;
        if (rc) return rc;
        TK_CHK(lat_finalise_sums(p, p->prow, p->pcol, p->nsegk, p->nchunk, p->rowinfo_prev, p->colinfo_prev));
    }
    p->sums_prev = 1;
    const int nblk = (p->nrows + 7) / 8;
    tk::lat_totparts_kernel<<<(unsigned)((2 * nblk * 4 + 255) / 256), 256, 0, p->stream>>>(
        p->nrows, nblk, p->rowinfo, p->rowinfo_prev, p->ftot);
    p->launches++;
    TK_CUDA(cudaGetLastError());
    return TK_OK;
}

// One pass = k steps.  buf[cur] stays u_t's buffer (it receives u_{t+k}), buf[cur^1] receives u_{t+k-1}.
int fuse_pass(tk_plan *p, int k, cudaEvent_t ev_begin, cudaEvent_t ev_end)
{
    const int S = p->lat.S, wc = p->fuse_wc;
    const int nct = (p->fuse_nsegk + wc - 1) / wc;
    tk::FusedProbes pr{};
    pr.n = p->n_probes;
    for (int i = 0; i < p->n_probes; ++i) {
        pr.cell[i] = p->probe_idx_host[i] >> 2;
        pr.q[i] = (int)(p->probe_idx_host[i] & 3);
    }
    const double g = p->damping;
    const double ca = (2.0 - g) - p->coeff * (2.0 * S + 4.0), cb = -(1.0 - g);
    const int R = p->nrows;
    const int nrowblk = (R + 7) / 8, ncolblk = (S * 4 + 31) / 32;
    if (k > 0)
        tk::lat_tables_kernel<<<(unsigned)((R + S + 127) / 128), 128, 0, p->stream>>>(
            S, R, k, p->coeff, g, ca, cb, p->rowinfo, p->rowinfo_prev, p->colinfo, p->colinfo_prev, p->ftot, nrowblk,
            p->kind == kSheetBand ? p->ftot_global : nullptr, p->tabR, p->tabC);
    const unsigned grid = (unsigned)(p->fuse_nchunk * nct);
    const size_t sm = fuse_smem(wc, p->fuse_rchunk);
    if (ev_begin) TK_CUDA(cudaEventRecord(ev_begin, p->stream));
#define TK_FUSE_LAUNCH1(MB, UD)                                                                                 \
    tk::lat_fused2d_kernel<MB, UD><<<grid, wc * 32, sm, p->stream>>>(                                           \
        S, R, k, p->buf[p->cur], p->buf[p->cur ^ 1], p->tabR, p->tabC, p->prow, p->fpcol_a, p->prow_b, p->fpcol_b, ca, cb, \
        p->coeff, p->fuse_rchunk, p->fuse_nchunk, p->fuse_nsegk, nct, p->pf_rows, pr, p->probe_series,          \
        p->ctr_mode ? p->ctr_dev : nullptr, p->probe_row)
#define TK_FUSE_LAUNCH(MB)              \
    do {                                \
        if (g == 0.0) TK_FUSE_LAUNCH1(MB, true);  \
        else TK_FUSE_LAUNCH1(MB, false);          \
    } while (0)
    if (p->fuse_minb == 2) TK_FUSE_LAUNCH(2);
    else if (p->fuse_minb == 4) TK_FUSE_LAUNCH(4);
    else TK_FUSE_LAUNCH(3);
#undef TK_FUSE_LAUNCH
#undef TK_FUSE_LAUNCH1
    if (ev_end) TK_CUDA(cudaEventRecord(ev_end, p->stream));
    tk::lat_finalise2_kernel<<<(unsigned)(2 * (nrowblk + ncolblk)), 256, 0, p->stream>>>(
        S, R, p->fuse_nsegk, p->fuse_nchunk, nrowblk, ncolblk, p->prow, p->fpcol_a, p->prow_b, p->fpcol_b, p->rowinfo,
        p->colinfo, p->rowinfo_prev, p->colinfo_prev, p->ftot);
    p->launches += 3;
    TK_CUDA(cudaGetLastError());
    if (!p->ctr_mode && p->n_probes > 0) p->probe_row += k;
    p->sums_cur_valid = true;
    p->sums_prev = 1;
    return TK_OK;
}

int graph_step(tk_plan *p, double coeff, double damping)
{
    const unsigned grid = (unsigned)((p->n + 255) / 256);
    if (grid) {
        const double inv_coeff = coeff > 0.0 ? 1.0 / coeff : 0.0;
        if (p->energy_on && (p->ctr_mode || p->energy_row < p->energy_cap)) {
            tk::graph_step_kernel<true><<<grid, 256, 0, p->stream>>>(
                p->n, p->slice_off, p->ell_idx, p->deg, p->inv_mass, p->potential, p->buf[p->cur],
                p->buf[p->cur ^ 1], coeff, damping, p->pen, inv_coeff);
            if (p->ctr_mode)
                tk::energy_reduce_kernel<<<1, 256, 0, p->stream>>>(p->pen, grid, p->energy_series, p->ctr_dev);
            else
                tk::energy_reduce_kernel<<<1, 256, 0, p->stream>>>(p->pen, grid, p->energy_series + p->energy_row++, nullptr);
            p->launches += 2;
        } else {
            tk::graph_step_kernel<false><<<grid, 256, 0, p->stream>>>(
                p->n, p->slice_off, p->ell_idx, p->deg, p->inv_mass, p->potential, p->buf[p->cur],
                p->buf[p->cur ^ 1], coeff, damping, nullptr, inv_coeff);
            p->launches++;
        }
        TK_CUDA(cudaGetLastError());
    }
    p->cur ^= 1;
    if (double *row = next_probe_row(p)) {
        tk::gather_kernel<<<(p->n_probes + 127) / 128, 128, 0, p->stream>>>(
            p->buf[p->cur], p->probe_idx, row, p->n_probes, p->ctr_mode ? p->ctr_dev : nullptr);
        p->launches++;
        TK_CUDA(cudaGetLastError());
    }
    return TK_OK;
}

// fp32 plans keep the ABI in double: dense state moves through a bounded fp64 staging buffer
constexpr long long kStageElems = 1LL << 24;

int stage_f64_to_f32(tk_plan *p, const double *host, int b)
{
    double *d = nullptr;
    TK_CUDA(cudaMalloc((void **)&d, (size_t)std::min<long long>(p->n, kStageElems) * sizeof(double)));
    cudaError_t e = cudaSuccess;
    for (long long o = 0; o < p->n && e == cudaSuccess; o += kStageElems) {
        const long long m = std::min<long long>(kStageElems, p->n - o);
        e = cudaMemcpyAsync(d, host + o, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, p->stream);
        if (e != cudaSuccess) break;
        tk::f64_to_f32_kernel<<<1024, 256, 0, p->stream>>>(d, reinterpret_cast<float *>(bptr(p, b, p->state_off + o)), m);
        p->launches++;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    }
    cudaFree(d);
    TK_CUDA(e);
    return TK_OK;
}

int stage_f32_to_f64(tk_plan *p, int b, double *host)
{
    double *d = nullptr;
    TK_CUDA(cudaMalloc((void **)&d, (size_t)std::min<long long>(p->n, kStageElems) * sizeof(double)));
    cudaError_t e = cudaSuccess;
    for (long long o = 0; o < p->n && e == cudaSuccess; o += kStageElems) {
        const long long m = std::min<long long>(kStageElems, p->n - o);
        tk::f32_to_f64_kernel<<<1024, 256, 0, p->stream>>>(reinterpret_cast<const float *>(bptr(p, b, p->state_off + o)), d, m);
        p->launches++;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(host + o, d, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, p->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    }
    cudaFree(d);
    TK_CUDA(e);
    return TK_OK;
}

}  // namespace

// ====================================================================================== ABI
extern "C" {

int tk_abi_version(void) { return TK_ABI_VERSION; }

const char *tk_last_error(void) { return g_err.c_str(); }

int tk_device_count(int *count)
{
    TK_REQUIRE(count, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    return TK_OK;
}

int tk_device_info(int device, char *name, int name_len, int *sm_count, int *cc_major, int *cc_minor,
                   int64_t *total_mem_bytes)
{
    int rc = check_device(device);
    if (rc) return rc;
    cudaDeviceProp prop;
    TK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (name && name_len > 0) snprintf(name, name_len, "%s", prop.name);
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (total_mem_bytes) *total_mem_bytes = (int64_t)prop.totalGlobalMem;
    return TK_OK;
}

// ------------------------------------------------------------------------------ graph plan
int tk_graph_plan_create(int device, int64_t n, const int64_t *rowptr, const int32_t *colidx,
                         const double *degree, const double *inv_mass, const double *potential,
                         tk_plan **out)
{
    TK_REQUIRE(out, "out is NULL");
    *out = nullptr;
    TK_REQUIRE(n >= 0 && n < 0x7fffffffLL, "n_nodes %lld out of range", (long long)n);
    TK_REQUIRE(n == 0 || (rowptr && degree && inv_mass && potential), "NULL graph array");
    int rc = check_device(device);
    if (rc) return rc;
    tk_plan *p = new tk_plan();
    p->kind = kGraph;
    p->n = n;
    p->alloc = n + 1;  // +1: the always-zero slot that ELL padding points at
#define TK_TRY(x)          \
    do {                   \
        rc = (x);          \
        if (rc) {          \
            tk_plan_destroy(p); \
            return rc;     \
        }                  \
    } while (0)
    TK_TRY(plan_common_init(p, device));
    const long long nslice = (n + 31) / 32;
    std::vector<long long> soff(nslice + 1, 0);
    long long maxdeg = 0;
    for (long long s = 0; s < nslice; ++s) {
        long long wmax = 0;
        for (long long i = s * 32; i < std::min<long long>(n, s * 32 + 32); ++i) {
            const long long wd = rowptr[i + 1] - rowptr[i];
            if (wd < 0) { tk_plan_destroy(p); return fail(TK_ERR_INVALID, "rowptr not monotone at %lld", i); }
            wmax = std::max(wmax, wd);
        }
        soff[s + 1] = soff[s] + wmax * 32;
    }
    for (long long i = 0; i < n; ++i) maxdeg = std::max<long long>(maxdeg, (long long)degree[i]);
    p->max_degree = maxdeg;
    std::vector<int> ell((size_t)soff[nslice], (int)n);
    for (long long i = 0; i < n; ++i) {
        const long long s = i >> 5, lane = i & 31;
        for (long long k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            if (colidx[k] < 0 || colidx[k] >= n) {
                tk_plan_destroy(p);
                return fail(TK_ERR_INVALID, "colidx[%lld]=%d out of range", k, colidx[k]);
            }
            ell[(size_t)(soff[s] + (k - rowptr[i]) * 32 + lane)] = colidx[k];
        }
    }
    TK_TRY(plan_upload(p, &p->slice_off, soff));
    TK_TRY(plan_upload(p, &p->ell_idx, ell));
    TK_TRY(plan_upload(p, &p->deg, std::vector<double>(degree, degree + n)));
    TK_TRY(plan_upload(p, &p->inv_mass, std::vector<double>(inv_mass, inv_mass + n)));
    TK_TRY(plan_upload(p, &p->potential, std::vector<double>(potential, potential + n)));
    TK_TRY(plan_alloc(p, &p->buf[0], (size_t)p->alloc, true));
    TK_TRY(plan_alloc(p, &p->buf[1], (size_t)p->alloc, true));
    *out = p;
    return TK_OK;
}

// ------------------------------------------------------------------------------ lattice plan
int tk_lattice_plan_create(int device, const tk_lattice_desc *d, tk_plan **out)
{
    return tk_lattice_plan_create_ex(device, d, 0, out);
}

int tk_lattice_plan_create_ex(int device, const tk_lattice_desc *d, uint32_t flags, tk_plan **out)
{
    TK_REQUIRE(out, "out is NULL");
    TK_REQUIRE((flags & ~(uint32_t)TK_PLAN_F32) == 0, "unknown plan flags 0x%x", flags);
    *out = nullptr;
    TK_REQUIRE(d, "desc is NULL");
    TK_REQUIRE(d->size >= 1 && d->layers >= 1, "size and layers must be >= 1");
    TK_REQUIRE(0 <= d->z_begin && d->z_begin < d->z_end && d->z_end <= d->layers,
               "slab [%d,%d) not inside [0,%d)", d->z_begin, d->z_end, d->layers);
    TK_REQUIRE(d->n_special == 0 || (d->special_index && d->special_flags), "special arrays NULL");
    TK_REQUIRE(d->n_corr == 0 || (d->corr_a && d->corr_b && d->corr_sign), "corr arrays NULL");
    int rc = check_device(device);
    if (rc) return rc;

    const long long S = d->size, L = d->layers, zb = d->z_begin, ze = d->z_end, Lloc = ze - zb;
    const long long halo = L > 1 ? 1 : 0;
    const long long plane = S * S * 4, planes = Lloc + 2 * halo;
    const int ntab = (int)((S + tk::kTabSeg - 1) / tk::kTabSeg);

    tk_plan *p = new tk_plan();
    p->kind = kLattice;
    p->esz = (flags & TK_PLAN_F32) ? 4 : 8;
    p->n = Lloc * plane;
    p->alloc = planes * plane;
    p->state_off = halo * plane;
    p->layers_global = (int)L;
    p->z_begin = (int)zb;
    TK_TRY(plan_common_init(p, device));

    // ---- sparse defect description -> per-cell bits
    struct CellBits { unsigned char part = 0xF, alive = 0xF; };
    std::unordered_map<long long, CellBits> cells;            // key: global cell index (z*S+r)*S+c
    std::unordered_map<long long, std::pair<double, double>> attrs;  // key: global node index
    for (long long i = 0; i < d->n_special; ++i) {
        const long long g = d->special_index[i];
        if (g < 0 || g >= L * plane) { tk_plan_destroy(p); return fail(TK_ERR_INVALID, "special_index[%lld]=%lld out of range", i, g); }
        const long long z = g / plane;
        if (z < zb - 1 || z > ze) continue;  // irrelevant to this slab
        const int q = (int)(g & 3);
        const unsigned char fl = d->special_flags[i];
        const bool alive = fl & 1, part = (fl & 2) && alive;
        if (!alive || !part) {
            CellBits &cb = cells[g >> 2];
            if (!alive) cb.alive &= ~(1u << q);
            if (!part) cb.part &= ~(1u << q);
        }
        const double im = d->special_inv_mass ? d->special_inv_mass[i] : 1.0;
        const double v = d->special_potential ? d->special_potential[i] : 0.0;
        if (alive && (im != 1.0 || v != 0.0) && z >= zb && z < ze) attrs[g] = {im, v};
    }
    auto cell_bits = [&](long long z, long long r, long long c) -> CellBits {
        auto it = cells.find((z * S + r) * S + c);
        return it == cells.end() ? CellBits() : it->second;
    };

    // ---- participation counts (static) and removed-node list
    std::vector<double> rowinfo((size_t)Lloc * S * 8, 0.0), colinfo((size_t)Lloc * S * 8, 0.0);
    for (long long zl = 0; zl < Lloc; ++zl)
        for (long long x = 0; x < S; ++x)
            for (int q = 0; q < 4; ++q) {
                rowinfo[(size_t)(zl * S + x) * 8 + 4 + q] = (double)S;
                colinfo[(size_t)(zl * S + x) * 8 + 4 + q] = (double)S;
            }
    for (const auto &kv : cells) {
        const long long cellg = kv.first, z = cellg / (S * S);
        if (z < zb || z >= ze) continue;
        const long long r = (cellg / S) % S, c = cellg % S, zl = z - zb;
        for (int q = 0; q < 4; ++q) {
            if (!((kv.second.part >> q) & 1)) {
                rowinfo[(size_t)(zl * S + r) * 8 + 4 + q] -= 1.0;
                colinfo[(size_t)(zl * S + c) * 8 + 4 + q] -= 1.0;
            }
            if (!((kv.second.alive >> q) & 1)) p->dead_local.push_back(((zl * S + r) * S + c) * 4 + q);
        }
    }
    std::sort(p->dead_local.begin(), p->dead_local.end());

    // ---- corrections: global -> padded-local, sorted by receiving node
    struct Corr { long long a, b; double s; };
    std::vector<Corr> corr;
    auto to_padded = [&](long long g, long long *outp) -> bool {
        const long long z = g / plane;
        if (z < zb - 1 || z > ze || (halo == 0 && z != 0)) return false;
        *outp = g - zb * plane + halo * plane;
        return true;
    };
    for (long long i = 0; i < d->n_corr; ++i) {
        const long long a = d->corr_a[i], b = d->corr_b[i];
        if (a < 0 || a >= L * plane || b < 0 || b >= L * plane) { tk_plan_destroy(p); return fail(TK_ERR_INVALID, "corr index out of range"); }
        const long long za = a / plane;
        if (za < zb || za >= ze) continue;
        long long pa, pb;
        if (!to_padded(a, &pa) || !to_padded(b, &pb)) {
            tk_plan_destroy(p);
            return fail(TK_ERR_INVALID, "edge correction %lld->%lld reaches beyond the slab halo", a, b);
        }
        corr.push_back({pa, pb, d->corr_sign[i]});
    }
    std::sort(corr.begin(), corr.end(), [](const Corr &x, const Corr &y) { return x.a < y.a; });

    // ---- flagged table segments
    auto segkey = [&](long long zl, long long r, long long c) { return (zl * S + r) * ntab + c / tk::kTabSeg; };
    std::vector<long long> keys;
    for (const auto &kv : cells) {
        const long long cellg = kv.first, z = cellg / (S * S), r = (cellg / S) % S, c = cellg % S;
        for (long long dz = -1; dz <= 1; ++dz) {
            const long long zl = z + dz - zb;
            if (zl >= 0 && zl < Lloc) keys.push_back(segkey(zl, r, c));
        }
    }
    for (const auto &kv : attrs) {
        const long long cellg = kv.first >> 2;
        keys.push_back(segkey(cellg / (S * S) - zb, (cellg / S) % S, cellg % S));
    }
    for (const Corr &c : corr) {
        const long long cellp = c.a >> 2;  // padded-local cell
        keys.push_back(segkey(cellp / (S * S) - halo, (cellp / S) % S, cellp % S));
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    const size_t ndet = keys.size();

    std::vector<int> seg_flag((size_t)Lloc * S * ntab, -1);
    std::vector<unsigned short> det_flags(ndet * tk::kTabSeg, 0);
    std::vector<int> det_attr(ndet, -1);
    std::vector<int2> det_corr(ndet, make_int2(0, 0));
    std::vector<double> attr_im, attr_v;
    size_t ci = 0;
    for (size_t k = 0; k < ndet; ++k) {
        const long long key = keys[k];
        seg_flag[(size_t)key] = (int)k;
        const long long tab = key % ntab, zr = key / ntab, r = zr % S, zl = zr / S, z = zl + zb;
        bool any_attr = false;
        for (int cc = 0; cc < tk::kTabSeg; ++cc) {
            const long long c = tab * tk::kTabSeg + cc;
            if (c >= S) continue;
            const CellBits me = cell_bits(z, r, c);
            unsigned f = (unsigned)me.part | ((unsigned)me.alive << 4);
            if (z - 1 >= 0) f |= (unsigned)cell_bits(z - 1, r, c).part << 8;
            if (z + 1 < L) f |= (unsigned)cell_bits(z + 1, r, c).part << 12;
            det_flags[k * tk::kTabSeg + cc] = (unsigned short)f;
            for (int q = 0; q < 4 && !any_attr; ++q)
                if (attrs.count((((z * S + r) * S + c) << 2) + q)) any_attr = true;
        }
        if (any_attr) {
            det_attr[k] = (int)(attr_im.size() / 256);
            attr_im.resize(attr_im.size() + 256, 1.0);
            attr_v.resize(attr_v.size() + 256, 0.0);
            for (int cc = 0; cc < tk::kTabSeg && tab * tk::kTabSeg + cc < S; ++cc)
                for (int q = 0; q < 4; ++q) {
                    auto it = attrs.find(((((z * S + r) * S + tab * tk::kTabSeg + cc)) << 2) + q);
                    if (it != attrs.end()) {
                        attr_im[attr_im.size() - 256 + cc * 4 + q] = it->second.first;
                        attr_v[attr_v.size() - 256 + cc * 4 + q] = it->second.second;
                    }
                }
        }
        // corrections whose receiving node lies in this segment (corr is sorted by a, and the
        // padded node order follows (zl, r, c), i.e. the key order)
        const long long lo = ((((zl + halo) * S + r) * S) + tab * tk::kTabSeg) * 4;
        const long long hi = lo + tk::kTabSeg * 4;
        while (ci < corr.size() && corr[ci].a < lo) ++ci;
        size_t cj = ci;
        while (cj < corr.size() && corr[cj].a < hi) ++cj;
        det_corr[k] = make_int2((int)ci, (int)cj);
    }
    std::vector<long long> ca(corr.size()), cb(corr.size());
    std::vector<double> cs(corr.size());
    for (size_t i = 0; i < corr.size(); ++i) { ca[i] = corr[i].a; cb[i] = corr[i].b; cs[i] = corr[i].s; }

    // ---- kernel variant + tiling
    p->cpl = env_int("TK_LAT_CPL", 1);
    // 2-D: 8 adjacent segments per CTA -- unless the rows are too short to have 8 (idle warps otherwise)
    const int nseg32 = (int)((S + 31) / 32);
    p->wc = env_int("TK_LAT_WC", halo ? 1 : (nseg32 >= 8 ? 8 : (nseg32 >= 2 ? 2 : 1)));
    // defaults from the round-1 sweeps on B200 (profiles/r01_sweeps.md): 2-D 3 CTAs/SM + 1-row L2
    // prefetch; 3-D 2 CTAs/SM + 2-row prefetch
    p->minb = env_int("TK_LAT_MINB", p->cpl == 1 ? (halo ? 2 : 3) : 1);
    p->pf_rows = env_int("TK_LAT_PREFETCH", halo ? 2 : 1);
    if (!lat_supported(p->cpl, p->wc, p->minb)) { p->cpl = 1; p->wc = halo ? 1 : 8; p->minb = 3; }
    if (p->esz == 4) {  // fp32 (sweep, profiles/r01_sweeps.md): 2-D 2 cells per lane, 3-D 1 cell per lane
        p->cpl = env_int("TK_LAT_CPL", halo ? 1 : 2);
        if (p->cpl != 1 && p->cpl != 2) p->cpl = halo ? 1 : 2;
        p->minb = env_int("TK_LAT_MINB", p->cpl == 1 ? 3 : 2);
        if (!lat_supported(p->cpl, p->wc, p->minb)) { p->cpl = 1; p->wc = halo ? 1 : 8; p->minb = 3; }
    }
    if (halo && p->esz == 8 && env_int("TK_LAT3D", 0)) {  // 3-D: shared-memory staged kernel (1 cell per lane)
        p->wz3d = env_int("TK_LAT3D_WZ", 8);
        p->stages3d = env_int("TK_LAT3D_STAGES", 4);
        p->minb3d = env_int("TK_LAT3D_MINB", p->wz3d == 16 ? 1 : (p->wz3d == 4 ? 4 : 2));
        if (!lat3d_supported(p->wz3d, p->stages3d, p->minb3d)) { p->wz3d = 8; p->stages3d = 4; p->minb3d = 2; }
        p->cpl = 1;
        p->wc = 1;
    }
    p->plain2d = halo == 0 && Lloc == 1 && ndet == 0 && corr.empty() && p->esz == 8;
    p->nrows = (int)S;
    if (p->esz == 8 && !p->wz3d && env_int("TK_LAT_BULK", 0)) {  // TMA-staged variant of the default shapes
        p->bulk_stages = env_int("TK_LAT_BULK_STAGES", halo ? 3 : 4);
        if (p->bulk_stages != 2 && p->bulk_stages != 3 && p->bulk_stages != 4 && p->bulk_stages != 6)
            p->bulk_stages = halo ? 3 : 4;
        p->cpl = 1;
        p->wc = halo ? 1 : 8;
        p->minb = halo ? 2 : 3;
    }
    const int nsegk_max = (int)((S + 31) / 32);
    p->nsegk = (int)((S + 32 * p->cpl - 1) / (32 * p->cpl));
    {
        // rows per CTA: trade the per-CTA column traffic (colinfo read + pcol write = 96 B per
        // cell per chunk, i.e. a fraction 1/rchunk of the streaming bytes) against wave
        // quantisation on (SMs x resident CTAs) slots
        const int wz = p->wz3d ? p->wz3d : 8 / p->wc;
        const long long nct = p->wz3d ? p->nsegk : (p->nsegk + p->wc - 1) / p->wc;
        const long long nzg = (Lloc + wz - 1) / wz;
        const long long slots =
            (long long)p->sm_count *
            (p->wz3d ? lat3d_occupancy(p->wz3d, p->stages3d, p->minb3d)
                     : p->bulk_stages ? (halo ? lat_bulk_occ<1, true, 2>(p->bulk_stages) : lat_bulk_occ<8, false, 3>(p->bulk_stages))
                                      : lat_occupancy(p->cpl, p->wc, p->minb, halo != 0));
        int best = 16;
        double best_eff = -1.0;
        // Small lattices never fill the machine: a step is then bound by the serial march of one warp
        // down its rows (~1 us per row out of L2), not by bandwidth, so use the SHORTEST march that still
        // gives about one CTA per slot (measured: 512^2 went from 20 to ~8 us per step).
        const long long tiles16 = ((S + 15) / 16) * nct * nzg;
        if (tiles16 < slots) {
            const long long per_row_block = std::max<long long>(1, nct * nzg);
            const long long want_chunks = std::max<long long>(1, slots / per_row_block);
            long long rc_small = (S + want_chunks - 1) / want_chunks;
            rc_small = std::max<long long>(4, std::min<long long>(rc_small, 16));
            best = (int)rc_small;
            best_eff = 2.0;  // skip the throughput search below
        }
        // (cap at 128 rows: longer marches measured slower -- cfg4 512^3: 256 rows 226 GNUPS, 64-128
        // rows 244 -- CTAs of neighbouring z-groups drift apart and lose the shared z+-1 lines in L2)
        for (int rc_ = 16; rc_ <= 128 && rc_ <= std::max<long long>(16, S); ++rc_) {
            const long long tiles = ((S + rc_ - 1) / rc_) * nct * nzg;
            const long long waves = (tiles + slots - 1) / slots;
            const double eff = (double)tiles / (double)(waves * slots) * (1.0 / (1.0 + 1.0 / rc_));
            if (eff > best_eff + 1e-9) { best_eff = eff; best = rc_; }
        }
        p->rchunk = env_int("TK_LAT_RCHUNK", best);
        if (p->rchunk < 1) p->rchunk = best;
    }
    p->nchunk = (int)((S + p->rchunk - 1) / p->rchunk);
    const int nchunk_alloc = p->nchunk;

    const int *d_seg_flag; const unsigned short *d_det_flags; const int *d_det_attr; const int2 *d_det_corr;
    const double *d_attr_im, *d_attr_v; const long long *d_ca, *d_cb; const double *d_cs;
    seg_flag.resize(seg_flag.size() + 4, -1);  // the TMA variant fetches the aligned 16-byte quad around an entry
    TK_TRY(plan_upload(p, const_cast<int **>(&d_seg_flag), seg_flag));
    TK_TRY(plan_upload(p, const_cast<unsigned short **>(&d_det_flags), det_flags));
    TK_TRY(plan_upload(p, const_cast<int **>(&d_det_attr), det_attr));
    TK_TRY(plan_upload(p, const_cast<int2 **>(&d_det_corr), det_corr));
    TK_TRY(plan_upload(p, const_cast<double **>(&d_attr_im), attr_im));
    TK_TRY(plan_upload(p, const_cast<double **>(&d_attr_v), attr_v));
    TK_TRY(plan_upload(p, const_cast<long long **>(&d_ca), ca));
    TK_TRY(plan_upload(p, const_cast<long long **>(&d_cb), cb));
    TK_TRY(plan_upload(p, const_cast<double **>(&d_cs), cs));
    TK_TRY(plan_upload(p, &p->rowinfo, rowinfo));
    TK_TRY(plan_upload(p, &p->colinfo, colinfo));
    TK_TRY(plan_alloc(p, &p->prow, (size_t)Lloc * S * nsegk_max * 4, true));
    TK_TRY(plan_alloc(p, &p->pcol, (size_t)Lloc * nchunk_alloc * S * 4, true));
    {   // state buffers are typed by p->esz; buffer 0 carries 256 B of halo arrival flags at its tail
        const size_t bytes = ((size_t)p->alloc * p->esz + 255) / 256 * 256;
        TK_TRY(plan_alloc(p, reinterpret_cast<char **>(&p->buf[0]), bytes + 256, true));
        p->flags = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(p->buf[0]) + bytes);
        TK_TRY(plan_alloc(p, reinterpret_cast<char **>(&p->buf[1]), bytes, true));
    }

    tk::LatDev &P = p->lat;
    P.S = (int)S; P.Lloc = (int)Lloc; P.halo = (int)halo; P.ntab = ntab;
    P.has_below = zb > 0; P.has_above = ze < L;
    P.seg_flag = d_seg_flag; P.det_flags = d_det_flags; P.det_attr = d_det_attr; P.det_corr = d_det_corr;
    P.attr_im = d_attr_im; P.attr_v = d_attr_v; P.corr_a = d_ca; P.corr_b = d_cb; P.corr_sign = d_cs;
    P.rowinfo = p->rowinfo; P.colinfo = p->colinfo;
    *out = p;
    return TK_OK;
}
#undef TK_TRY

// ------------------------------------------------------------------------------ common
int tk_plan_set_stream(tk_plan *p, void *cuda_stream)
{
    TK_REQUIRE(p, "plan is NULL");
    TK_CUDA(cudaSetDevice(p->device));
    if (p->own_stream && p->stream) {
        TK_CUDA(cudaStreamSynchronize(p->stream));
        TK_CUDA(cudaStreamDestroy(p->stream));
    }
    p->stream = (cudaStream_t)cuda_stream;
    p->own_stream = false;
    return TK_OK;
}

int tk_plan_num_nodes(const tk_plan *p, int64_t *n)
{
    TK_REQUIRE(p && n, "NULL argument");
    *n = p->n;
    return TK_OK;
}

int tk_plan_max_degree(tk_plan *p, int64_t *max_degree)
{
    TK_REQUIRE(p && max_degree, "NULL argument");
    if (p->max_degree < 0) {
        TK_CUDA(cudaSetDevice(p->device));
        int *d_out = nullptr;
        TK_CUDA(cudaMalloc((void **)&d_out, sizeof(int)));
        TK_CUDA(cudaMemsetAsync(d_out, 0, sizeof(int), p->stream));
        tk::lat_maxdeg_kernel<<<p->sm_count * 8, 256, 0, p->stream>>>(p->lat, d_out);
        p->launches++;
        int h = 0;
        cudaError_t e = cudaMemcpyAsync(&h, d_out, sizeof(int), cudaMemcpyDeviceToHost, p->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
        cudaFree(d_out);
        TK_CUDA(e);
        p->max_degree = h;
    }
    *max_degree = p->max_degree;
    return TK_OK;
}

int tk_plan_set_state_sparse(tk_plan *p, int64_t n, const int64_t *index, const double *value)
{
    TK_REQUIRE(p, "plan is NULL");
    TK_REQUIRE(n == 0 || (index && value), "NULL kick arrays");
    TK_CUDA(cudaSetDevice(p->device));
    TK_CUDA(cudaMemsetAsync(p->buf[0], 0, (size_t)p->alloc * p->esz, p->stream));
    TK_CUDA(cudaMemsetAsync(p->buf[1], 0, (size_t)p->alloc * p->esz, p->stream));
    p->cur = 0;
    p->sums_cur_valid = false;
    p->sums_prev = 2;  // u_{t-1} = 0
    if (n > 0) {
        std::vector<long long> h(n);
        for (int64_t i = 0; i < n; ++i) {
            TK_REQUIRE(index[i] >= 0 && index[i] < p->n, "kick index %lld out of range [0,%lld)",
                       (long long)index[i], p->n);
            h[i] = index[i] + p->state_off;
        }
        long long *d_idx = nullptr;
        double *d_val = nullptr;
        TK_CUDA(cudaMalloc((void **)&d_idx, n * sizeof(long long)));
        cudaError_t e = cudaMalloc((void **)&d_val, n * sizeof(double));
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_idx, h.data(), n * sizeof(long long), cudaMemcpyHostToDevice, p->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_val, value, n * sizeof(double), cudaMemcpyHostToDevice, p->stream);
        if (e == cudaSuccess) {
            if (is_f32(p))
                tk::scatter_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, p->stream>>>(
                    reinterpret_cast<float *>(p->buf[0]), d_idx, d_val, n);
            else
                tk::scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, p->stream>>>(p->buf[0], d_idx, d_val, n);
            p->launches++;
            e = cudaGetLastError();
        }
        // a defect-free sheet with a handful of distinct kicks: the row/column sums follow from the kicks
        bool distinct = p->kind == kLattice && p->plain2d && n <= 64;
        for (int64_t i = 0; distinct && i < n; ++i)
            for (int64_t j = 0; j < i; ++j)
                if (h[i] == h[j]) { distinct = false; break; }
        if (e == cudaSuccess && distinct) {
            tk::lat_sparse_sums_kernel<<<(unsigned)((2 * p->lat.S * 4 + 255) / 256), 256, 0, p->stream>>>(
                p->lat.S, (int)n, d_idx, d_val, p->rowinfo, p->colinfo);
            p->launches++;
            e = cudaGetLastError();
            p->sums_cur_valid = e == cudaSuccess;
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
        cudaFree(d_idx);
        cudaFree(d_val);
        TK_CUDA(e);
    }
    return TK_OK;
}

int tk_plan_set_state(tk_plan *p, const double *u, const double *u_prev)
{
    TK_REQUIRE(p && u, "NULL argument");
    TK_CUDA(cudaSetDevice(p->device));
    p->sums_cur_valid = false;
    p->sums_prev = u_prev ? 0 : 2;
    TK_CUDA(cudaMemsetAsync(p->buf[0], 0, (size_t)p->alloc * p->esz, p->stream));
    TK_CUDA(cudaMemsetAsync(p->buf[1], 0, (size_t)p->alloc * p->esz, p->stream));
    p->cur = 0;
    if (is_f32(p)) {
        int rc = stage_f64_to_f32(p, u, 0);
        if (rc == TK_OK && u_prev) rc = stage_f64_to_f32(p, u_prev, 1);
        if (rc) return rc;
    } else {
        TK_CUDA(cudaMemcpyAsync(p->buf[0] + p->state_off, u, (size_t)p->n * sizeof(double),
                                cudaMemcpyHostToDevice, p->stream));
        if (u_prev)
            TK_CUDA(cudaMemcpyAsync(p->buf[1] + p->state_off, u_prev, (size_t)p->n * sizeof(double),
                                    cudaMemcpyHostToDevice, p->stream));
    }
    if (!p->dead_local.empty()) {  // removed nodes never carry amplitude
        const long long nd = (long long)p->dead_local.size();
        std::vector<long long> h(nd);
        for (long long i = 0; i < nd; ++i) h[i] = p->dead_local[i] + p->state_off;
        std::vector<double> z(nd, 0.0);
        long long *d_idx = nullptr;
        double *d_val = nullptr;
        TK_CUDA(cudaMalloc((void **)&d_idx, nd * sizeof(long long)));
        cudaError_t e = cudaMalloc((void **)&d_val, nd * sizeof(double));
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_idx, h.data(), nd * sizeof(long long), cudaMemcpyHostToDevice, p->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_val, z.data(), nd * sizeof(double), cudaMemcpyHostToDevice, p->stream);
        if (e == cudaSuccess) {
            const unsigned g = (unsigned)((nd + 255) / 256);
            if (is_f32(p)) {
                tk::scatter_f32_kernel<<<g, 256, 0, p->stream>>>(reinterpret_cast<float *>(p->buf[0]), d_idx, d_val, nd);
                tk::scatter_f32_kernel<<<g, 256, 0, p->stream>>>(reinterpret_cast<float *>(p->buf[1]), d_idx, d_val, nd);
            } else {
                tk::scatter_kernel<<<g, 256, 0, p->stream>>>(p->buf[0], d_idx, d_val, nd);
                tk::scatter_kernel<<<g, 256, 0, p->stream>>>(p->buf[1], d_idx, d_val, nd);
            }
            p->launches += 2;
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
        cudaFree(d_idx);
        cudaFree(d_val);
        TK_CUDA(e);
    }
    TK_CUDA(cudaStreamSynchronize(p->stream));
    return TK_OK;
}

int tk_plan_get_state(tk_plan *p, double *u, double *u_prev)
{
    TK_REQUIRE(p, "plan is NULL");
    TK_CUDA(cudaSetDevice(p->device));
    if (is_f32(p)) {
        int rc = TK_OK;
        if (u) rc = stage_f32_to_f64(p, p->cur, u);
        if (rc == TK_OK && u_prev) rc = stage_f32_to_f64(p, p->cur ^ 1, u_prev);
        return rc;
    }
    if (u)
        TK_CUDA(cudaMemcpyAsync(u, p->buf[p->cur] + p->state_off, (size_t)p->n * sizeof(double),
                                cudaMemcpyDeviceToHost, p->stream));
    if (u_prev)
        TK_CUDA(cudaMemcpyAsync(u_prev, p->buf[p->cur ^ 1] + p->state_off, (size_t)p->n * sizeof(double),
                                cudaMemcpyDeviceToHost, p->stream));
    TK_CUDA(cudaStreamSynchronize(p->stream));
    return TK_OK;
}

int tk_lattice_halo_ptrs(tk_plan *p, int32_t next_state, void **send_lo, void **send_hi,
                         void **recv_lo, void **recv_hi, int64_t *plane_bytes)
{
    TK_REQUIRE(p && p->kind == kLattice, "not a lattice plan");
    TK_REQUIRE(p->lat.halo, "2-D sheets have no halo planes");
    const long long plane = (long long)p->lat.S * p->lat.S * 4;
    const int b = next_state ? (p->cur ^ 1) : p->cur;
    if (recv_lo) *recv_lo = bptr(p, b, 0);
    if (send_lo) *send_lo = bptr(p, b, plane);
    if (send_hi) *send_hi = bptr(p, b, plane * p->lat.Lloc);
    if (recv_hi) *recv_hi = bptr(p, b, plane * (p->lat.Lloc + 1));
    if (plane_bytes) *plane_bytes = (int64_t)(plane * p->esz);
    return TK_OK;
}

int tk_lattice_step_begin(tk_plan *p, double coeff, double damping)
{
    TK_REQUIRE(p && p->kind == kLattice, "not a lattice plan");
    TK_CUDA(cudaSetDevice(p->device));
    p->coeff = coeff;
    p->damping = damping;
    if (p->plain2d && p->sums_cur_valid) {  // the sums of u_t are current: only the probes' row 0 is due
        if (double *row = next_probe_row(p)) {
            tk::gather_kernel<<<(p->n_probes + 127) / 128, 128, 0, p->stream>>>(
                p->buf[p->cur], p->probe_idx, row, p->n_probes, nullptr);
            p->launches++;
            TK_CUDA(cudaGetLastError());
        }
        return TK_OK;
    }
    int rc = launch_lat_dispatch<1>(p, 0, p->lat.Lloc);
    if (rc) return rc;
    rc = lat_finalise(p, p->buf[p->cur], next_probe_row(p));
    p->sums_cur_valid = rc == TK_OK;
    return rc;
}

int tk_lattice_step_layers(tk_plan *p, int32_t zl_begin, int32_t zl_end)
{
    TK_REQUIRE(p && p->kind == kLattice, "not a lattice plan");
    TK_REQUIRE(0 <= zl_begin && zl_begin <= zl_end && zl_end <= p->lat.Lloc, "bad layer range");
    if (p->wz3d) return launch_lat3d_dispatch(p, zl_begin, zl_end);
    // A thin launch (the one-layer boundary launches of a slab run) with the 8-layers-per-CTA shape
    // would leave 7 of 8 warps of every CTA idle while still occupying its slot; use the
    // 8-segments-per-CTA shape for it (same partial-sum layout, so launches can be mixed in a step).
    const int wc_saved = p->wc;
    if (p->lat.halo && p->cpl == 1 && p->wc == 1 && zl_end - zl_begin <= 2 && lat_supported(1, 8, p->minb))
        p->wc = 8;
    const int rc = p->energy_on ? launch_lat_dispatch<2>(p, zl_begin, zl_end)
                                : launch_lat_dispatch<0>(p, zl_begin, zl_end);
    p->wc = wc_saved;
    return rc;
}

int tk_lattice_step_layers_on(tk_plan *p, int32_t zl_begin, int32_t zl_end, void *cuda_stream)
{
    TK_REQUIRE(p && p->kind == kLattice, "not a lattice plan");
    TK_REQUIRE(p->wz3d == 0, "side-stream launches are not available with TK_LAT3D=1");
    p->launch_stream = (cudaStream_t)cuda_stream;
    const int rc = tk_lattice_step_layers(p, zl_begin, zl_end);
    p->launch_stream = nullptr;
    return rc;
}

int tk_lattice_step_end(tk_plan *p)
{
    TK_REQUIRE(p && p->kind == kLattice, "not a lattice plan");
    p->sums_prev = 0;  // rowinfo now describes the new state; the old sums are gone
    p->cur ^= 1;
    return lat_finalise(p, p->buf[p->cur], next_probe_row(p));
}

int tk_plan_run(tk_plan *p, const tk_run_args *a)
{
    TK_REQUIRE(p && a, "NULL argument");
    TK_REQUIRE(p->kind != kSheetBand, "row-band plans are stepped with tk_sheet_pass_local / tk_sheet_pass_commit");
    TK_REQUIRE(a->steps >= 0, "steps must be >= 0");
    TK_REQUIRE(a->n_probes == 0 || (a->probe_index && a->probe_out), "probe arrays NULL");
    TK_CUDA(cudaSetDevice(p->device));
    int rc = set_probes(p, a->n_probes, a->probe_index, a->steps + 1);
    if (rc) return rc;
    TK_REQUIRE(!(is_f32(p) && a->energy_out), "the energy series is fp64-only");
    rc = set_energy(p, a->energy_out != nullptr && a->steps > 0, a->steps);
    if (rc) return rc;

    // optional full history (physics.py:104,129), staged through a bounded device buffer
    const size_t row_bytes = (size_t)p->n * sizeof(double);
    long long chunk_rows = 0;
    double *d_hist = nullptr;
    if (a->history_out && p->n > 0) {
        const long long rows_total = a->steps / (a->history_stride > 1 ? a->history_stride : 1) + 1;
        chunk_rows = std::max<long long>(1, std::min<long long>(rows_total, (long long)(kHistoryChunkBytes / row_bytes)));
        TK_CUDA(cudaMalloc((void **)&d_hist, (size_t)chunk_rows * row_bytes));
    }
    long long hist_fill = 0, hist_done = 0, hist_step = 0;
    const long long hist_stride = a->history_stride > 1 ? a->history_stride : 1;
    auto record = [&]() -> cudaError_t {
        if (!d_hist) return cudaSuccess;
        if ((hist_step++ % hist_stride) != 0) return cudaSuccess;
        cudaError_t e = cudaSuccess;
        if (is_f32(p)) {
            tk::f32_to_f64_kernel<<<1024, 256, 0, p->stream>>>(
                reinterpret_cast<const float *>(bptr(p, p->cur, p->state_off)), d_hist + (size_t)hist_fill * p->n, p->n);
            p->launches++;
            e = cudaGetLastError();
        } else {
            e = cudaMemcpyAsync(d_hist + (size_t)hist_fill * p->n, p->buf[p->cur] + p->state_off, row_bytes,
                                cudaMemcpyDeviceToDevice, p->stream);
        }
        if (e != cudaSuccess) return e;
        if (++hist_fill == chunk_rows) {
            e = cudaMemcpyAsync(a->history_out + (size_t)hist_done * p->n, d_hist,
                                (size_t)hist_fill * row_bytes, cudaMemcpyDeviceToHost, p->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
            hist_done += hist_fill;
            hist_fill = 0;
        }
        return e;
    };
#define TK_RUN_TRY(x)                \
    do {                             \
        rc = (x);                    \
        if (rc) {                    \
            if (d_hist) cudaFree(d_hist); \
            return rc;               \
        }                            \
    } while (0)
#define TK_RUN_CUDA(x)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (x);                                                                 \
        if (e_ != cudaSuccess) {                                                              \
            if (d_hist) cudaFree(d_hist);                                                     \
            return fail(TK_ERR_CUDA, "%s failed: %s", #x, cudaGetErrorString(e_));            \
        }                                                                                     \
    } while (0)

    std::vector<cudaEvent_t> kev;  // pairs around each step kernel (bounded pool)
    if (a->step_kernel_ms) {
        const int64_t pairs = std::min<int64_t>(a->steps, 4096);
        kev.resize((size_t)pairs * 2, nullptr);
        for (auto &e : kev) TK_RUN_CUDA(cudaEventCreate(&e));
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (a->elapsed_ms) {
        TK_RUN_CUDA(cudaEventCreate(&ev0));
        TK_RUN_CUDA(cudaEventCreate(&ev1));
    }
    if (p->kind == kGraph) {
        if (double *row = next_probe_row(p)) {
            tk::gather_kernel<<<(p->n_probes + 127) / 128, 128, 0, p->stream>>>(p->buf[p->cur], p->probe_idx, row, p->n_probes, nullptr);
            p->launches++;
        }
        TK_RUN_CUDA(record());
        if (ev0) TK_RUN_CUDA(cudaEventRecord(ev0, p->stream));
    } else {
        TK_RUN_TRY(tk_lattice_step_begin(p, a->coeff, a->damping));
        TK_RUN_CUDA(record());
        if (ev0) TK_RUN_CUDA(cudaEventRecord(ev0, p->stream));
    }

    // ---- launch-bound problems: capture two steps (the buffers alternate) into a CUDA graph and replay
    // it; output rows (probes / history / energy) are addressed through device counters so the captured
    // launches are identical every time.  Large lattices gain nothing (their kernels take ~1 ms) and the
    // per-kernel event timing cannot be captured, so both keep the direct launches below.
    int64_t steps_left = a->steps;
    size_t kev_used = 0;      // event pairs recorded
    int64_t kev_steps = 0;    // steps those pairs cover
    p->fused_steps_last = 0;

    // ---- defect-free 2-D sheets: K steps per pass over the state (kernels_fused.cuh); the remainder
    // (steps mod K) runs step by step below
    const int fk = fuse_choice(p, a);
    if (fk > 1 && steps_left >= fk) {
        TK_RUN_TRY(fuse_setup(p, fk));
        TK_RUN_TRY(fuse_prepare(p));
        p->fuse_k_run = fk;
        const int64_t npass = steps_left / fk;
        const bool fgraph = npass >= 4 && !a->step_kernel_ms && p->n <= (1LL << 24) && env_int("TK_CUDA_GRAPH", 1) != 0;
        if (fgraph) {
            if (!p->ctr_dev) TK_RUN_CUDA(cudaMalloc((void **)&p->ctr_dev, 4 * sizeof(long long)));
            const long long ctr0[4] = {p->probe_row, 0, 0, 0};
            TK_RUN_CUDA(cudaMemcpyAsync(p->ctr_dev, ctr0, sizeof(ctr0), cudaMemcpyHostToDevice, p->stream));
            TK_RUN_CUDA(cudaStreamSynchronize(p->stream));
            p->ctr_mode = true;
            const long long launches_before = p->launches;
            cudaGraph_t graph = nullptr;
            cudaGraphExec_t exec = nullptr;
            cudaError_t ce = cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal);
            int rcap = TK_OK;
            if (ce == cudaSuccess) {
                rcap = fuse_pass(p, fk, nullptr, nullptr);
                if (rcap == TK_OK) {
                    tk::advance_kernel<<<1, 32, 0, p->stream>>>(p->ctr_dev, fk);
                    p->launches++;
                }
                ce = cudaStreamEndCapture(p->stream, &graph);
            }
            if (ce == cudaSuccess && rcap == TK_OK) ce = cudaGraphInstantiate(&exec, graph, 0);
            const long long per_pass = p->launches - launches_before;
            for (int64_t i = 0; i < npass && ce == cudaSuccess && rcap == TK_OK; ++i) ce = cudaGraphLaunch(exec, p->stream);
            if (exec) cudaGraphExecDestroy(exec);
            if (graph) cudaGraphDestroy(graph);
            p->ctr_mode = false;
            if (ce != cudaSuccess || rcap != TK_OK) {
                if (d_hist) cudaFree(d_hist);
                if (rcap != TK_OK) return rcap;
                return fail(TK_ERR_CUDA, "CUDA graph of the fused pass failed: %s", cudaGetErrorString(ce));
            }
            p->launches = launches_before + per_pass * npass;
            if (p->n_probes > 0) p->probe_row += npass * fk;
        } else {
            for (int64_t i = 0; i < npass; ++i) {
                const bool tk_ = kev_used + 2 <= kev.size();
                TK_RUN_TRY(fuse_pass(p, fk, tk_ ? kev[kev_used] : nullptr, tk_ ? kev[kev_used + 1] : nullptr));
                if (tk_) { kev_used += 2; kev_steps += fk; }
            }
        }
        steps_left -= npass * fk;
        p->fused_steps_last = npass * fk;
        if (steps_left >= 2) {  // the remainder as one shorter pass (a single left-over step runs below)
            const int kr = (int)steps_left;
            const bool tk_ = kev_used + 2 <= kev.size();
            TK_RUN_TRY(fuse_pass(p, kr, tk_ ? kev[kev_used] : nullptr, tk_ ? kev[kev_used + 1] : nullptr));
            if (tk_) { kev_used += 2; kev_steps += kr; }
            p->fused_steps_last += kr;
            steps_left = 0;
        }
    }
    const bool hist_in_one_chunk = !d_hist || (hist_stride == 1 && chunk_rows >= a->steps + 1);
    const bool use_graph = steps_left >= 8 && !a->step_kernel_ms && hist_in_one_chunk && p->n <= (1LL << 24) &&
                           env_int("TK_CUDA_GRAPH", 1) != 0;
    if (use_graph) {
        if (!p->ctr_dev) {
            TK_RUN_CUDA(cudaMalloc((void **)&p->ctr_dev, 4 * sizeof(long long)));
        }
        const long long ctr0[4] = {p->probe_row, hist_fill, p->energy_row, 0};
        TK_RUN_CUDA(cudaMemcpyAsync(p->ctr_dev, ctr0, sizeof(ctr0), cudaMemcpyHostToDevice, p->stream));
        TK_RUN_CUDA(cudaStreamSynchronize(p->stream));  // ctr0 is a stack array
        auto one_step = [&]() -> int {
            int r = p->kind == kGraph ? graph_step(p, a->coeff, a->damping)
                                      : tk_lattice_step_layers(p, 0, p->lat.Lloc);
            if (r == TK_OK && p->kind != kGraph) r = tk_lattice_step_end(p);
            if (r) return r;
            if (d_hist) {
                tk::history_row_kernel<<<256, 256, 0, p->stream>>>(bptr(p, p->cur, p->state_off), is_f32(p) ? 1 : 0,
                                                                  d_hist, p->n, p->ctr_dev);
                p->launches++;
            }
            tk::advance_kernel<<<1, 32, 0, p->stream>>>(p->ctr_dev, 1);
            p->launches++;
            return cudaGetLastError() == cudaSuccess ? TK_OK : fail(TK_ERR_CUDA, "graph-mode launch failed");
        };
        p->ctr_mode = true;
        const long long launches_before = p->launches;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaError_t ce = cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal);
        int rcap = TK_OK;
        if (ce == cudaSuccess) {
            rcap = one_step();
            if (rcap == TK_OK) rcap = one_step();
            ce = cudaStreamEndCapture(p->stream, &graph);
        }
        if (ce == cudaSuccess && rcap == TK_OK) ce = cudaGraphInstantiate(&exec, graph, 0);
        const long long per_pair = p->launches - launches_before;
        if (ce != cudaSuccess || rcap != TK_OK) {
            p->ctr_mode = false;
            if (graph) cudaGraphDestroy(graph);
            if (d_hist) cudaFree(d_hist);
            if (rcap != TK_OK) return rcap;
            return fail(TK_ERR_CUDA, "CUDA graph capture failed: %s", cudaGetErrorString(ce));
        }
        const int64_t pairs = steps_left / 2;
        for (int64_t i = 0; i < pairs && ce == cudaSuccess; ++i) ce = cudaGraphLaunch(exec, p->stream);
        p->launches = launches_before + per_pair * pairs;
        if (ce == cudaSuccess && (steps_left & 1)) {
            const int r = one_step();
            if (r) ce = cudaErrorUnknown;
        }
        cudaGraphExecDestroy(exec);
        cudaGraphDestroy(graph);
        p->ctr_mode = false;
        if (ce != cudaSuccess) {
            if (d_hist) cudaFree(d_hist);
            return fail(TK_ERR_CUDA, "CUDA graph replay failed: %s", cudaGetErrorString(ce));
        }
        p->probe_row += p->n_probes > 0 ? steps_left : 0;
        p->energy_row += p->energy_on ? steps_left : 0;
        if (d_hist) hist_fill += steps_left;
        steps_left = 0;
    }
    if (p->kind == kGraph) {
        for (int64_t s = 0; s < steps_left; ++s) {
            const bool tk_ = kev_used + 2 <= kev.size();
            if (tk_) TK_RUN_CUDA(cudaEventRecord(kev[kev_used], p->stream));
            TK_RUN_TRY(graph_step(p, a->coeff, a->damping));
            if (tk_) TK_RUN_CUDA(cudaEventRecord(kev[kev_used + 1], p->stream));
            if (tk_) { kev_used += 2; kev_steps += 1; }
            TK_RUN_CUDA(record());
        }
    } else {
        for (int64_t s = 0; s < steps_left; ++s) {
            const bool tk_ = kev_used + 2 <= kev.size();
            if (tk_) TK_RUN_CUDA(cudaEventRecord(kev[kev_used], p->stream));
            TK_RUN_TRY(tk_lattice_step_layers(p, 0, p->lat.Lloc));
            if (tk_) TK_RUN_CUDA(cudaEventRecord(kev[kev_used + 1], p->stream));
            if (tk_) { kev_used += 2; kev_steps += 1; }
            TK_RUN_TRY(tk_lattice_step_end(p));
            TK_RUN_CUDA(record());
        }
    }
    if (ev1) TK_RUN_CUDA(cudaEventRecord(ev1, p->stream));
    if (d_hist && hist_fill > 0)
        TK_RUN_CUDA(cudaMemcpyAsync(a->history_out + (size_t)hist_done * p->n, d_hist,
                                    (size_t)hist_fill * row_bytes, cudaMemcpyDeviceToHost, p->stream));
    if (p->energy_on)
        TK_RUN_CUDA(cudaMemcpyAsync(a->energy_out, p->energy_series, (size_t)a->steps * sizeof(double),
                                    cudaMemcpyDeviceToHost, p->stream));
    if (p->n_probes > 0)
        TK_RUN_CUDA(cudaMemcpyAsync(a->probe_out, p->probe_series,
                                    (size_t)(a->steps + 1) * p->n_probes * sizeof(double),
                                    cudaMemcpyDeviceToHost, p->stream));
    TK_RUN_CUDA(cudaStreamSynchronize(p->stream));
    if (a->elapsed_ms) {
        float ms = 0.f;
        TK_RUN_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        *a->elapsed_ms = (double)ms;
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
    }
    if (a->step_kernel_ms) {
        double total = 0.0;
        for (size_t i = 0; i + 1 < kev_used; i += 2) {
            float ms = 0.f;
            TK_RUN_CUDA(cudaEventElapsedTime(&ms, kev[i], kev[i + 1]));
            total += ms;
        }
        // scale to all steps if the pool was smaller than the step count
        if (kev_steps > 0) total *= (double)a->steps / (double)kev_steps;
        *a->step_kernel_ms = total;
        for (auto &e : kev) cudaEventDestroy(e);
    }
    if (d_hist) cudaFree(d_hist);
    p->energy_on = false;
    return TK_OK;
#undef TK_RUN_TRY
#undef TK_RUN_CUDA
}

// ---- fused halo push over peer memory
namespace {
typedef int (*stream_value64_fn)(void *stream, unsigned long long addr, unsigned long long value, unsigned int flags);
stream_value64_fn g_write64 = nullptr, g_wait64 = nullptr;

int load_stream_memops()
{
    if (g_write64 && g_wait64) return TK_OK;
    cudaDriverEntryPointQueryResult q;
    void *fw = nullptr, *fq = nullptr;
    TK_CUDA(cudaGetDriverEntryPoint("cuStreamWriteValue64", &fw, cudaEnableDefault, &q));
    TK_CUDA(cudaGetDriverEntryPoint("cuStreamWaitValue64", &fq, cudaEnableDefault, &q));
    if (!fw || !fq) return fail(TK_ERR_CUDA, "driver has no 64-bit stream memory operations");
    g_write64 = (stream_value64_fn)fw;
    g_wait64 = (stream_value64_fn)fq;
    return TK_OK;
}
}  // namespace

int tk_lattice_ipc_export(tk_plan *p, void *handles)
{
    TK_REQUIRE(p && p->kind == kLattice && handles, "bad argument");
    TK_CUDA(cudaSetDevice(p->device));
    cudaIpcMemHandle_t h[2];
    TK_CUDA(cudaIpcGetMemHandle(&h[0], p->buf[0]));
    TK_CUDA(cudaIpcGetMemHandle(&h[1], p->buf[1]));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    memcpy(handles, h, 128);
    return TK_OK;
}

int tk_ipc_open(int device, const void *handle, void **dev_ptr)
{
    TK_REQUIRE(handle && dev_ptr, "NULL argument");
    int rc = check_device(device);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    TK_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TK_OK;
}

int tk_ipc_close(int device, void *dev_ptr)
{
    int rc = check_device(device);
    if (rc) return rc;
    if (dev_ptr) TK_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return TK_OK;
}

int tk_lattice_state_ptrs(tk_plan *p, void **buf0, void **buf1)
{
    TK_REQUIRE(p && p->kind == kLattice, "not a lattice plan");
    if (buf0) *buf0 = p->buf[0];
    if (buf1) *buf1 = p->buf[1];
    return TK_OK;
}

int tk_lattice_peer_attach(tk_plan *p, int32_t side, void *b0, void *b1, int32_t peer_layers)
{
    TK_REQUIRE(p && p->kind == kLattice && p->lat.halo, "needs a 3-D lattice plan");
    TK_REQUIRE(side == 0 || side == 1, "side must be 0 or 1");
    TK_REQUIRE((b0 == nullptr) == (b1 == nullptr), "attach both buffers or none");
    TK_REQUIRE(b0 == nullptr || peer_layers >= 1, "peer_layers must be >= 1");
    TK_REQUIRE(p->wz3d == 0, "peer halo push is not available with TK_LAT3D=1");
    p->peer_buf[side][0] = (double *)b0;
    p->peer_buf[side][1] = (double *)b1;
    p->peer_layers[side] = peer_layers;
    return TK_OK;
}

int tk_lattice_set_push_mode(tk_plan *p, int32_t in_kernel)
{
    TK_REQUIRE(p && p->kind == kLattice, "not a lattice plan");
    p->push_in_kernel = in_kernel != 0;
    return TK_OK;
}

int tk_lattice_push_halos(tk_plan *p, void *cuda_stream)
{
    TK_REQUIRE(p && p->kind == kLattice && p->lat.halo, "needs a 3-D lattice plan");
    TK_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : p->stream;
    const long long plane = (long long)p->lat.S * p->lat.S * 4;
    const size_t bytes = (size_t)plane * p->esz;
    const int b = p->cur;
    if (p->peer_buf[0][b])  // my first owned plane -> lower neighbour's upper halo
        TK_CUDA(cudaMemcpyAsync(pptr(p, 0, b, (p->peer_layers[0] + 1) * plane), bptr(p, b, plane), bytes,
                                cudaMemcpyDeviceToDevice, st));
    if (p->peer_buf[1][b])  // my last owned plane -> upper neighbour's lower halo
        TK_CUDA(cudaMemcpyAsync(pptr(p, 1, b, 0), bptr(p, b, plane * p->lat.Lloc), bytes,
                                cudaMemcpyDeviceToDevice, st));
    return TK_OK;
}

int tk_lattice_signal(tk_plan *p, uint64_t value, void *cuda_stream)
{
    TK_REQUIRE(p && p->kind == kLattice && p->lat.halo, "needs a 3-D lattice plan");
    TK_CUDA(cudaSetDevice(p->device));
    int rc = load_stream_memops();
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : p->stream;
    const size_t plane = (size_t)p->lat.S * p->lat.S * 4;
    auto peer_tail = [&](int side) {  // the neighbour's flag block sits after its 256-byte aligned state
        const size_t bytes = ((size_t)(p->peer_layers[side] + 2) * plane * p->esz + 255) / 256 * 256;
        return reinterpret_cast<double *>(reinterpret_cast<char *>(p->peer_buf[side][0]) + bytes);
    };
    if (p->peer_buf[0][0]) {  // lower neighbour's "from above" flag: tail of its buffer 0, word 16
        double *tail = peer_tail(0);
        const int e = g_write64(st, (unsigned long long)(reinterpret_cast<unsigned long long *>(tail) + 16), value, 0);
        if (e) return fail(TK_ERR_CUDA, "cuStreamWriteValue64 failed with CUresult %d", e);
    }
    if (p->peer_buf[1][0]) {  // upper neighbour's "from below" flag: word 0
        double *tail = peer_tail(1);
        const int e = g_write64(st, (unsigned long long)reinterpret_cast<unsigned long long *>(tail), value, 0);
        if (e) return fail(TK_ERR_CUDA, "cuStreamWriteValue64 failed with CUresult %d", e);
    }
    return TK_OK;
}

int tk_lattice_wait(tk_plan *p, uint64_t value, void *cuda_stream)
{
    TK_REQUIRE(p && p->kind == kLattice && p->lat.halo, "needs a 3-D lattice plan");
    TK_CUDA(cudaSetDevice(p->device));
    int rc = load_stream_memops();
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : p->stream;
    const unsigned int GEQ = 0x1;  // CU_STREAM_WAIT_VALUE_GEQ
    if (p->peer_buf[0][0]) {
        const int e = g_wait64(st, (unsigned long long)(p->flags), value, GEQ);
        if (e) return fail(TK_ERR_CUDA, "cuStreamWaitValue64 failed with CUresult %d", e);
    }
    if (p->peer_buf[1][0]) {
        const int e = g_wait64(st, (unsigned long long)(p->flags + 16), value, GEQ);
        if (e) return fail(TK_ERR_CUDA, "cuStreamWaitValue64 failed with CUresult %d", e);
    }
    return TK_OK;
}

int tk_lattice_tiling(const tk_plan *p, int32_t *cpl, int32_t *wc, int32_t *rchunk, int32_t *occ)
{
    TK_REQUIRE(p && p->kind == kLattice, "not a lattice plan");
    if (cpl) *cpl = p->cpl;
    if (wc) *wc = p->wz3d ? -p->wz3d * 100 - p->stages3d : p->wc;  // 3-D: -(100*layers_per_cta + stages)
    if (rchunk) *rchunk = p->rchunk;
    if (occ) *occ = p->wz3d ? lat3d_occupancy(p->wz3d, p->stages3d, p->minb3d)
                    : p->bulk_stages ? (p->lat.halo ? lat_bulk_occ<1, true, 2>(p->bulk_stages)
                                                    : lat_bulk_occ<8, false, 3>(p->bulk_stages))
                                     : lat_occupancy(p->cpl, p->wc, p->minb, p->lat.halo != 0);
    return TK_OK;
}

int tk_plan_fused_info(tk_plan *p, int32_t *steps_per_pass, int64_t *fused_steps, int32_t *rows_per_cta,
                       int32_t *ctas_per_sm)
{
    TK_REQUIRE(p, "NULL plan");
    if (steps_per_pass) *steps_per_pass = p->fused_steps_last > 0 ? p->fuse_k_run : 0;
    if (fused_steps) *fused_steps = p->fused_steps_last;
    if (rows_per_cta) *rows_per_cta = p->fuse_rchunk;
    if (ctas_per_sm) *ctas_per_sm = p->fuse_occ;
    return TK_OK;
}

// ------------------------------------------------------------------ one sheet sharded over GPUs by rows
#define TK_BAND_TRY(x)     \
    do {                   \
        rc = (x);          \
        if (rc) {          \
            tk_plan_destroy(p); \
            return rc;     \
        }                  \
    } while (0)
int tk_sheet_band_plan_create(int device, int32_t size, int32_t row_begin, int32_t row_end, tk_plan **out)
{
    TK_REQUIRE(out, "out is NULL");
    *out = nullptr;
    TK_REQUIRE(size >= 1 && 0 <= row_begin && row_begin < row_end && row_end <= size,
               "row band [%d,%d) not inside [0,%d)", row_begin, row_end, size);
    int rc = check_device(device);
    if (rc) return rc;
    tk_plan *p = new tk_plan();
    p->kind = kSheetBand;
    p->esz = 8;
    const long long S = size, R = row_end - row_begin;
    p->n = R * S * 4;
    p->alloc = p->n;
    p->state_off = 0;
    p->nrows = (int)R;
    p->row_begin = row_begin;
    p->lat.S = size;
    p->lat.Lloc = 1;
    p->max_degree = 2 * S + 1;  // lattice.py:30-73 without defects: 3 in the cell + 2 (S-1) in the row and column cliques
    p->pf_rows = env_int("TK_LAT_PREFETCH", 1);
    TK_BAND_TRY(plan_common_init(p, device));
    TK_BAND_TRY(plan_alloc(p, &p->buf[0], (size_t)p->alloc, true));
    TK_BAND_TRY(plan_alloc(p, &p->buf[1], (size_t)p->alloc, true));
    TK_BAND_TRY(plan_alloc(p, &p->rowinfo, (size_t)R * 8, true));
    TK_BAND_TRY(plan_alloc(p, &p->colinfo, (size_t)S * 8, true));
    TK_BAND_TRY(plan_alloc(p, &p->prow, (size_t)R * ((S + 31) / 32) * 4, true));
    TK_BAND_TRY(plan_alloc(p, &p->ftot_global, 8, true));
    TK_BAND_TRY(fuse_setup(p, 64));  // response tables, previous-level sums, partial buffers, tiling
    *out = p;
    return TK_OK;
}

#undef TK_BAND_TRY

int tk_sheet_exchange_doubles(tk_plan *p, int64_t *count)
{
    TK_REQUIRE(p && p->kind == kSheetBand && count, "not a row-band plan");
    *count = 8LL * p->lat.S + 8;
    return TK_OK;
}

int tk_sheet_pass_local(tk_plan *p, int32_t k, double coeff, double damping, void *exchange)
{
    TK_REQUIRE(p && p->kind == kSheetBand && exchange, "not a row-band plan / NULL exchange buffer");
    TK_REQUIRE(k >= 0, "k must be >= 0");
    TK_CUDA(cudaSetDevice(p->device));
    p->coeff = coeff;
    p->damping = damping;
    int rc = fuse_setup(p, k);
    if (rc) return rc;
    if (k == 0) {  // start of a run: probes' row 0; the pass below only produces the sums of both levels
        if (double *row = next_probe_row(p)) {
            tk::gather_kernel<<<(p->n_probes + 127) / 128, 128, 0, p->stream>>>(p->buf[p->cur], p->probe_idx, row,
                                                                              p->n_probes, nullptr);
            p->launches++;
        }
    }
    if (p->band_ev.empty()) {
        p->band_ev.resize(128, nullptr);
        for (auto &e : p->band_ev) TK_CUDA(cudaEventCreate(&e));
    }
    const size_t slot = (size_t)(p->band_ev_next % 64) * 2;
    rc = fuse_pass(p, k, p->band_ev[slot], p->band_ev[slot + 1]);
    if (rc) return rc;
    p->band_ev_next++;
    const int S = p->lat.S;
    tk::sheet_pack_kernel<<<(unsigned)((8 * S + 8 + 255) / 256), 256, 0, p->stream>>>(
        S, (p->nrows + 7) / 8, p->colinfo, p->colinfo_prev, p->ftot, static_cast<double *>(exchange));
    p->launches++;
    TK_CUDA(cudaGetLastError());
    return TK_OK;
}

int tk_sheet_kernel_ms(tk_plan *p, double *total_ms, int64_t *passes)
{
    TK_REQUIRE(p && p->kind == kSheetBand && total_ms && passes, "not a row-band plan / NULL argument");
    TK_CUDA(cudaSetDevice(p->device));
    TK_CUDA(cudaStreamSynchronize(p->stream));
    const long long first = std::max(p->band_ev_read, p->band_ev_next - 64);
    double sum = 0.0;
    for (long long i = first; i < p->band_ev_next; ++i) {
        float ms = 0.f;
        TK_CUDA(cudaEventElapsedTime(&ms, p->band_ev[(size_t)(i % 64) * 2], p->band_ev[(size_t)(i % 64) * 2 + 1]));
        sum += ms;
    }
    *total_ms = sum;
    *passes = p->band_ev_next - first;
    p->band_ev_read = p->band_ev_next;
    return TK_OK;
}

int tk_sheet_pass_commit(tk_plan *p, const void *exchange)
{
    TK_REQUIRE(p && p->kind == kSheetBand && exchange, "not a row-band plan / NULL exchange buffer");
    TK_CUDA(cudaSetDevice(p->device));
    const int S = p->lat.S;
    tk::sheet_unpack_kernel<<<(unsigned)((8 * S + 8 + 255) / 256), 256, 0, p->stream>>>(
        S, static_cast<const double *>(exchange), p->colinfo, p->colinfo_prev, p->ftot_global);
    p->launches++;
    TK_CUDA(cudaGetLastError());
    return TK_OK;
}

int tk_plan_set_probes(tk_plan *p, int64_t n_probes, const int64_t *probe_index, int64_t rows)
{
    TK_REQUIRE(p, "plan is NULL");
    TK_REQUIRE(n_probes >= 0 && rows >= 0 && (n_probes == 0 || probe_index), "bad probe arguments");
    TK_CUDA(cudaSetDevice(p->device));
    return set_probes(p, n_probes, probe_index, rows);
}

int tk_plan_read_probes(tk_plan *p, double *out, int64_t *rows_recorded)
{
    TK_REQUIRE(p && rows_recorded, "NULL argument");
    TK_CUDA(cudaSetDevice(p->device));
    TK_CUDA(cudaStreamSynchronize(p->stream));
    *rows_recorded = p->probe_row;
    if (out && p->n_probes > 0 && p->probe_row > 0)
        TK_CUDA(cudaMemcpy(out, p->probe_series, (size_t)p->probe_row * p->n_probes * sizeof(double),
                           cudaMemcpyDeviceToHost));
    return TK_OK;
}

int tk_plan_launch_count(const tk_plan *p, int64_t *launches)
{
    TK_REQUIRE(p && launches, "NULL argument");
    *launches = p->launches;
    return TK_OK;
}

int tk_plan_destroy(tk_plan *p)
{
    if (!p) return TK_OK;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    for (void *q : p->owned) cudaFree(q);
    if (p->probe_idx) cudaFree(p->probe_idx);
    if (p->probe_series) cudaFree(p->probe_series);
    if (p->pen) cudaFree(p->pen);
    if (p->energy_series) cudaFree(p->energy_series);
    if (p->ctr_dev) cudaFree(p->ctr_dev);
    for (cudaEvent_t e : p->band_ev)
        if (e) cudaEventDestroy(e);
    if (p->own_stream && p->stream) cudaStreamDestroy(p->stream);
    delete p;
    return TK_OK;
}

// ------------------------------------------------------------------------------ DFT
int tk_dft_abs(int device, const double *series, int64_t batch, int64_t n, double *spectrum,
               int64_t *argmax)
{
    TK_REQUIRE(series && spectrum, "NULL argument");
    TK_REQUIRE(batch >= 1 && batch <= 65535 && n >= 1 && n < 0x7fffffffLL, "bad batch/n");
    int rc = check_device(device);
    if (rc) return rc;
    double *d_x = nullptr, *d_s = nullptr;
    long long *d_a = nullptr;
    const size_t bytes = (size_t)batch * n * sizeof(double);
    cudaError_t e = cudaMalloc((void **)&d_x, bytes);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_s, bytes);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_a, batch * sizeof(long long));
    if (e == cudaSuccess) e = cudaMemcpy(d_x, series, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        tk::dft_abs_kernel<<<dim3((unsigned)n, (unsigned)batch), 128>>>(d_x, n, d_s);
        tk::argmax_kernel<<<(unsigned)batch, 256>>>(d_s, n, d_a);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(spectrum, d_s, bytes, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && argmax) {
        static_assert(sizeof(long long) == sizeof(int64_t), "int64");
        e = cudaMemcpy(argmax, d_a, batch * sizeof(long long), cudaMemcpyDeviceToHost);
    }
    cudaFree(d_x);
    cudaFree(d_s);
    cudaFree(d_a);
    TK_CUDA(e);
    return TK_OK;
}

// ------------------------------------------------------------------------------ ensembles
int tk_ensemble_run(int device, const tk_ensemble_desc *d, double *probe_series, double *spectrum,
                    int64_t *argmax, double *elapsed_ms)
{
    TK_REQUIRE(d && probe_series, "NULL argument");
    TK_REQUIRE(d->size >= 1 && d->layers >= 1 && d->n_masks >= 1 && d->n_sims >= 0 && d->steps >= 0,
               "bad ensemble shape");
    TK_REQUIRE(d->cell_bits && (d->n_sims == 0 || (d->sim_mask && d->sim_kick && d->sim_coeff && d->sim_damping)),
               "NULL ensemble array");
    TK_REQUIRE(!argmax || spectrum, "argmax needs the spectrum buffer");
    int rc = check_device(device);
    if (rc) return rc;
    const long long S = d->size, L = d->layers, ncell = L * S * S, N = ncell * 4, nt = d->steps + 1;
    TK_REQUIRE(N < (1LL << 24), "lattice too large for an SM-resident ensemble");
    TK_REQUIRE(S <= 64 && L * S * 4 <= 65535, "ensemble lattices are limited to size <= 64 and layers*size <= 16383");
    const size_t smem = tk::ensemble_smem_bytes((int)S, (int)L, (int)d->steps);
    cudaDeviceProp prop;
    TK_CUDA(cudaGetDeviceProperties(&prop, device));
    TK_REQUIRE(smem <= (size_t)prop.sharedMemPerBlockOptin,
               "ensemble state (%zu B) exceeds shared memory per CTA (%zu B); use a lattice plan", smem,
               (size_t)prop.sharedMemPerBlockOptin);
    if (d->n_sims == 0) return TK_OK;
    for (long long i = 0; i < d->n_sims; ++i) {
        TK_REQUIRE(d->sim_mask[i] >= 0 && d->sim_mask[i] < d->n_masks, "sim_mask[%lld] out of range", i);
        TK_REQUIRE(d->sim_kick[i] >= 0 && d->sim_kick[i] < N, "sim_kick[%lld] out of range", i);
    }
    // dense attribute rows for the masks that carry mass / potential tags
    std::vector<int> attr_row(d->n_masks, -1);
    std::vector<double> attr_im, attr_v;
    if (d->exc_offset) {
        for (long long m = 0; m < d->n_masks; ++m) {
            const int b = d->exc_offset[m], e = d->exc_offset[m + 1];
            if (e <= b) continue;
            attr_row[m] = (int)(attr_im.size() / N);
            attr_im.resize(attr_im.size() + N, 1.0);
            attr_v.resize(attr_v.size() + N, 0.0);
            for (int k = b; k < e; ++k) {
                TK_REQUIRE(d->exc_node[k] >= 0 && d->exc_node[k] < N, "exc_node out of range");
                attr_im[attr_im.size() - N + d->exc_node[k]] = d->exc_inv_mass[k];
                attr_v[attr_v.size() - N + d->exc_node[k]] = d->exc_potential[k];
            }
        }
    }
    std::vector<int> corr_off(d->n_masks + 1, 0);
    long long ncorr = 0;
    if (d->corr_offset) {
        for (long long m = 0; m <= d->n_masks; ++m) corr_off[m] = d->corr_offset[m];
        ncorr = corr_off[d->n_masks];
        for (long long m = 0; m < d->n_masks; ++m)
            TK_REQUIRE(corr_off[m + 1] - corr_off[m] <= tk::kEnsMaxCorr, "more than %d edge corrections in mask %lld",
                       tk::kEnsMaxCorr, m);
    }
    std::vector<double> kickv(d->n_sims, 1.0);
    if (d->sim_kick_value) kickv.assign(d->sim_kick_value, d->sim_kick_value + d->n_sims);

    std::vector<void *> owned;
    auto up = [&](auto **dst, const auto *src, size_t count) -> int {
        using T = std::remove_cv_t<std::remove_pointer_t<decltype(src)>>;
        T *p = nullptr;
        TK_CUDA(cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T)));
        owned.push_back(p);
        if (count) TK_CUDA(cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice));
        *dst = p;
        return TK_OK;
    };
    tk::EnsDev P{};
    P.S = (int)S; P.L = (int)L; P.n_masks = (int)d->n_masks; P.steps = (int)d->steps;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&]() {
        for (void *q : owned) cudaFree(q);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    };
#define TK_ENS(x) do { rc = (x); if (rc) { cleanup(); return rc; } } while (0)
    TK_ENS(up(&P.cell_bits, d->cell_bits, (size_t)d->n_masks * ncell));
    TK_ENS(up(&P.attr_im, attr_im.data(), attr_im.size()));
    TK_ENS(up(&P.attr_v, attr_v.data(), attr_v.size()));
    TK_ENS(up(&P.mask_attr_row, attr_row.data(), attr_row.size()));
    TK_ENS(up(&P.corr_off, corr_off.data(), corr_off.size()));
    TK_ENS(up(&P.corr_a, d->corr_a, (size_t)ncorr));
    TK_ENS(up(&P.corr_b, d->corr_b, (size_t)ncorr));
    TK_ENS(up(&P.corr_sign, d->corr_sign, (size_t)ncorr));
    TK_ENS(up(&P.sim_mask, d->sim_mask, (size_t)d->n_sims));
    TK_ENS(up(&P.sim_kick, d->sim_kick, (size_t)d->n_sims));
    TK_ENS(up(&P.sim_coeff, d->sim_coeff, (size_t)d->n_sims));
    TK_ENS(up(&P.sim_damping, d->sim_damping, (size_t)d->n_sims));
    TK_ENS(up(&P.sim_kick_value, kickv.data(), kickv.size()));
    auto dalloc = [&](auto **dst, size_t count) -> int {
        using T = std::remove_pointer_t<std::remove_pointer_t<decltype(dst)>>;
        T *p = nullptr;
        TK_CUDA(cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T)));
        owned.push_back(p);
        *dst = p;
        return TK_OK;
    };
    TK_ENS(dalloc(&P.series, (size_t)d->n_sims * nt));
    if (spectrum) TK_ENS(dalloc(&P.spectrum, (size_t)d->n_sims * nt));
    if (argmax) TK_ENS(dalloc(&P.argmax, (size_t)d->n_sims));
#define TK_ENS_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); \
        return fail(TK_ERR_CUDA, "%s failed: %s", #x, cudaGetErrorString(e_)); } } while (0)
    TK_ENS_CUDA(cudaFuncSetAttribute(tk::ensemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // threads: whole iterations over the half cells (one lane per 2 quadrants), at most 448 per CTA
    const long long work = 2 * ncell;
    const long long iters = (work + 447) / 448;
    int threads = (int)(((work + iters - 1) / iters + 31) / 32 * 32);
    threads = std::max(64, std::min(threads, 448));
    int occ = 1;
    TK_ENS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tk::ensemble_kernel, threads, smem));
    const long long grid = std::min<long long>(d->n_sims, (long long)prop.multiProcessorCount * std::max(occ, 1));
    TK_ENS_CUDA(cudaEventCreate(&e0));
    TK_ENS_CUDA(cudaEventCreate(&e1));
    TK_ENS_CUDA(cudaEventRecord(e0, 0));
    tk::ensemble_kernel<<<(unsigned)grid, threads, smem>>>(P, d->n_sims);
    TK_ENS_CUDA(cudaGetLastError());
    TK_ENS_CUDA(cudaEventRecord(e1, 0));
    TK_ENS_CUDA(cudaMemcpy(probe_series, P.series, (size_t)d->n_sims * nt * sizeof(double), cudaMemcpyDeviceToHost));
    if (spectrum) TK_ENS_CUDA(cudaMemcpy(spectrum, P.spectrum, (size_t)d->n_sims * nt * sizeof(double), cudaMemcpyDeviceToHost));
    if (argmax) TK_ENS_CUDA(cudaMemcpy(argmax, P.argmax, (size_t)d->n_sims * sizeof(long long), cudaMemcpyDeviceToHost));
    if (elapsed_ms) {
        float ms = 0.f;
        TK_ENS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        *elapsed_ms = ms;
    }
    cleanup();
    return TK_OK;
#undef TK_ENS
#undef TK_ENS_CUDA
}

}  // extern "C"
