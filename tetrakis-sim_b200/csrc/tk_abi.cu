// C ABI of the tetrakis-b200 hot path (see include/tetrakis_b200.h for the contract).
// Host side: plan construction (graph -> sliced ELL; lattice descriptor -> defect tables),
// the stepping loops and all CUDA resource management.  No CPU compute fallback anywhere.
#include "../../include/tetrakis_b200.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "kernels_cross.cuh"
#include "kernels_ensemble.cuh"
#include "kernels_fft.cuh"
#include "kernels_fused.cuh"
#include "kernels_graph.cuh"
#include "kernels_lattice.cuh"

namespace {

thread_local std::string g_err;

// ---------------------------------------------------------------------------------------------------------------
// Device-memory cache.  A run_wave_sim call on a reference-sized graph makes ~20 cudaMalloc / cudaFree calls (plan
// arrays, probe and history staging, kick upload); on the B200 boxes of this pool a single one of them takes anything
// from 50 us to hundreds of ms (measured: scripts/micro/alloc_cost.cu), which is more than the whole simulation.  Freed
// blocks up to 256 MB are therefore kept per device (at most 2 GB in total) and handed out again to requests of about
// the same size; everything else goes straight to the driver.  tk_device_trim() returns the cache to the driver;
// TK_DEV_CACHE=0 switches it off.  A block is only reused after cudaDeviceSynchronize() at the time it was freed.
struct DevCache {
    std::mutex mu;
    std::unordered_map<void *, std::pair<int, size_t>> live;     // ptr -> (device, capacity)
    std::multimap<size_t, void *> free_[16];                     // per device: capacity -> ptr
    size_t cached[16] = {0};
    int enabled = -1;
};
DevCache &dev_cache()
{
    static DevCache *c = new DevCache();  // never destroyed: the CUDA context may be gone at exit
    return *c;
}
constexpr size_t kCacheMaxBlock = (size_t)256 << 20, kCacheMaxTotal = (size_t)2 << 30;

size_t cache_class(size_t bytes)
{
    if (bytes <= 65536) return (bytes + 511) / 512 * 512;
    size_t g = 4096;  // granule: 1/8 of the next power of two below the size
    while (g * 16 <= bytes) g <<= 1;
    return (bytes + g - 1) / g * g;
}

cudaError_t tk_malloc(void **p, size_t bytes)
{
    DevCache &c = dev_cache();
    if (c.enabled < 0) {
        const char *e = getenv("TK_DEV_CACHE");
        c.enabled = (e && *e == '0') ? 0 : 1;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    if (!c.enabled || dev < 0 || dev >= 16) return cudaMalloc(p, bytes);
    const size_t want = cache_class(std::max<size_t>(bytes, 1));
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.free_[dev].lower_bound(want);
        if (it != c.free_[dev].end() && it->first <= want + want / 4 + 65536) {
            *p = it->second;
            c.cached[dev] -= it->first;
            c.live[*p] = {dev, it->first};
            c.free_[dev].erase(it);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, want);
    if (e == cudaErrorMemoryAllocation) {  // give the cache back to the driver and try once more
        cudaGetLastError();
        std::vector<void *> drop;
        {
            std::lock_guard<std::mutex> lk(c.mu);
            for (auto &kv : c.free_[dev]) drop.push_back(kv.second);
            c.free_[dev].clear();
            c.cached[dev] = 0;
        }
        for (void *q : drop) cudaFree(q);
        e = cudaMalloc(p, want);
    }
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(c.mu);
        c.live[*p] = {dev, want};
    }
    return e;
}

cudaError_t tk_free(void *p)
{
    if (!p) return cudaSuccess;
    DevCache &c = dev_cache();
    int dev = -1;
    size_t cap = 0;
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.live.find(p);
        if (it != c.live.end()) {
            dev = it->second.first;
            cap = it->second.second;
            c.live.erase(it);
        }
    }
    if (dev < 0) return cudaFree(p);
    if (cap > kCacheMaxBlock) return cudaFree(p);
    cudaDeviceSynchronize();  // nobody still works on the block (cudaFree would have synchronised too)
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.cached[dev] + cap > kCacheMaxTotal) {
        cudaFree(p);
    } else {
        c.free_[dev].emplace(cap, p);
        c.cached[dev] += cap;
    }
    return cudaSuccess;
}

int cached_sm_count(int device)
{
    static int sm[16] = {0};
    if (device >= 0 && device < 16 && sm[device] > 0) return sm[device];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
        cudaGetLastError();
        n = 148;
    }
    if (device >= 0 && device < 16) sm[device] = n;
    return n;
}

// Sustained rates the pass-length heuristics are calibrated on, derived from the device's own attributes instead of
// literals measured on one box: FP64 instructions per second = 0.88 x SMs x 64 lanes x SM clock (the fused passes reach
// 88 % of the FP64 pipe: 16.4 us per step of 6.7e7 cells at 4 DFMA per cell on a 148-SM, 1 965 MHz B200), HBM bytes per
// second = 0.82 x 2 x memory clock x bus width (6.3 TB/s of the 7.67 TB/s a B200 reports: 3 996 MHz, 7 680 bits).  Falls back to those B200
// figures when an attribute is unavailable.
struct DevRates { double fp64_per_s, hbm_bytes_per_s; };

DevRates device_rates(int device)
{
    static DevRates cache[16] = {};
    if (device >= 0 && device < 16 && cache[device].fp64_per_s > 0.0) return cache[device];
    DevRates r{1.634e13, 6.3e12};
    int clk_khz = 0, mem_khz = 0, bus_bits = 0;
    const bool ok = cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, device) == cudaSuccess &&
                    cudaDeviceGetAttribute(&mem_khz, cudaDevAttrMemoryClockRate, device) == cudaSuccess &&
                    cudaDeviceGetAttribute(&bus_bits, cudaDevAttrGlobalMemoryBusWidth, device) == cudaSuccess;
    if (!ok) cudaGetLastError();
    if (ok && clk_khz > 0) r.fp64_per_s = 0.88 * cached_sm_count(device) * 64.0 * clk_khz * 1e3;
    if (ok && mem_khz > 0 && bus_bits > 0) r.hbm_bytes_per_s = 0.82 * 2.0 * mem_khz * 1e3 * (bus_bits / 8.0);
    if (device >= 0 && device < 16) cache[device] = r;
    return r;
}

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
    TK_CUDA(tk_malloc((void **)p, count * sizeof(T)));
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
    unsigned long long armed_signal = 0;     // != 0: the next full-range step launch signals this value itself
    unsigned int *sig_cnt = nullptr;         // [64] arrival counters of the early signal (own allocation: the flag
                                             // block is written by the neighbours over NVLink)
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

    // K-step passes on lattices with defects / 3-D lattices: the cross scheme (kernels_cross.cuh, host_cross.inc)
    std::vector<int> special_rows, special_cols;  // rows / columns holding a special cell in any layer (sorted)
    bool fx_whole = false;                   // the plan owns the whole lattice (no slab neighbours): eligible
    bool fx_ready = false;
    std::vector<long long> fx_probe_key;     // the probe set the cross was built for
    tk::FxDev fx{};
    std::vector<void *> fx_owned;            // device allocations of the current cross setup
    double *fxB[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [kind][level slot] bulk parts, [L][S][8]
    double *fxY[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [kind][level slot] responses, [L][S][4]
    double *fxTot[2] = {nullptr, nullptr};   // per-block totals over bulk lines, two generations
    int fx_bslot = 0, fx_yslot = 0, fx_tslot = 0;  // which slot is "current"
    int fx_k_run = 0;                        // steps per pass of the last run (0 = not used)
    long long fx_passes = 0;                 // passes since the bulk parts were last summed from the data
    double fx_cross_frac = 0.0;
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
    p->sm_count = cached_sm_count(device);  // (cudaGetDeviceProperties costs 2-15 ms per call here)
    return TK_OK;
}

// Uploads and zero-fills of plan-owned memory are ordered on the plan's own stream (a non-blocking stream does
// not synchronise with the legacy default stream, so cudaMemset / cudaMemcpy there could land after kernels the
// plan has already enqueued).  The host vector may be a temporary: wait for the copy.
template <typename T>
int plan_upload(tk_plan *p, T **dst, const std::vector<T> &v)
{
    int rc = dev_alloc(dst, v.size());
    if (rc) return rc;
    p->owned.push_back((void *)*dst);
    if (!v.empty()) {
        TK_CUDA(cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, p->stream));
        TK_CUDA(cudaStreamSynchronize(p->stream));
    }
    return TK_OK;
}

template <typename T>
int plan_alloc(tk_plan *p, T **dst, size_t count, bool zero)
{
    int rc = dev_alloc(dst, count);
    if (rc) return rc;
    p->owned.push_back((void *)*dst);
    if (zero) TK_CUDA(cudaMemsetAsync(*dst, 0, std::max<size_t>(count, 1) * sizeof(T), p->stream));
    return TK_OK;
}

int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// The host side is one translation unit split by subject:
#include "host_launch.inc"
#include "host_fused.inc"
#include "host_cross.inc"
#include "host_device.inc"
#include "host_plans.inc"
#include "host_run.inc"
#include "host_peer.inc"
#include "host_sheet.inc"
#include "host_ensemble.inc"
