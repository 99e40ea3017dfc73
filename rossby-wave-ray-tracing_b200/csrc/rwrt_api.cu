// rwrt_api.cu -- the C ABI of librwrt.so (include/rwrt.h): handle, buffers, launch sequencing.
#include <sched.h>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "../../include/rwrt.h"
#include "rwrt_internal.h"
#include "rwrt_rk45_params.h"
#include "rwrt_hx.h"

using namespace rwrt;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(RWRT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            cudaGetLastError();
            return -1;
        }
        cap = bytes;
        return 0;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace

namespace rwrt {
void set_error_string(const std::string &s) { g_err = s; }   // for the other translation units (rwrt_shsf.cu)
}  // namespace rwrt

struct rwrt_ctx {
    int device = 0, nlon = 0, nlat = 0, xcyclic = 1, nmember = 1, nlon_f = 0;
    int sm_count = 0;
    bool ready = false;
    double lon0 = 0, dlon = 0, lat0 = 0, dlat = 0;
    DevBuf fpk, scratch, trig, live, state, counters, blockcnt, inv, rows_done, records, packed, direct, row0;
    DevBuf arena;                      // inputs / counters / t = 0 row of the host-buffer calls; kept between calls: a cudaFree per
                                       // call stalled for ~0.5 s every few calls (rwrt_last_host_phases showed 280 -> 770 ms)
    cudaStream_t copy_stream = nullptr;   // D2H stream of the host-buffer calls, created on first use
    // per-call objects of the host-buffer calls, pooled on the handle (creating ~100 events and faulting in a fresh slot
    // map in every call showed up as occasional 70 ms stalls before the first window)
    std::vector<cudaEvent_t> hx_ev_ready, hx_ev_copied;   // per time window
    cudaEvent_t hx_ev_t0 = nullptr;
    int32_t *hx_inv = nullptr;            // pinned host copy of the slot -> ray map
    size_t hx_inv_cap = 0;
    void release_hx_pools()
    {
        for (auto e : hx_ev_ready) cudaEventDestroy(e);
        for (auto e : hx_ev_copied) cudaEventDestroy(e);
        hx_ev_ready.clear();
        hx_ev_copied.clear();
        if (hx_ev_t0) cudaEventDestroy(hx_ev_t0);
        hx_ev_t0 = nullptr;
        if (hx_inv) cudaFreeHost(hx_inv);
        hx_inv = nullptr;
        hx_inv_cap = 0;
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join = nullptr;
    std::vector<cudaEvent_t> ev_win;   // per time window: start / end of the integrate kernel (timing of the dominant kernel alone)
    int n_win_timed = 0;
    RunParams last_p;                  // parameters of the last run that kept its ray records on the handle (rwrt_ray_run_rk45_records)
    bool have_records = false;
    bool last_strict = false;
    cudaStream_t aux = nullptr;  // side stream for work that is independent of the integrator
    bool timed = false;
    double last_launches = 0, last_nlive = 0;
    // host-side expansion (rwrt_real2d_host): pinned staging buffers, kept between calls
    void *hx_stage[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t hx_stage_cap = 0;
    int hx_nstage = 0;
    double host_phases[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // last rwrt_real2d_host: host-clock ms at the end of each phase (rwrt_last_host_phases)
    std::vector<double> host_windows;        // last host-buffer call: RWRT_HOST_WINDOW_FIELDS doubles per time window (rwrt_last_host_windows)
    double host_stats[5] = {0, 0, 0, 0, 0};  // last host-buffer call: bytes D2H, bytes H2D, time windows, host threads, direct windows
    void release_hx_staging()
    {
        for (int i = 0; i < 4; ++i) {
            if (hx_stage[i]) cudaFreeHost(hx_stage[i]);
            hx_stage[i] = nullptr;
        }
        hx_stage_cap = 0;
        hx_nstage = 0;
    }

    Grid grid() const
    {
        Grid g;
        g.fpk = (const double *)fpk.p;
        g.member_stride = (size_t)nlon_f * nlat * NPK;
        g.nlon_f = nlon_f;
        g.nlat = nlat;
        g.lon0 = lon0;
        g.dlon = dlon;
        g.lat0 = lat0;
        g.dlat = dlat;
        g.rdlon = 1.0 / dlon;
        g.rdlat = 1.0 / dlat;
        return g;
    }
};

// ---------------------------------------------------------------------------------------------
// live-slot compaction (deterministic order)
// ---------------------------------------------------------------------------------------------
namespace rwrt {
namespace {

__device__ __forceinline__ bool slot_is_live(const double *out, size_t pl, long long slot)
{
    // a slot is integrated iff its t=0 state has no NaN: np.mean(y) is not NaN (rkf45.py:400-402)
    double s = out[slot] + out[pl + slot] + out[2 * pl + slot] + out[3 * pl + slot] + out[4 * pl + slot];
    return s == s;
}

// Work-list order: rays that share a root branch and a zonal wavenumber and start at neighbouring
// sources travel together, so the list walks (root, k, source) -- consecutive lanes then gather the
// same / adjacent field cells -- while slots stay in the reference's (root, source, k) order.
__device__ __forceinline__ long long list_to_slot(long long v, int nsource, int nzwn)
{
    if (nsource == 1 && nzwn == 1) return v;  // identity (also the RWRT_ORDER=slot switch)
    long long isrc = v % nsource, t = v / nsource;
    long long iz = t % nzwn, ir = t / nzwn;
    return (ir * nsource + isrc) * nzwn + iz;
}

__global__ void __launch_bounds__(256) count_live_kernel(long long N, size_t pl, const double *out, int nsource,
                                                         int nzwn, int *block_counts)
{
    long long v = (long long)blockIdx.x * 256 + threadIdx.x;
    bool live = v < N && slot_is_live(out, pl, list_to_slot(v, nsource, nzwn));
    int c = __syncthreads_count(live);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024) scan_blocks_kernel(int nblocks, int *block_counts, unsigned long long *nlive)
{
    // exclusive scan of block_counts in place by one block (nblocks <= a few 10^4)
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        int i = base + threadIdx.x;
        int v = (i < nblocks) ? block_counts[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warp_sums[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            warp_sums[threadIdx.x] = w;
        }
        __syncthreads();
        int warp_off = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0;
        int incl = x + warp_off;
        int c = carry;
        if (i < nblocks) block_counts[i] = c + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *nlive = (unsigned long long)carry;
}

__global__ void __launch_bounds__(256) scatter_live_kernel(long long N, size_t pl, const double *out, int nsource,
                                                           int nzwn, const int *block_offsets, int32_t *live,
                                                           int32_t *inv)
{
    __shared__ int warp_cnt[8];
    long long v = (long long)blockIdx.x * 256 + threadIdx.x;
    long long slot = v < N ? list_to_slot(v, nsource, nzwn) : 0;
    bool is_live = v < N && slot_is_live(out, pl, slot);
    unsigned m = __ballot_sync(0xffffffffu, is_live);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) warp_cnt[w] = __popc(m);
    __syncthreads();
    int off = block_offsets[blockIdx.x];
    for (int k = 0; k < w; ++k) off += warp_cnt[k];
    int idx = off + __popc(m & ((1u << l) - 1));
    if (is_live) live[idx] = (int32_t)slot;
    if (inv && v < N) inv[slot] = is_live ? idx : -1;
}

}  // namespace

void launch_compact_live(long long N, size_t pl, const double *out7, int nsource, int nzwn, int32_t *live, int32_t *inv,
                         unsigned long long *nlive_dev, int *block_counts, int nblocks, cudaStream_t s)
{
    count_live_kernel<<<nblocks, 256, 0, s>>>(N, pl, out7, nsource, nzwn, block_counts);
    scan_blocks_kernel<<<1, 1024, 0, s>>>(nblocks, block_counts, nlive_dev);
    scatter_live_kernel<<<nblocks, 256, 0, s>>>(N, pl, out7, nsource, nzwn, block_counts, live, inv);
}

// ---------------------------------------------------------------------------------------------
// pack_records: the ray records of one time window, [ray][rec_rows][8] as the integrator writes them, become seven
// planes over the LIVE rays, packed[plane][row][ray] -- the form that crosses the link in rwrt_real2d_host (56 bytes
// per live ray and row, no pad, rows of dead rays NaN like wr.py:838-855 leaves them).  Shared-memory transpose:
// a ray's rows are read 1 KB at a time, planes are written 256 bytes at a time.
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int PK_RAYS = 32, PK_ROWS = 16, PK_STRIDE = PK_ROWS * 8 + 1;   // odd stride: conflict-free column reads
__global__ void __launch_bounds__(256) pack_records_kernel(const double *rec, int rec_rows, const int32_t *rows_done,
                                                           long long nlive, int it_begin, int it_end, double *packed)
{
    __shared__ double tile[PK_RAYS * PK_STRIDE];
    const long long ray0 = (long long)blockIdx.x * PK_RAYS;
    const int row0 = it_begin + blockIdx.y * PK_ROWS;
    const int len = it_end - it_begin;
    const double n = __longlong_as_double(0x7ff8000000000000LL);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rr = lane >> 2, c = lane & 3;     // row within a pass of 8, 16-byte piece of the record
#pragma unroll
    for (int q = 0; q < PK_RAYS / 8; ++q) {
        const int s = warp + 8 * q;
        const long long li = ray0 + s;
        if (li >= nlive) break;
        const int rd = rows_done[li];
        const double2 *r2 = reinterpret_cast<const double2 *>(rec + ((size_t)li * rec_rows) * 8);
        double *t = tile + s * PK_STRIDE;
#pragma unroll
        for (int hh = 0; hh < PK_ROWS / 8; ++hh) {
            const int it = row0 + rr + 8 * hh;
            double2 v = make_double2(n, n);
            if (it < rd && it < it_end) v = __ldcs(r2 + (size_t)(it - it_begin) * 4 + c);
            t[(rr + 8 * hh) * 8 + 2 * c] = v.x;
            t[(rr + 8 * hh) * 8 + 2 * c + 1] = v.y;
        }
    }
    __syncthreads();
    const long long li = ray0 + lane;
#pragma unroll
    for (int j = 0; j < PK_ROWS / 8; ++j) {
        const int r = warp + 8 * j, it = row0 + r;
        if (it >= it_end) break;
        if (li < nlive) {
#pragma unroll
            for (int plane = 0; plane < 7; ++plane)
                __stcs(packed + ((size_t)plane * len + (it - it_begin)) * (size_t)nlive + li, tile[lane * PK_STRIDE + r * 8 + plane]);
        }
    }
}
}  // namespace

void launch_pack_records(const double *rec, int rec_rows, const int32_t *rows_done, long long nlive, int it_begin, int it_end,
                         double *packed, cudaStream_t s)
{
    const int len = it_end - it_begin;
    if (nlive <= 0 || len <= 0) return;
    pack_records_kernel<<<dim3((unsigned)((nlive + PK_RAYS - 1) / PK_RAYS), (len + PK_ROWS - 1) / PK_ROWS), 256, 0, s>>>(
        rec, rec_rows, rows_done, nlive, it_begin, it_end, packed);
}
}  // namespace rwrt

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int rwrt_abi_version(void) { return RWRT_ABI_VERSION; }

const char *rwrt_last_error(void) { return g_err.c_str(); }

int rwrt_device_count(int *count)
{
    if (!count) return fail(RWRT_ERR_ARG, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        cudaGetLastError();
        return fail(RWRT_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return RWRT_OK;
}

int rwrt_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return fail(RWRT_ERR_ARG, "ptr is NULL");
    CU(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return RWRT_OK;
}

int rwrt_host_free(void *ptr)
{
    if (ptr) CU(cudaFreeHost(ptr));
    return RWRT_OK;
}

int rwrt_create(int device, int nlon, int nlat, int xcyclic, int nmember, rwrt_handle *out)
{
    if (!out) return fail(RWRT_ERR_ARG, "out handle is NULL");
    *out = nullptr;
    if (nlon < 2 || nlat < 3 || nmember < 1) return fail(RWRT_ERR_ARG, "need nlon >= 2, nlat >= 3, nmember >= 1");
    // the basic-state kernels put the longitude index on gridDim.y and the ensemble member on gridDim.z
    if (nlon + 1 > 65535 || nmember > 65535)
        return fail(RWRT_ERR_ARG, "nlon (%d) and nmember (%d) must be below 65535 (grid dimensions of the stencil kernels): split the ensemble over handles", nlon, nmember);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(RWRT_ERR_CUDA, "no CUDA device visible: librwrt has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(RWRT_ERR_ARG, "device %d out of range (%d visible)", device, ndev);
    CU(cudaSetDevice(device));
    rwrt_ctx *h = new (std::nothrow) rwrt_ctx();
    if (!h) return fail(RWRT_ERR_NOMEM, "host allocation failed");
    h->device = device;
    h->nlon = nlon;
    h->nlat = nlat;
    h->xcyclic = xcyclic ? 1 : 0;
    h->nmember = nmember;
    h->nlon_f = nlon + h->xcyclic;
    {
        int sms = 0;
        cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (e != cudaSuccess) {
            delete h;
            return fail(RWRT_ERR_CUDA, "rwrt_create: %s", cudaGetErrorString(e));
        }
        h->sm_count = sms;
    }
    size_t npk = (size_t)nmember * h->nlon_f * nlat * NPK * sizeof(double);
    if (h->fpk.reserve(npk) || h->trig.reserve(2 * (size_t)nlat * sizeof(double)) ||
        h->counters.reserve(8 * sizeof(unsigned long long))) {
        delete h;
        return fail(RWRT_ERR_NOMEM, "device allocation of the packed fields (%zu B) failed", npk);
    }
    {
        int lo = 0, hi = 0;   // lo = numerically largest = lowest priority
        cudaError_t e = cudaEventCreate(&h->ev0);
        if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->aux, cudaStreamNonBlocking, lo);
        if (e != cudaSuccess) {
            rwrt_destroy(h);   // releases whatever exists so far
            return fail(RWRT_ERR_CUDA, "rwrt_create: %s", cudaGetErrorString(e));
        }
    }
    *out = h;
    return RWRT_OK;
}

int rwrt_destroy(rwrt_handle h)
{
    if (!h) return RWRT_OK;
    cudaSetDevice(h->device);
    h->fpk.release();
    h->inv.release();
    h->rows_done.release();
    h->records.release();
    h->packed.release();
    h->direct.release();
    h->row0.release();
    h->arena.release();
    h->release_hx_pools();
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (auto e : h->ev_win) cudaEventDestroy(e);
    h->ev_win.clear();
    h->scratch.release();
    h->trig.release();
    h->live.release();
    h->state.release();
    h->counters.release();
    h->blockcnt.release();
    h->release_hx_staging();
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->aux) cudaStreamDestroy(h->aux);
    delete h;
    return RWRT_OK;
}

static int set_coords(rwrt_handle h, const double *lon, const double *lat, cudaStream_t s)
{
    // the sampler needs lon[0], lon[1]-lon[0], lat[0], lat[1]-lat[0] (interpolation.py:78-79) on the host
    double c[4];
    CU(cudaMemcpyAsync(c, lon, 2 * sizeof(double), cudaMemcpyDefault, s));
    CU(cudaMemcpyAsync(c + 2, lat, 2 * sizeof(double), cudaMemcpyDefault, s));
    CU(cudaStreamSynchronize(s));
    h->lon0 = c[0];
    h->dlon = c[1] - c[0];
    h->lat0 = c[2];
    h->dlat = c[3] - c[2];
    return RWRT_OK;
}

int rwrt_bs_ready(rwrt_handle h, const float *u, const float *v, const double *lon, const double *lat,
                  const double *coslat, const double *sinlat, double *fields18, double *diag9, void *stream)
{
    if (!h || !u || !v || !lon || !lat) return fail(RWRT_ERR_ARG, "rwrt_bs_ready: NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaSetDevice(h->device));
    int rc = set_coords(h, lon, lat, s);
    if (rc) return rc;
    size_t n = (size_t)h->nmember * h->nlon * h->nlat;
    if (h->scratch.reserve(4 * n * sizeof(double))) return fail(RWRT_ERR_NOMEM, "scratch planes (%zu B)", 4 * n * 8);
    double *trig = (double *)h->trig.p;
    if (!coslat || !sinlat) {
        launch_trig_tables(h->nlat, lat, trig, trig + h->nlat, s);
        coslat = trig;
        sinlat = trig + h->nlat;
    }
    BsArgs a;
    a.nlon = h->nlon;
    a.nlat = h->nlat;
    a.nmember = h->nmember;
    a.xcyclic = h->xcyclic;
    a.u = u;
    a.v = v;
    a.coslat = coslat;
    a.sinlat = sinlat;
    a.q = (double *)h->scratch.p;
    a.qxx = a.q + n;
    a.qyy = a.q + 2 * n;
    a.qxy = a.q + 3 * n;
    a.fpk = (double *)h->fpk.p;
    a.fields18 = fields18;
    a.diag9 = diag9;
    launch_bs_ready(a, s);
    CU(cudaGetLastError());
    h->ready = true;
    return RWRT_OK;
}

int rwrt_bs_set_fields(rwrt_handle h, const double *fields18, const double *lon, const double *lat, void *stream)
{
    if (!h || !fields18 || !lon || !lat) return fail(RWRT_ERR_ARG, "rwrt_bs_set_fields: NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaSetDevice(h->device));
    int rc = set_coords(h, lon, lat, s);
    if (rc) return rc;
    launch_pack_from_fields18(h->nmember, h->nlon_f, h->nlat, fields18, (double *)h->fpk.p, s);
    CU(cudaGetLastError());
    h->ready = true;
    return RWRT_OK;
}

int rwrt_sample_mercator(rwrt_handle h, int64_t n, const double *lon, const double *lat, const int32_t *member,
                         double *out, void *stream)
{
    if (!h || (n > 0 && (!lon || !lat || !out))) return fail(RWRT_ERR_ARG, "rwrt_sample_mercator: NULL argument");
    if (!h->ready) return fail(RWRT_ERR_STATE, "rwrt_sample_mercator before rwrt_bs_ready");
    CU(cudaSetDevice(h->device));
    launch_sample_mercator(h->grid(), n, lon, lat, member, out, (cudaStream_t)stream);
    CU(cudaGetLastError());
    return RWRT_OK;
}

int rwrt_sample_mercator18(rwrt_handle h, const double *fields18, int64_t n, const double *lon, const double *lat,
                           const int32_t *member, double *out, void *stream)
{
    if (!h || !fields18 || (n > 0 && (!lon || !lat || !out)))
        return fail(RWRT_ERR_ARG, "rwrt_sample_mercator18: NULL argument");
    if (!h->ready) return fail(RWRT_ERR_STATE, "rwrt_sample_mercator18 before rwrt_bs_ready");
    CU(cudaSetDevice(h->device));
    launch_sample18(h->grid(), fields18, n, lon, lat, member, out, (cudaStream_t)stream);
    CU(cudaGetLastError());
    return RWRT_OK;
}

int rwrt_rhs(rwrt_handle h, int64_t n, const double *y, const int32_t *member, double *dy, uint32_t flags, void *stream)
{
    if (!h || (n > 0 && (!y || !dy))) return fail(RWRT_ERR_ARG, "rwrt_rhs: NULL argument");
    if (!h->ready) return fail(RWRT_ERR_STATE, "rwrt_rhs before rwrt_bs_ready");
    CU(cudaSetDevice(h->device));
    if (flags & RWRT_MODE_STRICT) launch_rhs_strict(h->grid(), n, y, member, dy, (cudaStream_t)stream);
    else launch_rhs_fast(h->grid(), n, y, member, dy, (cudaStream_t)stream);
    CU(cudaGetLastError());
    return RWRT_OK;
}

int rwrt_ray_init(rwrt_handle h, int nsource, const double *src_lon, const double *src_lat, const int32_t *src_member,
                  int nzwn, const double *zwn, double freq, int nt, double *out7, void *stream)
{
    if (!h || nsource < 0 || nzwn < 0 || nt < 1) return fail(RWRT_ERR_ARG, "rwrt_ray_init: bad sizes");
    if ((long long)nsource * nzwn > 0 && (!src_lon || !src_lat || !zwn || !out7))
        return fail(RWRT_ERR_ARG, "rwrt_ray_init: NULL argument");
    if (!h->ready) return fail(RWRT_ERR_STATE, "rwrt_ray_init before rwrt_bs_ready");
    CU(cudaSetDevice(h->device));
    launch_ray_init(h->grid(), nsource, src_lon, src_lat, src_member, nzwn, zwn, freq, nt, out7, (cudaStream_t)stream);
    CU(cudaGetLastError());
    return RWRT_OK;
}

// After each time window's kernel has been queued on `s`, hook(ctx, it_begin, it_end, s) may queue
// follow-up work (the host entry copies the finished rows to the host on another stream).
typedef int (*window_hook)(void *ctx, int it_begin, int it_end, cudaStream_t s);

// Host-side expansion (rwrt_real2d_host, rwrt_ray_run_rk45_d2h): the rows of every time window travel to the host as
// packed planes of the live rays instead of reference-layout rows (11.6 GB instead of 23.2 GB on C2) and host threads
// write the caller's arrays from them (details above hx_worker and hx_on_window).
constexpr int HX_STAGES = 4;
struct HostExpand {
    // set by the caller
    double *out7_host = nullptr;            // (7, nt, N)
    long long N = 0;
    int nt = 0;
    cudaStream_t copy_stream = nullptr;
    bool row0_only = true;                  // the device holds only the t = 0 row (7, N); false: full planes (7, nt, N) that
                                            // expand_records fills as well (rwrt_ray_run_rk45_d2h)
    int route_mode = 0;                     // 0 auto, 1 always through the host threads, 2 always direct, 3 alternate (tests)
    bool can_direct = false;                // out7_host is pinned: a window's expanded rows may be copied straight into it
    // filled by hx_begin
    long long nlive = 0;
    int rec_rows = 0, nwin = 0;
    const int32_t *inv = nullptr;           // slot -> live index (pinned host copy, owned by the handle)
    double *stage[HX_STAGES] = {};          // pinned (owned by the handle): packed rows of one window, 7 x rows x nlive doubles
    std::vector<cudaEvent_t> packed_ev, copied;     // per window
    std::vector<int> w_begin, w_end;
    std::vector<int> route, stage_of;       // per window: 1 = direct copy of device-expanded rows; staging buffer of a host window
    std::vector<int> hx_windows;            // the windows that went through the host threads, in order
    std::atomic<int> hx_finished{0};        // ... and how many of them are completely in the caller's arrays
    int direct_count = 0, direct_last[2] = {-1, -1};   // direct windows so far; last window that used direct buffer 0 / 1
    std::atomic<int> enqueued{0};           // windows whose D2H has been queued
    std::vector<std::atomic<int>> *chunks_done = nullptr;   // per window: finished chunks
    int nchunk = 0;
    long long chunk_slots = 0;
    std::vector<std::thread> workers;
    std::atomic<int> next_task{0};
    std::atomic<bool> abort{false};
    // timeline of the call (rwrt_last_host_windows): one clock origin for the device events and the host clock
    cudaEvent_t ev_t0 = nullptr;
    std::chrono::steady_clock::time_point t0;
    std::vector<double> bytes_w, t_enqueued;                 // per window: bytes on the link; host ms when its copy was queued
    std::vector<std::atomic<long long>> *t_first = nullptr;  // per window: host ns of the first / last chunk written by the threads
    std::vector<std::atomic<long long>> *t_last = nullptr;
    rwrt_handle owner = nullptr;
    std::mutex mu;                          // waits (window queued / window expanded) sleep on cv instead of spinning:
    std::condition_variable cv;             // the launching thread and the CUDA driver threads need the cores too
};

static int hx_on_window(HostExpand *hx, rwrt_handle h, int w, const RunParams &p, bool strict, const double *out7_dev, cudaStream_t s);
static int hx_begin(HostExpand *hx, rwrt_handle h, const double *out7_dev, size_t dev_pl, long long nlive, int rec_rows, int nwin,
                    cudaStream_t s);
#define RWRT_HX_UNSUPPORTED 12345   /* internal: the caller falls back to device-side expansion */

// Gridding request of a run (rwrt_ray_sweep_grid): accumulators (ncx, ncy) int64 each, cells of (dlon, dlat) degrees.
struct GridSpec {
    double dlon = 0, dlat = 0;
    int ncx = 0, ncy = 0;
    long long *count = nullptr, *fx = nullptr, *fy = nullptr;
    bool from_records = false;   // false: fused into the integrator (RED per output row); true: grid_records_kernel per window
};
// What a run keeps of its output rows
enum { KEEP_PLANES = 0,          // the seven reference-layout planes in out7 (7, nt, N)         -- rwrt_ray_run_rk45
       KEEP_RECORDS = 1,         // the per-ray records on the handle, out7 = row 0 only (7, N)  -- rwrt_ray_run_rk45_records
       KEEP_NOTHING = 2 };       // nothing (the grids are the result), out7 = row 0 only         -- rwrt_ray_sweep_grid
struct RunOpts {
    int nwin = 1;
    window_hook hook = nullptr;
    void *hook_ctx = nullptr;
    HostExpand *hx = nullptr;
    const GridSpec *grid = nullptr;
    int keep = KEEP_PLANES;
};

static cudaEvent_t *win_events(rwrt_handle h, int w)
{
    while ((int)h->ev_win.size() < 2 * (w + 1)) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        h->ev_win.push_back(e);
    }
    return h->ev_win.data() + 2 * w;
}

static int run_rk45(rwrt_handle h, int nsource, const int32_t *src_member, int nzwn, int nt, double tstep,
                    double ttotal, double rtol, double atol, double min_step_factor, double cut_off, uint32_t flags,
                    double *out7, int64_t *accepted, int64_t *attempts, int8_t *status, int64_t *total_accepted,
                    void *stream, const RunOpts &opt)
{
    int nwin = opt.nwin;
    window_hook hook = opt.hook;
    void *hook_ctx = opt.hook_ctx;
    HostExpand *hx = opt.hx;
    const GridSpec *grid = opt.grid;
    const bool row0_only = (hx && hx->row0_only) || opt.keep != KEEP_PLANES;
    const bool fused_grid = grid && !grid->from_records;
    if (!h || nsource < 0 || nzwn < 0 || nt < 1) return fail(RWRT_ERR_ARG, "rwrt_ray_run_rk45: bad sizes");
    if (!(tstep > 0)) return fail(RWRT_ERR_ARG, "rwrt_ray_run_rk45: tstep must be positive");
    if (!h->ready) return fail(RWRT_ERR_STATE, "rwrt_ray_run_rk45 before rwrt_bs_ready");
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaSetDevice(h->device));
    long long N = 3LL * nsource * nzwn;
    h->timed = false;
    h->last_launches = 0;
    h->last_nlive = 0;
    h->n_win_timed = 0;
    h->have_records = false;
    unsigned long long *ctr = (unsigned long long *)h->counters.p;  // [0] queue, [1] total accepted, [2] nlive
    CU(cudaMemsetAsync(ctr, 0, 8 * sizeof(unsigned long long), s));
    if (N == 0 || nt < 2) {
        if (total_accepted) CU(cudaMemsetAsync(total_accepted, 0, sizeof(int64_t), s));
        return RWRT_OK;
    }
    if (N > 2147483647LL) return fail(RWRT_ERR_ARG, "more than 2^31-1 ray slots in one call: shard the sources");
    if (!out7) return fail(RWRT_ERR_ARG, "rwrt_ray_run_rk45: out7 is NULL");

    // 1. compact the live slots
    bool use_records = false;
    int nblocks = (int)((N + 255) / 256);
    if (h->live.reserve((size_t)N * sizeof(int32_t)) || h->blockcnt.reserve((size_t)nblocks * sizeof(int)))
        return fail(RWRT_ERR_NOMEM, "work list allocation failed");
    {
        // default: work list in slot order (root, source, k); RWRT_ORDER=rks walks (root, k, source) instead
        const char *e = getenv("RWRT_ORDER");
        bool slot_order = !(e && e[0] == 'r');  // measured: slot order 2.54e9 vs 2.32e9 steps/s (output rows coalesce)
        // Record layout (default): the integrator writes 64-byte per-ray records and expand_records produces the
        // reference layout -- 13 GB of full-sector DRAM writes instead of 79 GB of read-modify-write, and the integrator
        // itself is 19 % faster (profiles/).  RWRT_RECORDS=0 writes the reference layout directly (v1-v13).
        const char *er = getenv("RWRT_RECORDS");
        use_records = !(er && atoi(er) == 0);
        if (opt.keep == KEEP_RECORDS || (grid && grid->from_records)) use_records = true;
        if (opt.keep == KEEP_NOTHING && fused_grid) use_records = false;   // no rows are stored at all
        if (use_records && h->inv.reserve((size_t)N * sizeof(int32_t))) return fail(RWRT_ERR_NOMEM, "slot map allocation failed");
        launch_compact_live(N, row0_only ? (size_t)N : (size_t)nt * N, out7, slot_order ? (int)1 : nsource, slot_order ? (int)1 : nzwn, (int32_t *)h->live.p,
                            use_records ? (int32_t *)h->inv.p : nullptr, ctr + 2, (int *)h->blockcnt.p, nblocks, s);
    }
    unsigned long long nlive = 0;
    CU(cudaMemcpyAsync(&nlive, ctr + 2, sizeof nlive, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));  // the one host sync of this call: grid sizes depend on nlive
    if (nlive && h->state.reserve((size_t)nlive * RWRT_STATE * sizeof(double)))
        return fail(RWRT_ERR_NOMEM, "ray state allocation failed (%llu rays)", nlive);

    // 2. parameters (rkf45.py:21-26 validate_tol; rkf45.py:362 Global_Minstep; wr.py:170 cut_off)
    RunParams p;
    p.g = h->grid();
    p.src_member = src_member;
    p.nsource = nsource;
    p.nzwn = nzwn;
    p.nt = nt;
    p.N = N;
    p.tstep = tstep;
    p.ttotal = ttotal;
    const double eps = 2.220446049250313e-16;
    p.rtol = rtol < 100 * eps ? 100 * eps : rtol;
    p.atol = atol;
    p.gmin = fmin(min_step_factor * tstep, (tstep - 0) * 0.001);
    p.cut = cut_off * tstep / 3600.;
    p.out = out7;
    p.out_pl = row0_only ? (size_t)N : (size_t)nt * N;
    p.dst = out7;                                   // expand_records writes the planes themselves ...
    p.dst_pl = (size_t)nt * N;                      // (... except for a direct window of rwrt_real2d_host, see hx_on_window)
    p.dst_it0 = 0;
    p.live = (const int32_t *)h->live.p;
    p.nlive = (long long)nlive;
    p.state = (double *)h->state.p;
    p.accepted = (long long *)accepted;
    p.attempts = (long long *)attempts;
    p.status = (signed char *)status;
    p.queue = ctr;
    p.total_accepted = ctr + 1;
    p.g_count = p.g_fx = p.g_fy = nullptr;
    p.g_dlon = p.g_dlat = 1.0;
    p.g_ncx = p.g_ncy = 0;
    p.no_rows = (opt.keep == KEEP_NOTHING && fused_grid) ? 1 : 0;
    if (fused_grid) {
        p.g_count = (unsigned long long *)grid->count;
        p.g_fx = (unsigned long long *)grid->fx;
        p.g_fy = (unsigned long long *)grid->fy;
        p.g_dlon = grid->dlon;
        p.g_dlat = grid->dlat;
        p.g_ncx = grid->ncx;
        p.g_ncy = grid->ncy;
    }

    // 3. launch: static slots, prepare, persistent integrator (one wave of resident blocks)
    bool strict = (flags & RWRT_MODE_STRICT) != 0;
    // The cooperative-gather kernel is the product path for every grid; RWRT_PER_LANE_GATHER=1 selects the per-lane
    // kernel instead (tests: two independent gather implementations must agree bit for bit).
    bool coop = true;
    if (const char *e = getenv("RWRT_PER_LANE_GATHER")) coop = (atoi(e) == 0);
    if ((long long)h->nmember * h->nlon_f * h->nlat >= (1LL << 30) - 1) coop = false;  // its fetch keys hold 30-bit point indices (103 GB of fields)
    p.coop = coop ? 1 : 0;
    if (!coop || nlive == 0) use_records = false;   // the record layout is a mode of the cooperative kernel
    if (hx && !use_records) return RWRT_HX_UNSUPPORTED;
    if (!coop && (grid || opt.keep != KEEP_PLANES))
        return fail(RWRT_ERR_ARG, "ray records / gridding sweeps need the cooperative-gather kernel (fields of < 2^30 points, "
                                  "RWRT_PER_LANE_GATHER unset)");
    p.rec = nullptr;
    p.rec_rows = 0;
    p.rec_it0 = 1;
    p.nquanta = 1;
    p.qshift = 0;
    p.rows_done = nullptr;
    p.inv = nullptr;
    int per_sm = strict ? occupancy_rk45_strict(coop) : occupancy_rk45_fast(coop);
    if (per_sm < 1) per_sm = 1;
    if (const char *e = getenv("RWRT_BLOCKS_PER_SM")) {  // tuning knob (profiles/): cap resident blocks per SM
        int v = atoi(e);
        if (v >= 1 && v < per_sm) per_sm = v;
    }
    const int bt = strict ? threads_rk45_strict(coop) : threads_rk45_fast(coop);   // threads per block of the integrator
    long long want = ((long long)nlive + bt - 1) / bt;
    int integrate_blocks = (int)(want < (long long)per_sm * h->sm_count ? want : (long long)per_sm * h->sm_count);
    if (integrate_blocks < 1) integrate_blocks = 1;
    int prepare_blocks = (int)(((long long)nlive + 127) / 128);
    int fill_blocks = (int)((N + 255) / 256);
    if (!coop || nwin < 1 || opt.keep == KEEP_RECORDS) nwin = 1;   // only the cooperative kernel can park rays at a window end
    if (nwin > nt - 1) nwin = nt - 1;
    if (use_records) {
        // the records of one time window must fit beside everything else: more windows when memory is short
        // host path: + two packed window buffers (7 doubles per live ray and row each) and, when windows may be copied
        // straight into pinned caller memory while only the t = 0 row lives on the device, two window buffers of expanded rows
        const bool direct_bufs = hx && hx->row0_only && hx->can_direct;
        const size_t per_row = (size_t)nlive * (8 + (hx ? 14 : 0)) * sizeof(double) + (direct_bufs ? (size_t)N * 14 * sizeof(double) : 0);
        size_t budget = h->records.cap;             // a buffer from an earlier call is reused as it is ...
        if (budget < per_row * (size_t)((nt - 1 + nwin - 1) / nwin + 1)) {   // ... unless it is too small (cudaMemGetInfo costs ms)
            size_t free_b = 0, total_b = 0;
            CU(cudaMemGetInfo(&free_b, &total_b));
            free_b += h->records.cap;
            budget = free_b > (3ull << 30) ? free_b - (2ull << 30) : free_b / 2;
        }
        int max_rows = (int)(budget / per_row < 2147483647ull ? budget / per_row : 2147483647ull);
        if (max_rows < 2) return fail(RWRT_ERR_NOMEM, "no room for the ray records (%llu rays): shard the sources", nlive);
        int need_win = (nt - 1 + (max_rows - 1) - 1) / (max_rows - 1);
        if (need_win > nwin && opt.keep == KEEP_RECORDS)
            return fail(RWRT_ERR_NOMEM, "the ray records of the whole run (%llu rays x %d rows x 64 B) do not fit on the device: "
                                        "shard the sources, or use rwrt_ray_run_rk45 / rwrt_ray_sweep_grid (time windows)", nlive, nt - 1);
        if (need_win > nwin) nwin = need_win;
        if (nwin > nt - 1) nwin = nt - 1;
        int rows = (nt - 1 + nwin - 1) / nwin + (hx ? 0 : 1);  // rows of the largest window (the host path copies whole buffers: no slack row)
        if (h->records.reserve((size_t)nlive * rows * 8 * sizeof(double)) || h->rows_done.reserve((size_t)nlive * sizeof(int32_t)) ||
            (hx && h->packed.reserve((size_t)nlive * rows * 14 * sizeof(double))) ||
            (direct_bufs && h->direct.reserve((size_t)N * rows * 14 * sizeof(double))))
            return fail(RWRT_ERR_NOMEM, "record buffer allocation failed (%llu rays x %d rows x 64 B)", nlive, rows);
        p.rec = (double *)h->records.p;
        p.rec_rows = rows;
        p.rows_done = (int32_t *)h->rows_done.p;
        p.inv = (const int32_t *)h->inv.p;
    }
    if (coop && nlive) {   // progress flags of the in-kernel time quanta (also: number of valid rows per ray, for the record consumers)
        if (h->rows_done.reserve((size_t)nlive * sizeof(int32_t))) return fail(RWRT_ERR_NOMEM, "ray progress flags (%llu rays)", nlive);
        p.rows_done = (int32_t *)h->rows_done.p;
    }
    // Time quanta inside a launch: a ray is integrated a quantum of output rows at a time and re-queued, so that the persistent
    // grid ends with a tail of one (short) quantum instead of one whole ray -- and rays differ a lot in length.  The quanta
    // shrink geometrically (by 2^RWRT_QUANTUM_SHIFT, default 4x) down to RWRT_QUANTUM_MIN rows.  RWRT_QUANTUM_SHIFT=0: off.
    int qshift = 2, qmin = 8;
    if (const char *e = getenv("RWRT_QUANTUM_SHIFT")) qshift = atoi(e);
    if (const char *e = getenv("RWRT_QUANTUM_MIN")) qmin = atoi(e);
    if (qshift < 0 || qshift > 8) qshift = 0;
    if (qmin < 1) qmin = 1;
    // window bounds: rows [bounds[w], bounds[w + 1]) belong to window w.  Host delivery: the first and the last window
    // are split (1/4, 1/4, 1/2) so that the pipeline fills and drains on quarter-size windows -- the host threads start
    // after a quarter of a window has been integrated, packed and copied, and finish a quarter window after the last copy.
    std::vector<int> bounds(nwin + 1);
    for (int w = 0; w <= nwin; ++w) bounds[w] = 1 + (int)((long long)(nt - 1) * w / nwin);
    if (hx && use_records) {
        bool ramp = true;
        if (const char *e = getenv("RWRT_HOST_RAMP")) ramp = atoi(e) != 0;
        if (ramp && nwin >= 2) {
            std::vector<int> b2;
            auto split = [&](int b, int e, bool head) {
                const int len = e - b, q = len / 4;
                if (q < 1) { b2.push_back(b); return; }
                if (head) { b2.push_back(b); b2.push_back(b + q); b2.push_back(b + 2 * q); }
                else { b2.push_back(b); b2.push_back(e - 2 * q); b2.push_back(e - q); }
            };
            split(bounds[0], bounds[1], true);
            for (int w = 1; w + 1 < nwin; ++w) b2.push_back(bounds[w]);
            split(bounds[nwin - 1], bounds[nwin], false);
            b2.push_back(bounds[nwin]);
            bounds.swap(b2);
            nwin = (int)bounds.size() - 1;
        }
        int rc = hx_begin(hx, h, out7, p.out_pl, (long long)nlive, p.rec_rows, nwin, s);
        if (rc) return rc;
    }
    CU(cudaEventRecord(h->ev0, s));
    CU(cudaEventRecord(h->ev_fork, s));            // fork: the side stream starts after everything queued so far
    CU(cudaStreamWaitEvent(h->aux, h->ev_fork, 0));
    p.it_begin = 1;
    p.it_end = nt;
    if (strict) launch_rk45_setup_strict(p, prepare_blocks, fill_blocks, s, h->aux);
    else launch_rk45_setup_fast(p, prepare_blocks, fill_blocks, s, h->aux);
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev_join, h->aux));
    CU(cudaStreamWaitEvent(s, h->ev_join, 0));     // join: the pre-fill (all rows >= 1) precedes the integrator's writes
    int launches = 0;
    for (int w = 0; w < nwin; ++w) {
        p.it_begin = bounds[w];
        p.it_end = bounds[w + 1];
        p.rec_it0 = p.it_begin;                     // record row r of this window is output row it_begin + r
        p.qshift = qshift;
        p.nquanta = 1;                              // quanta off: one work item per ray and window
        if (qshift > 0)                             // quantum q + 1 exists while the rows left after quantum q are >= qmin
            while (p.nquanta < 15 && qshift * p.nquanta < 31 && ((p.it_end - p.it_begin) >> (qshift * p.nquanta)) >= qmin) ++p.nquanta;
        if (w > 0) CU(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), s));  // rewind the work queue
        cudaEvent_t *we = win_events(h, w);
        if (we) CU(cudaEventRecord(we[0], s));
        if (strict) launch_rk45_window_strict(p, integrate_blocks, s);
        else launch_rk45_window_fast(p, integrate_blocks, s);
        CU(cudaGetLastError());
        if (we) {
            CU(cudaEventRecord(we[1], s));
            h->n_win_timed = w + 1;
        }
        ++launches;
        if (grid && grid->from_records && use_records) {   // the window's records -> ray-density / flux grids
            launch_grid_records(p.rec, p.rec_rows, p.rows_done, p.nlive, p.it_begin, p.it_end, grid->dlon, grid->dlat, grid->ncx,
                                grid->ncy, grid->count, grid->fx, grid->fy, h->sm_count, s);
            CU(cudaGetLastError());
            ++launches;
        }
        if (use_records && !row0_only) {   // rows [it_begin, it_end) of the reference layout from the records
            if (strict) launch_expand_records_strict(p, s);
            else launch_expand_records_fast(p, s);
            CU(cudaGetLastError());
            ++launches;
        }
        if (hx) {                                   // rows [it_begin, it_end) -> host (packed planes of the live rays, or direct)
            int rc = hx_on_window(hx, h, w, p, strict, out7, s);
            if (rc) return rc;
        }
        if (hook) {
            int rc = hook(hook_ctx, p.it_begin, p.it_end, s);
            if (rc) return rc;
        }
    }
    CU(cudaEventRecord(h->ev1, s));
    h->timed = true;
    h->last_launches = 3 + 1 + (nlive ? 1 + launches : 0);  // compaction (3) + fill + prepare + integrate windows
    h->last_nlive = (double)nlive;
    if (total_accepted)
        CU(cudaMemcpyAsync(total_accepted, ctr + 1, sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
    if (opt.keep == KEEP_RECORDS && use_records) {
        h->last_p = p;
        h->last_p.it_begin = 1;
        h->last_p.it_end = nt;
        h->last_p.accepted = nullptr;                // caller memory: not touched again after this call
        h->last_p.attempts = nullptr;
        h->last_p.status = nullptr;
        h->last_strict = strict;
        h->have_records = true;
    }
    return RWRT_OK;
}


// ---------------------------------------------------------------------------------------------
// Host-side expansion (rwrt_real2d_host, rwrt_ray_run_rk45_d2h).
//
// The link to the host is the bottleneck of the host-buffer calls (23.2 GB of reference-layout rows per C2 pass at
// 30-55 GB/s).  Half of those bytes are redundant: slots without a root are constant / NaN in every row.  So after
// each time window pack_records turns the ray records into seven planes over the live rays (56 B per live ray and
// row: 11.6 GB per C2 pass), those cross the link into pinned staging buffers while the next window integrates, and
// host threads write the caller's seven arrays from them with non-temporal stores (rwrt_hx.cpp).  Two packed buffers
// alternate on the device, HX_STAGES staging buffers on the host.
// ---------------------------------------------------------------------------------------------

static void hx_worker(HostExpand *hx, int device)
{
    cudaSetDevice(device);
    const int total = hx->nwin * hx->nchunk;
    for (;;) {
        const int task = hx->next_task.fetch_add(1);
        if (task >= total || hx->abort.load()) return;
        const int w = task / hx->nchunk, c = task % hx->nchunk;
        if (hx->enqueued.load(std::memory_order_acquire) <= w) {        // D2H of window w not queued yet
            std::unique_lock<std::mutex> lk(hx->mu);
            hx->cv.wait(lk, [&] { return hx->enqueued.load(std::memory_order_acquire) > w || hx->abort.load(); });
        }
        if (hx->abort.load()) return;
        if (hx->route[w]) continue;                                     // a direct window: the DMA engine writes it
        if (cudaEventSynchronize(hx->copied[w]) != cudaSuccess) {
            {
                std::lock_guard<std::mutex> lk(hx->mu);
                hx->abort.store(true);
            }
            hx->cv.notify_all();
            return;
        }
        const long long s0 = (long long)c * hx->chunk_slots;
        const long long s1 = s0 + hx->chunk_slots < hx->N ? s0 + hx->chunk_slots : hx->N;
        {
            const long long now = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - hx->t0).count();
            long long cur = (*hx->t_first)[w].load(std::memory_order_relaxed);
            while ((cur < 0 || now < cur) && !(*hx->t_first)[w].compare_exchange_weak(cur, now, std::memory_order_relaxed)) {}
        }
        HxView v;
        v.out = hx->out7_host;
        v.pl = (size_t)hx->nt * hx->N;
        v.N = hx->N;
        v.inv = hx->inv;
        v.packed = hx->stage[hx->stage_of[w]];
        v.nray = hx->nlive;
        v.rows = hx->w_end[w] - hx->w_begin[w];
        hx_expand_rows(v, hx->w_begin[w], s0, s1);
        if ((*hx->chunks_done)[w].fetch_add(1, std::memory_order_acq_rel) + 1 == hx->nchunk) {   // window w is in the caller's arrays
            (*hx->t_last)[w].store(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - hx->t0).count(),
                                   std::memory_order_relaxed);
            std::lock_guard<std::mutex> lk(hx->mu);
            hx->hx_finished.fetch_add(1, std::memory_order_release);
            hx->cv.notify_all();
        }
    }
}

static int hx_begin(HostExpand *hx, rwrt_handle h, const double *out7_dev, size_t dev_pl, long long nlive, int rec_rows, int nwin,
                    cudaStream_t s)
{
    hx->nlive = nlive;
    hx->rec_rows = rec_rows;
    hx->nwin = nwin;
    if (h->hx_inv_cap < (size_t)hx->N) {
        if (h->hx_inv) cudaFreeHost(h->hx_inv);
        h->hx_inv = nullptr;
        h->hx_inv_cap = 0;
        if (cudaMallocHost((void **)&h->hx_inv, (size_t)hx->N * sizeof(int32_t)) != cudaSuccess) {
            cudaGetLastError();
            return fail(RWRT_ERR_NOMEM, "pinned slot map (%lld slots)", hx->N);
        }
        h->hx_inv_cap = (size_t)hx->N;
    }
    hx->inv = h->hx_inv;
    CU(cudaMemcpyAsync(h->hx_inv, h->inv.p, (size_t)hx->N * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    // row 0 (the t = 0 state, wr.py:768-781) goes to the caller's planes now: the workers read the constant values of
    // NaN-root slots from it
    CU(cudaMemcpy2DAsync(hx->out7_host, (size_t)hx->nt * hx->N * sizeof(double), out7_dev, dev_pl * sizeof(double),
                         (size_t)hx->N * sizeof(double), 7, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    // pinned staging buffers live on the handle: pinning 2 GB costs more than a whole C2 pass
    const size_t stage_bytes = (size_t)nlive * rec_rows * 7 * sizeof(double);
    const int nstage = nwin < HX_STAGES ? nwin : HX_STAGES;
    if (h->hx_stage_cap < stage_bytes || h->hx_nstage < nstage) {
        h->release_hx_staging();
        for (int i = 0; i < nstage; ++i) {
            if (cudaMallocHost((void **)&h->hx_stage[i], stage_bytes) != cudaSuccess) {
                cudaGetLastError();
                h->release_hx_staging();
                return fail(RWRT_ERR_NOMEM, "pinned staging allocation failed (%d x %zu bytes)", nstage, stage_bytes);
            }
        }
        h->hx_stage_cap = stage_bytes;
        h->hx_nstage = nstage;
    }
    for (int i = 0; i < HX_STAGES; ++i) hx->stage[i] = (double *)h->hx_stage[i];
    hx->packed_ev.assign(nwin, nullptr);
    hx->copied.assign(nwin, nullptr);
    hx->w_begin.assign(nwin, 0);
    hx->w_end.assign(nwin, 0);
    hx->route.assign(nwin, 0);
    hx->stage_of.assign(nwin, 0);
    hx->hx_windows.clear();
    while ((int)h->hx_ev_ready.size() < nwin) {    // timing enabled: the per-window timeline of rwrt_last_host_windows
        cudaEvent_t a = nullptr, b = nullptr;
        CU(cudaEventCreate(&a));
        h->hx_ev_ready.push_back(a);
        CU(cudaEventCreateWithFlags(&b, cudaEventBlockingSync));
        h->hx_ev_copied.push_back(b);
    }
    for (int w = 0; w < nwin; ++w) {
        hx->packed_ev[w] = h->hx_ev_ready[w];
        hx->copied[w] = h->hx_ev_copied[w];
    }
    hx->chunks_done = new std::vector<std::atomic<int>>(nwin);
    for (auto &a : *hx->chunks_done) a.store(0);
    hx->t_first = new std::vector<std::atomic<long long>>(nwin);
    hx->t_last = new std::vector<std::atomic<long long>>(nwin);
    for (auto &a : *hx->t_first) a.store(-1);
    for (auto &a : *hx->t_last) a.store(-1);
    hx->bytes_w.assign(nwin, 0.0);
    hx->t_enqueued.assign(nwin, 0.0);
    hx->owner = h;
    if (!h->hx_ev_t0) CU(cudaEventCreate(&h->hx_ev_t0));
    hx->ev_t0 = h->hx_ev_t0;
    CU(cudaEventRecord(hx->ev_t0, s));           // the stream is idle here (synchronised above): device and host clocks start together
    hx->t0 = std::chrono::steady_clock::now();
    hx->chunk_slots = HX_CHUNK;                 // multiple of 8: 64-byte lines stay whole when N is a multiple of 8
    hx->nchunk = (int)((hx->N + hx->chunk_slots - 1) / hx->chunk_slots);
    // worker threads: the cores this process may run on, shared between the ranks of a multi-GPU launch
    int nthreads = (int)std::thread::hardware_concurrency();
    cpu_set_t cpus;
    if (sched_getaffinity(0, sizeof cpus, &cpus) == 0 && CPU_COUNT(&cpus) > 0) nthreads = CPU_COUNT(&cpus);
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) {       // torchrun: ranks on this host
        const int ranks = atoi(e);
        // ranks share the host's memory bandwidth: two thirds of a rank's cores write as fast as all of them and
        // leave room for the launching and driver threads (2 ranks on 24 cores: 8 threads 686 ms per C2 pass, 12: 766)
        if (ranks > 1) nthreads = (2 * nthreads / ranks + 2) / 3;
        else if (nthreads > 3) nthreads -= 2;
    }
    else if (nthreads > 3) nthreads -= 2;                   // one rank: leave two cores to the launching thread and the CUDA driver's threads
    if (nthreads > 32) nthreads = 32;                       // the expansion is bound by memory bandwidth long before
    if (const char *e = getenv("RWRT_HOST_THREADS")) nthreads = atoi(e);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    try {
        for (int i = 0; i < nthreads; ++i) hx->workers.emplace_back(hx_worker, hx, h->device);
    } catch (...) {     // thread limit of the process: go on with the threads that exist
        if (hx->workers.empty()) return fail(RWRT_ERR_NOMEM, "could not start a host thread for the result expansion");
        nthreads = (int)hx->workers.size();
    }
    h->host_phases[7] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - hx->t0).count();  // thread start-up
    h->host_stats[0] += (double)hx->N * (sizeof(int32_t) + 7 * sizeof(double));   // slot map + the t = 0 row; windows add theirs
    h->host_stats[2] = nwin;
    h->host_stats[3] = nthreads;
    return RWRT_OK;
}

// Window w has been integrated (and, when the device holds the full planes, expanded).  Two ways to the host:
//   host route   pack_records -> one copy of the packed planes of the live rays into a pinned staging buffer -> the
//                worker threads write the caller's planes (half the link bytes, costs host cores and memory bandwidth);
//   direct route the expanded rows of the window are copied straight into the caller's (pinned) planes by the DMA engine
//                (twice the link bytes, no host work).
// Default: the host route, and the direct route whenever all HX_STAGES staging buffers are still waiting for the host
// threads -- i.e. when the cores, not the link, are what the delivery is waiting for.  Decisions are taken with at most
// two copies outstanding, so that they see the current backlog.
static int hx_on_window(HostExpand *hx, rwrt_handle h, int w, const RunParams &p, bool strict, const double *out7_dev, cudaStream_t s)
{
    const int it_begin = p.it_begin, it_end = p.it_end, len = it_end - it_begin;
    const long long nlive = p.nlive, N = hx->N;
    hx->w_begin[w] = it_begin;
    hx->w_end[w] = it_end;
    // records -> packed planes of the live rays (device buffer w & 1: last read by the copy of window w - 2); queued
    // before the route is known so that the link never waits for it (0.15 ms of GPU time if the window goes direct)
    double *packed = (double *)h->packed.p + (size_t)(w & 1) * (size_t)nlive * p.rec_rows * 7;
    if (w >= 2) CU(cudaStreamWaitEvent(s, hx->copied[w - 2], 0));
    launch_pack_records(p.rec, p.rec_rows, p.rows_done, nlive, it_begin, it_end, packed, s);
    CU(cudaGetLastError());
    CU(cudaEventRecord(hx->packed_ev[w], s));
    if (w >= 2) CU(cudaEventSynchronize(hx->copied[w - 2]));           // at most two copies in flight when the route is chosen
    const int nhx = (int)hx->hx_windows.size();
    bool direct = false;
    if (hx->can_direct) {
        const int backlog = nhx - hx->hx_finished.load(std::memory_order_acquire);
        direct = hx->route_mode == 2 || (hx->route_mode == 3 && (w & 1)) || (hx->route_mode == 0 && backlog >= HX_STAGES);
    }
    size_t bytes;
    if (!direct) {
        if (nhx >= HX_STAGES) {                  // the staging buffer is reused: wait for its previous window on the host
            const int prev = hx->hx_windows[nhx - HX_STAGES];
            std::unique_lock<std::mutex> lk(hx->mu);
            hx->cv.wait(lk, [&] { return (*hx->chunks_done)[prev].load(std::memory_order_acquire) >= hx->nchunk || hx->abort.load(); });
            if (hx->abort.load()) return fail(RWRT_ERR_CUDA, "host expansion aborted");
        }
        hx->stage_of[w] = nhx % HX_STAGES;
        hx->hx_windows.push_back(w);
        bytes = (size_t)nlive * (size_t)len * 7 * sizeof(double);
        CU(cudaStreamWaitEvent(hx->copy_stream, hx->packed_ev[w], 0));
        CU(cudaMemcpyAsync(hx->stage[hx->stage_of[w]], packed, bytes, cudaMemcpyDeviceToHost, hx->copy_stream));
    } else {
        const double *src = out7_dev;            // full planes on the device: rows [it_begin, it_end) are final after expand_records
        size_t src_pl = (size_t)hx->nt * N, src_row0 = (size_t)it_begin;
        cudaEvent_t ready = hx->packed_ev[w];    // recorded after expand_records of this window (stream order)
        if (hx->row0_only) {                     // only the t = 0 row lives on the device: expand this window into a buffer of its own
            const int b = hx->direct_count & 1;
            double *buf = (double *)h->direct.p + (size_t)b * (size_t)N * p.rec_rows * 7;
            if (hx->direct_last[b] >= 0) CU(cudaStreamWaitEvent(s, hx->copied[hx->direct_last[b]], 0));
            RunParams q = p;
            q.dst = buf;
            q.dst_pl = (size_t)N * len;
            q.dst_it0 = it_begin;
            if (strict) launch_expand_records_strict(q, s);
            else launch_expand_records_fast(q, s);
            CU(cudaGetLastError());
            CU(cudaEventRecord(hx->packed_ev[w], s));                   // re-recorded: now also after the expansion
            hx->direct_last[b] = w;
            ++hx->direct_count;
            src = buf;
            src_pl = (size_t)N * len;
            src_row0 = 0;
        }
        CU(cudaStreamWaitEvent(hx->copy_stream, ready, 0));
        bytes = (size_t)N * (size_t)len * 7 * sizeof(double);
        for (int pl = 0; pl < 7; ++pl)
            CU(cudaMemcpyAsync(hx->out7_host + ((size_t)pl * hx->nt + it_begin) * (size_t)N, src + pl * src_pl + src_row0 * (size_t)N,
                               (size_t)N * len * sizeof(double), cudaMemcpyDeviceToHost, hx->copy_stream));
        hx->route[w] = 1;
        h->host_stats[4] += 1;
    }
    CU(cudaEventRecord(hx->copied[w], hx->copy_stream));
    hx->bytes_w[w] = (double)bytes;
    hx->t_enqueued[w] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - hx->t0).count();
    h->host_stats[0] += (double)bytes;
    {
        std::lock_guard<std::mutex> lk(hx->mu);
        hx->enqueued.store(w + 1, std::memory_order_release);
    }
    hx->cv.notify_all();
    return RWRT_OK;
}

// joins the workers (after every window has been expanded when ok, at once otherwise) and frees the per-call objects
static int hx_finish(HostExpand *hx, bool ok)
{
    if (!ok) {
        std::lock_guard<std::mutex> lk(hx->mu);
        hx->abort.store(true);
    }
    hx->cv.notify_all();
    for (auto &t : hx->workers) t.join();
    hx->workers.clear();
    // timeline of the call, one row per window (all events have completed: the caller synchronised both streams)
    if (ok && hx->owner && hx->ev_t0 && !hx->abort.load()) {
        std::vector<double> &tl = hx->owner->host_windows;
        tl.assign((size_t)hx->nwin * RWRT_HOST_WINDOW_FIELDS, 0.0);
        float prev_end = 0;
        for (int w = 0; w < hx->nwin; ++w) {
            float t_ready = 0, t_end = 0;
            if (cudaEventElapsedTime(&t_ready, hx->ev_t0, hx->packed_ev[w]) != cudaSuccess ||
                cudaEventElapsedTime(&t_end, hx->ev_t0, hx->copied[w]) != cudaSuccess) {
                cudaGetLastError();
                tl.clear();
                break;
            }
            double *r = tl.data() + (size_t)w * RWRT_HOST_WINDOW_FIELDS;
            r[0] = hx->route[w];
            r[1] = hx->bytes_w[w];
            r[2] = t_ready;                                      // device: rows of the window ready to travel
            r[3] = hx->t_enqueued[w];                            // host: copy queued (after the routing decision)
            r[4] = t_ready > prev_end ? t_ready : prev_end;      // device: the link is free and the data ready (copies are serial) ...
            if (r[3] > r[4]) r[4] = r[3];                        // ... and the copy has been queued
            r[5] = t_end;                                        // device: copy finished
            r[6] = (*hx->t_first)[w].load() < 0 ? -1.0 : (*hx->t_first)[w].load() * 1e-6;   // host threads: first chunk started
            r[7] = (*hx->t_last)[w].load() < 0 ? -1.0 : (*hx->t_last)[w].load() * 1e-6;     // ... last chunk written
            prev_end = t_end;
        }
    }
    hx->ev_t0 = nullptr;                    // the events belong to the handle's pools
    delete hx->t_first;
    delete hx->t_last;
    hx->t_first = hx->t_last = nullptr;
    hx->packed_ev.clear();
    hx->copied.clear();
    delete hx->chunks_done;
    hx->chunks_done = nullptr;
    return hx->abort.load() ? RWRT_ERR_CUDA : RWRT_OK;
}

// Host half of rwrt_real2d_host on its own (tests, and callers that move the packed rows themselves): rows [row_begin,
// row_end) of slots [slot_begin, slot_end) of the seven planes from the packed rows of one time window.  Pure host code.
int rwrt_host_expand_rows(double *out7_host, int nt, int64_t nslot, const int32_t *slot_to_ray, const double *packed,
                          int64_t nray, int row_begin, int row_end, int64_t slot_begin, int64_t slot_end)
{
    if (!out7_host || !slot_to_ray || !packed || nt < 1 || nslot < 0 || nray < 1 || row_begin < 1 || row_end > nt ||
        slot_begin < 0 || slot_end > nslot)
        return fail(RWRT_ERR_ARG, "rwrt_host_expand_rows: bad argument");
    HxView v;
    v.out = out7_host;
    v.pl = (size_t)nt * (size_t)nslot;
    v.N = nslot;
    v.inv = slot_to_ray;
    v.packed = packed;
    v.nray = nray;
    v.rows = row_end - row_begin;
    if (v.rows > 0)
        for (int64_t s0 = slot_begin; s0 < slot_end; s0 += HX_CHUNK)
            hx_expand_rows(v, row_begin, s0, s0 + HX_CHUNK < slot_end ? s0 + HX_CHUNK : slot_end);
    return RWRT_OK;
}

int rwrt_ray_run_rk45(rwrt_handle h, int nsource, const int32_t *src_member, int nzwn, int nt, double tstep,
                      double ttotal, double rtol, double atol, double min_step_factor, double cut_off, uint32_t flags,
                      double *out7, int64_t *accepted, int64_t *attempts, int8_t *status, int64_t *total_accepted,
                      void *stream)
{
    int nwin = 1;
    if (const char *e = getenv("RWRT_WINDOWS")) nwin = atoi(e);  // tuning / testing: split the run into time windows
    RunOpts o;
    o.nwin = nwin;
    return run_rk45(h, nsource, src_member, nzwn, nt, tstep, ttotal, rtol, atol, min_step_factor, cut_off, flags, out7,
                    accepted, attempts, status, total_accepted, stream, o);
}

int rwrt_trim(rwrt_handle h)
{
    if (!h) return fail(RWRT_ERR_ARG, "rwrt_trim: null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    h->records.release();
    h->packed.release();
    h->direct.release();
    h->row0.release();
    h->arena.release();
    h->release_hx_pools();
    h->have_records = false;
    h->rows_done.release();
    h->inv.release();
    h->state.release();
    h->live.release();
    h->blockcnt.release();
    h->release_hx_staging();
    return RWRT_OK;
}

int rwrt_last_run_stats(rwrt_handle h, double *stats3)
{
    if (!h || !stats3) return fail(RWRT_ERR_ARG, "rwrt_last_run_stats: NULL argument");
    stats3[0] = h->last_launches;
    stats3[1] = h->last_nlive;
    stats3[2] = 0;
    if (h->timed) {
        CU(cudaSetDevice(h->device));
        CU(cudaEventSynchronize(h->ev1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        stats3[2] = ms;
        unsigned long long lost = 0;     // counters[3]: the integrator's watchdog (a ray's progress flag never arrived)
        CU(cudaMemcpy(&lost, (const unsigned long long *)h->counters.p + 3, sizeof lost, cudaMemcpyDeviceToHost));
        if (lost) return fail(RWRT_ERR_CUDA, "rwrt_last_run_stats: the integrator gave up waiting for a ray's previous time quantum; the results of this run are incomplete");
    }
    return RWRT_OK;
}

int rwrt_last_run_times(rwrt_handle h, double *ms4)
{
    if (!h || !ms4) return fail(RWRT_ERR_ARG, "rwrt_last_run_times: NULL argument");
    ms4[0] = ms4[1] = ms4[2] = ms4[3] = 0;
    if (!h->timed) return RWRT_OK;
    CU(cudaSetDevice(h->device));
    CU(cudaEventSynchronize(h->ev1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    ms4[0] = ms;
    for (int w = 0; w < h->n_win_timed; ++w) {
        CU(cudaEventElapsedTime(&ms, h->ev_win[2 * w], h->ev_win[2 * w + 1]));
        ms4[1] += ms;
    }
    if (h->n_win_timed > 0) {
        CU(cudaEventElapsedTime(&ms, h->ev0, h->ev_win[0]));
        ms4[2] = ms;
    }
    ms4[3] = ms4[0] - ms4[1] - ms4[2];
    return RWRT_OK;
}

int rwrt_last_host_stats(rwrt_handle h, double *stats5)
{
    if (!h || !stats5) return fail(RWRT_ERR_ARG, "rwrt_last_host_stats: NULL argument");
    for (int i = 0; i < 5; ++i) stats5[i] = h->host_stats[i];
    return RWRT_OK;
}

int rwrt_last_host_phases(rwrt_handle h, double *ms8)
{
    if (!h || !ms8) return fail(RWRT_ERR_ARG, "rwrt_last_host_phases: NULL argument");
    for (int i = 0; i < 8; ++i) ms8[i] = h->host_phases[i];
    return RWRT_OK;
}

int rwrt_last_host_windows(rwrt_handle h, double *rows, int max_windows, int *nwindows)
{
    if (!h || !nwindows || (max_windows > 0 && !rows)) return fail(RWRT_ERR_ARG, "rwrt_last_host_windows: NULL argument");
    const int n = (int)(h->host_windows.size() / RWRT_HOST_WINDOW_FIELDS);
    *nwindows = n;
    const int m = n < max_windows ? n : max_windows;
    if (m > 0) memcpy(rows, h->host_windows.data(), (size_t)m * RWRT_HOST_WINDOW_FIELDS * sizeof(double));
    return RWRT_OK;
}

int rwrt_measure_fp64_peak(rwrt_handle h, double *tflops)
{
    if (!h || !tflops) return fail(RWRT_ERR_ARG, "rwrt_measure_fp64_peak: NULL argument");
    CU(cudaSetDevice(h->device));
    if (measure_fp64_peak(h->sm_count, tflops, nullptr)) return fail(RWRT_ERR_CUDA, "DFMA microbenchmark failed");
    return RWRT_OK;
}

int rwrt_grid_accumulate(rwrt_handle h, int64_t nslot, int nt, const double *out7, double dlon_deg, double dlat_deg,
                         int ncx, int ncy, int64_t *count, int64_t *fx, int64_t *fy, void *stream)
{
    if (!h || !out7 || !count || !fx || !fy || nslot < 0 || nt < 1 || ncx < 1 || ncy < 1 || !(dlon_deg > 0) ||
        !(dlat_deg > 0))
        return fail(RWRT_ERR_ARG, "rwrt_grid_accumulate: bad argument");
    CU(cudaSetDevice(h->device));
    launch_grid_accumulate(nslot, nt, out7, dlon_deg, dlat_deg, ncx, ncy, (long long *)count, (long long *)fx,
                           (long long *)fy, (cudaStream_t)stream);
    CU(cudaGetLastError());
    return RWRT_OK;
}

// ---------------------------------------------------------------------------------------------
// Grid-only sweep and device-resident ray records
// ---------------------------------------------------------------------------------------------
static int check_grid(const char *who, double dlon_deg, double dlat_deg, int ncx, int ncy, const int64_t *count, const int64_t *fx,
                      const int64_t *fy)
{
    if (!count || !fx || !fy || ncx < 1 || ncy < 1 || !(dlon_deg > 0) || !(dlat_deg > 0))
        return fail(RWRT_ERR_ARG, "%s: bad grid argument", who);
    if ((long long)ncx * ncy > 2147483647LL) return fail(RWRT_ERR_ARG, "%s: more than 2^31-1 grid cells", who);
    return RWRT_OK;
}

int rwrt_ray_sweep_grid(rwrt_handle h, int nsource, const double *src_lon, const double *src_lat, const int32_t *src_member,
                        int nzwn, const double *zwn, double freq, int nt, double tstep, double ttotal, double rtol, double atol,
                        double min_step_factor, double cut_off, uint32_t flags, double dlon_deg, double dlat_deg, int ncx,
                        int ncy, int64_t *count, int64_t *fx, int64_t *fy, int64_t *accepted, int8_t *status,
                        int64_t *total_accepted, void *stream)
{
    if (!h || nsource < 0 || nzwn < 0 || nt < 1) return fail(RWRT_ERR_ARG, "rwrt_ray_sweep_grid: bad sizes");
    int rc = check_grid("rwrt_ray_sweep_grid", dlon_deg, dlat_deg, ncx, ncy, count, fx, fy);
    if (rc) return rc;
    if (!h->ready) return fail(RWRT_ERR_STATE, "rwrt_ray_sweep_grid before rwrt_bs_ready");
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaSetDevice(h->device));
    const long long N = 3LL * nsource * nzwn;
    if (N == 0) {
        if (total_accepted) CU(cudaMemsetAsync(total_accepted, 0, sizeof(int64_t), s));
        return RWRT_OK;
    }
    // the t = 0 state (wr.py:344-395) is all a sweep keeps of the seven planes: 7 x N doubles on the handle
    if (h->row0.reserve((size_t)7 * N * sizeof(double))) return fail(RWRT_ERR_NOMEM, "t = 0 state (%lld slots)", N);
    rc = rwrt_ray_init(h, nsource, src_lon, src_lat, src_member, nzwn, zwn, freq, 1, (double *)h->row0.p, stream);
    if (rc) return rc;
    // slots without a real root count zero steps / status 0 (expand_records writes these on the trajectory path)
    if (accepted) CU(cudaMemsetAsync(accepted, 0, (size_t)N * sizeof(int64_t), s));
    if (status) CU(cudaMemsetAsync(status, 0, (size_t)N, s));
    GridSpec g;
    g.dlon = dlon_deg;
    g.dlat = dlat_deg;
    g.ncx = ncx;
    g.ncy = ncy;
    g.count = (long long *)count;
    g.fx = (long long *)fx;
    g.fy = (long long *)fy;
    g.from_records = (flags & RWRT_GRID_FROM_RECORDS) != 0;
    RunOpts o;
    o.grid = &g;
    o.keep = KEEP_NOTHING;
    if (const char *e = getenv("RWRT_WINDOWS")) o.nwin = atoi(e);
    return run_rk45(h, nsource, src_member, nzwn, nt, tstep, ttotal, rtol, atol, min_step_factor, cut_off, flags,
                    (double *)h->row0.p, accepted, nullptr, status, total_accepted, stream, o);
}

int rwrt_ray_run_rk45_records(rwrt_handle h, int nsource, const int32_t *src_member, int nzwn, int nt, double tstep,
                              double ttotal, double rtol, double atol, double min_step_factor, double cut_off, uint32_t flags,
                              const double *out7_row0, int64_t *accepted, int64_t *attempts, int8_t *status,
                              int64_t *total_accepted, void *stream)
{
    if (!h || nsource < 0 || nzwn < 0 || nt < 1) return fail(RWRT_ERR_ARG, "rwrt_ray_run_rk45_records: bad sizes");
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaSetDevice(h->device));
    const long long N = 3LL * nsource * nzwn;
    if (N > 0 && !out7_row0) return fail(RWRT_ERR_ARG, "rwrt_ray_run_rk45_records: out7_row0 is NULL");
    if (N > 0) {
        // the handle keeps its own copy of the t = 0 row: rwrt_records_expand reads the constant rows of the slots
        // without a root from it long after the caller's buffer may be gone
        if (h->row0.reserve((size_t)7 * N * sizeof(double))) return fail(RWRT_ERR_NOMEM, "t = 0 state (%lld slots)", N);
        CU(cudaMemcpyAsync(h->row0.p, out7_row0, (size_t)7 * N * sizeof(double), cudaMemcpyDeviceToDevice, s));
        if (accepted) CU(cudaMemsetAsync(accepted, 0, (size_t)N * sizeof(int64_t), s));
        if (attempts) CU(cudaMemsetAsync(attempts, 0, (size_t)N * sizeof(int64_t), s));
        if (status) CU(cudaMemsetAsync(status, 0, (size_t)N, s));
    }
    RunOpts o;
    o.keep = KEEP_RECORDS;
    return run_rk45(h, nsource, src_member, nzwn, nt, tstep, ttotal, rtol, atol, min_step_factor, cut_off, flags,
                    (double *)h->row0.p, accepted, attempts, status, total_accepted, stream, o);
}

int rwrt_records_info(rwrt_handle h, int64_t *nray, int *rec_rows, const double **rec, const int32_t **rows_done,
                      const int32_t **ray_slot, const int32_t **slot_to_ray)
{
    if (!h) return fail(RWRT_ERR_ARG, "rwrt_records_info: null handle");
    if (!h->have_records) return fail(RWRT_ERR_STATE, "rwrt_records_info: no ray records on the handle (run rwrt_ray_run_rk45_records)");
    const RunParams &p = h->last_p;
    if (nray) *nray = p.nlive;
    if (rec_rows) *rec_rows = p.rec_rows;
    if (rec) *rec = p.rec;
    if (rows_done) *rows_done = p.rows_done;
    if (ray_slot) *ray_slot = p.live;
    if (slot_to_ray) *slot_to_ray = p.inv;
    return RWRT_OK;
}

int rwrt_records_expand(rwrt_handle h, int row_begin, int row_end, double *dst7, void *stream)
{
    if (!h || !dst7) return fail(RWRT_ERR_ARG, "rwrt_records_expand: NULL argument");
    if (!h->have_records) return fail(RWRT_ERR_STATE, "rwrt_records_expand: no ray records on the handle");
    RunParams q = h->last_p;
    if (row_begin < 1 || row_end > q.nt || row_end < row_begin)
        return fail(RWRT_ERR_ARG, "rwrt_records_expand: rows [%d, %d) outside [1, %d)", row_begin, row_end, q.nt);
    CU(cudaSetDevice(h->device));
    q.it_begin = row_begin;
    q.it_end = row_end;
    q.dst = dst7;
    q.dst_pl = (size_t)q.N * (size_t)(row_end - row_begin);
    q.dst_it0 = row_begin;
    if (h->last_strict) launch_expand_records_strict(q, (cudaStream_t)stream);
    else launch_expand_records_fast(q, (cudaStream_t)stream);
    CU(cudaGetLastError());
    return RWRT_OK;
}

int rwrt_grid_accumulate_records(rwrt_handle h, double dlon_deg, double dlat_deg, int ncx, int ncy, int64_t *count, int64_t *fx,
                                 int64_t *fy, void *stream)
{
    if (!h) return fail(RWRT_ERR_ARG, "rwrt_grid_accumulate_records: null handle");
    int rc = check_grid("rwrt_grid_accumulate_records", dlon_deg, dlat_deg, ncx, ncy, count, fx, fy);
    if (rc) return rc;
    if (!h->have_records) return fail(RWRT_ERR_STATE, "rwrt_grid_accumulate_records: no ray records on the handle");
    const RunParams &p = h->last_p;
    CU(cudaSetDevice(h->device));
    launch_grid_records(p.rec, p.rec_rows, p.rows_done, p.nlive, 1, p.nt, dlon_deg, dlat_deg, ncx, ncy, (long long *)count,
                        (long long *)fx, (long long *)fy, h->sm_count, (cudaStream_t)stream);
    CU(cudaGetLastError());
    return RWRT_OK;
}

namespace {
struct HostCopyCtx {
    cudaStream_t copy_stream;
    cudaEvent_t ev[2];
    int n_ev;
    const double *dev;
    double *host;
    long long N;
    int nt;
    bool first;
};

// window hook of rwrt_real2d_host: rows [it_begin, it_end) of the seven planes are final once this window's
// kernel has finished -> ship them on the copy stream (7 contiguous copies) while the next window integrates.
int host_copy_hook(void *vctx, int it_begin, int it_end, cudaStream_t s)
{
    HostCopyCtx *c = (HostCopyCtx *)vctx;
    cudaEvent_t e = c->ev[c->n_ev & 1];
    ++c->n_ev;
    CU(cudaEventRecord(e, s));
    CU(cudaStreamWaitEvent(c->copy_stream, e, 0));
    int r0 = c->first ? 0 : it_begin;  // the t = 0 row travels with the first window
    c->first = false;
    for (int pl = 0; pl < 7; ++pl) {
        size_t off = ((size_t)pl * c->nt + r0) * (size_t)c->N;
        CU(cudaMemcpyAsync(c->host + off, c->dev + off, (size_t)(it_end - r0) * c->N * sizeof(double),
                           cudaMemcpyDeviceToHost, c->copy_stream));
    }
    return RWRT_OK;
}
}  // namespace

int rwrt_ray_run_rk4(rwrt_handle h, int nsource, const int32_t *src_member, int nzwn, int nt, double tstep, double cut_off,
                     uint32_t flags, double *out7, int8_t *status, void *stream)
{
    if (!h || nsource < 0 || nzwn < 0 || nt < 1) return fail(RWRT_ERR_ARG, "rwrt_ray_run_rk4: bad sizes");
    if (!(tstep > 0)) return fail(RWRT_ERR_ARG, "rwrt_ray_run_rk4: tstep must be positive");
    if (!h->ready) return fail(RWRT_ERR_STATE, "rwrt_ray_run_rk4 before rwrt_bs_ready");
    long long N = 3LL * nsource * nzwn;
    if (N == 0 || nt < 2) return RWRT_OK;
    if (!out7) return fail(RWRT_ERR_ARG, "rwrt_ray_run_rk4: out7 is NULL");
    CU(cudaSetDevice(h->device));
    RunParams p;
    memset(&p, 0, sizeof p);
    p.g = h->grid();
    p.src_member = src_member;
    p.nsource = nsource;
    p.nzwn = nzwn;
    p.nt = nt;
    p.N = N;
    p.tstep = tstep;
    p.cut = cut_off * tstep / 3600.;  // wr.py:170
    p.out = out7;
    p.status = (signed char *)status;
    // fast build: cooperative-gather kernel with FSAL (40.7 ms per C2 pass against 56.5); strict build: the per-lane
    // kernel (143 ms against 156: libdevice trigonometry dominates there).  RWRT_PER_LANE_GATHER=0 / 1 forces one.
    p.coop = (flags & RWRT_MODE_STRICT) ? 0 : 1;
    if (const char *e = getenv("RWRT_PER_LANE_GATHER")) p.coop = (atoi(e) == 0);
    if ((long long)h->nmember * h->nlon_f * h->nlat >= (1LL << 30) - 1) p.coop = 0;
    if (flags & RWRT_MODE_STRICT) launch_rk4_strict(p, (cudaStream_t)stream);
    else launch_rk4_fast(p, (cudaStream_t)stream);
    CU(cudaGetLastError());
    return RWRT_OK;
}

// how the windows of a host-buffer call travel (see hx_on_window): RWRT_HOST_ROUTE = auto (default) | hx | direct | alt
static void hx_configure(HostExpand *hx)
{
    cudaPointerAttributes at;
    hx->can_direct = cudaPointerGetAttributes(&at, hx->out7_host) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    hx->route_mode = 0;
    if (const char *e = getenv("RWRT_HOST_ROUTE"))
        hx->route_mode = !strcmp(e, "hx") ? 1 : !strcmp(e, "direct") ? 2 : !strcmp(e, "alt") ? 3 : 0;
    if (hx->route_mode == 1) hx->can_direct = false;
}

static int host_windows_default(size_t nout_doubles) { return (int)((nout_doubles * sizeof(double)) >> 29); }

int rwrt_ray_run_rk45_d2h(rwrt_handle h, int nsource, const int32_t *src_member, int nzwn, int nt, double tstep,
                          double ttotal, double rtol, double atol, double min_step_factor, double cut_off, uint32_t flags,
                          double *out7, int64_t *accepted, int64_t *attempts, int8_t *status, int64_t *total_accepted,
                          double *out7_host, void *stream)
{
    if (!h || !out7_host) return fail(RWRT_ERR_ARG, "rwrt_ray_run_rk45_d2h: NULL argument");
    for (double &x : h->host_stats) x = 0;
    h->host_windows.clear();
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaSetDevice(h->device));
    const long long N = 3LL * nsource * nzwn;
    const size_t nout = (size_t)7 * nt * N;
    int nwin = host_windows_default(nout);
    if (const char *e = getenv("RWRT_HOST_WINDOWS")) nwin = atoi(e);
    if (nwin < 1) nwin = 1;
    if (nwin > 64) nwin = 64;
    if (!h->copy_stream) CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    cudaStream_t copy_stream = h->copy_stream;
    // Default: the device planes are filled as in rwrt_ray_run_rk45 and, beside them, the packed rows of the live rays
    // travel to the host where threads write out7_host (see rwrt_real2d_host); RWRT_HOST_EXPAND=0 or a run without
    // ray records: the finished rows of the device planes are copied as they are.
    bool use_hx = nt >= 2 && N > 0 && out7;
    if (const char *e = getenv("RWRT_HOST_EXPAND")) use_hx = use_hx && atoi(e) != 0;
    int rc = RWRT_HX_UNSUPPORTED;
    if (use_hx) {
        HostExpand hx;
        hx.out7_host = out7_host;
        hx.N = N;
        hx.nt = nt;
        hx.copy_stream = copy_stream;
        hx.row0_only = false;
        hx_configure(&hx);
        RunOpts o;
        o.nwin = nwin;
        o.hx = &hx;
        rc = run_rk45(h, nsource, src_member, nzwn, nt, tstep, ttotal, rtol, atol, min_step_factor, cut_off, flags, out7,
                      accepted, attempts, status, total_accepted, stream, o);
        if (rc == RWRT_OK) {
            if (cudaStreamSynchronize(s) != cudaSuccess || cudaStreamSynchronize(copy_stream) != cudaSuccess)
                rc = fail(RWRT_ERR_CUDA, "rwrt_ray_run_rk45_d2h: stream synchronisation failed");
        }
        int rc2 = hx_finish(&hx, rc == RWRT_OK);
        if (rc == RWRT_OK && rc2) rc = fail(rc2, "host-side expansion failed");
    }
    if (rc == RWRT_HX_UNSUPPORTED) {
        HostCopyCtx ctx;
        ctx.copy_stream = copy_stream;
        ctx.ev[0] = ctx.ev[1] = nullptr;
        ctx.n_ev = 0;
        ctx.dev = out7;
        ctx.host = out7_host;
        ctx.N = N;
        ctx.nt = nt;
        ctx.first = true;
        CU(cudaEventCreateWithFlags(&ctx.ev[0], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx.ev[1], cudaEventDisableTiming));
        RunOpts o;
        o.nwin = nwin;
        o.hook = host_copy_hook;
        o.hook_ctx = &ctx;
        rc = run_rk45(h, nsource, src_member, nzwn, nt, tstep, ttotal, rtol, atol, min_step_factor, cut_off, flags, out7,
                      accepted, attempts, status, total_accepted, stream, o);
        if (rc == RWRT_OK && ctx.first && nout > 0 && out7) {
            cudaError_t e = cudaMemcpyAsync(out7_host, out7, nout * 8, cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess) rc = fail(RWRT_ERR_CUDA, "cudaMemcpyAsync failed: %s", cudaGetErrorString(e));
        }
        cudaStreamSynchronize(s);
        cudaStreamSynchronize(copy_stream);
        cudaEventDestroy(ctx.ev[0]);
        cudaEventDestroy(ctx.ev[1]);
    }
    return rc;
}

static int real2d_host_impl(rwrt_handle h, const float *u, const float *v, const double *lon, const double *lat,
                            const double *coslat, const double *sinlat, int nsource, const double *src_lon,
                            const double *src_lat, const int32_t *src_member, int nzwn, const double *zwn, double freq, int nt,
                            double tstep, double ttotal, double rtol, double atol, double min_step_factor, double cut_off,
                            uint32_t flags, double *out7_host, int64_t *accepted_host, int8_t *status_host,
                            int64_t *total_accepted_host, void *stream, bool use_hx)
{
    cudaStream_t s = (cudaStream_t)stream;
    const auto t_entry = std::chrono::steady_clock::now();
    auto phase = [&](int i) { h->host_phases[i] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count(); };
    for (double &x : h->host_phases) x = 0;
    CU(cudaSetDevice(h->device));
    const size_t ng = (size_t)h->nmember * h->nlon * h->nlat;
    const long long N = 3LL * nsource * nzwn;
    const size_t nout = (size_t)7 * nt * N;
    // The integration runs in time windows; what a finished window produced travels to the host on a second stream
    // while the next window integrates.
    // ~0.5 GB of result planes per window (device-expanded rows on C2, v9: 1 / 4 / 10 / 24 windows -> 683 / 631 / 548 /
    // 492 ms per pass on the same box; packed rows, v16: 10 / 21 / 30 / 40 windows -> 326 / 305 / 312 / 300 ms).  Small
    // windows also keep the pinned staging buffers small (4 x 0.14 GB on C2): pinning is what a one-shot call pays for.
    int nwin = host_windows_default(nout);
    if (const char *e = getenv("RWRT_HOST_WINDOWS")) nwin = atoi(e);
    if (nwin < 1) nwin = 1;
    if (nwin > 64) nwin = 64;
    // one device arena for every input and output of the call; with host-side expansion the device holds only the
    // t = 0 row of the seven planes (7 x N) -- the other rows exist as ray records in the library's window buffers
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    size_t o_u = carve(ng * 4), o_v = carve(ng * 4), o_lon = carve((size_t)h->nlon * 8), o_lat = carve((size_t)h->nlat * 8);
    size_t o_cl = carve((size_t)h->nlat * 8), o_sl = carve((size_t)h->nlat * 8);
    size_t o_slon = carve((size_t)nsource * 8), o_slat = carve((size_t)nsource * 8), o_mem = carve((size_t)nsource * 4);
    size_t o_zwn = carve((size_t)nzwn * 8), o_acc = carve((size_t)N * 8), o_st = carve((size_t)N), o_tot = carve(8);
    size_t o_out = carve(use_hx ? (size_t)7 * N * 8 : nout * 8);
    if (h->arena.reserve(off)) return fail(RWRT_ERR_NOMEM, "device arena of the host-buffer call (%zu bytes)", off);
    char *arena = (char *)h->arena.p;
    phase(0);                                       // device arena allocated
    h->host_stats[0] = (use_hx ? 0.0 : (double)nout * 8) + (accepted_host ? (double)N * 8 : 0.0) + (status_host ? (double)N : 0.0) + 8;
    h->host_stats[1] = (double)ng * 8 + (double)h->nlon * 8 + (double)h->nlat * ((coslat && sinlat) ? 24 : 8) +
                       (double)nsource * (src_member ? 20 : 16) + (double)nzwn * 8;
    h->host_stats[2] = nwin;
    h->host_stats[3] = 0;
    h->host_stats[4] = 0;
    h->host_windows.clear();
    HostCopyCtx ctx;
    ctx.copy_stream = nullptr;
    ctx.ev[0] = ctx.ev[1] = nullptr;
    ctx.n_ev = 0;
    ctx.dev = (const double *)(arena + o_out);
    ctx.host = out7_host;
    ctx.N = N;
    ctx.nt = nt;
    ctx.first = true;
    HostExpand hx;
    hx.out7_host = out7_host;
    hx.N = N;
    hx.nt = nt;
    int rc = RWRT_OK;
#define CUX(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            rc = fail(RWRT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));                   \
            goto done;                                                                                  \
        }                                                                                               \
    } while (0)
    if (!h->copy_stream) CUX(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    ctx.copy_stream = h->copy_stream;
    hx.copy_stream = ctx.copy_stream;
    hx_configure(&hx);
    CUX(cudaEventCreateWithFlags(&ctx.ev[0], cudaEventDisableTiming));
    CUX(cudaEventCreateWithFlags(&ctx.ev[1], cudaEventDisableTiming));
    CUX(cudaMemcpyAsync(arena + o_u, u, ng * 4, cudaMemcpyHostToDevice, s));
    CUX(cudaMemcpyAsync(arena + o_v, v, ng * 4, cudaMemcpyHostToDevice, s));
    CUX(cudaMemcpyAsync(arena + o_lon, lon, (size_t)h->nlon * 8, cudaMemcpyHostToDevice, s));
    CUX(cudaMemcpyAsync(arena + o_lat, lat, (size_t)h->nlat * 8, cudaMemcpyHostToDevice, s));
    if (coslat && sinlat) {
        CUX(cudaMemcpyAsync(arena + o_cl, coslat, (size_t)h->nlat * 8, cudaMemcpyHostToDevice, s));
        CUX(cudaMemcpyAsync(arena + o_sl, sinlat, (size_t)h->nlat * 8, cudaMemcpyHostToDevice, s));
    }
    if (nsource > 0) {
        CUX(cudaMemcpyAsync(arena + o_slon, src_lon, (size_t)nsource * 8, cudaMemcpyHostToDevice, s));
        CUX(cudaMemcpyAsync(arena + o_slat, src_lat, (size_t)nsource * 8, cudaMemcpyHostToDevice, s));
        if (src_member) CUX(cudaMemcpyAsync(arena + o_mem, src_member, (size_t)nsource * 4, cudaMemcpyHostToDevice, s));
    }
    if (nzwn > 0) CUX(cudaMemcpyAsync(arena + o_zwn, zwn, (size_t)nzwn * 8, cudaMemcpyHostToDevice, s));
    CUX(cudaMemsetAsync(arena + o_tot, 0, 8, s));
    if (use_hx) {   // counters and status of NaN-root slots are zero (expand_records writes them on the device path)
        CUX(cudaMemsetAsync(arena + o_acc, 0, (size_t)N * 8, s));
        CUX(cudaMemsetAsync(arena + o_st, 0, (size_t)N, s));
    }
    {
        const int32_t *dmem = src_member ? (const int32_t *)(arena + o_mem) : nullptr;
        rc = rwrt_bs_ready(h, (const float *)(arena + o_u), (const float *)(arena + o_v), (const double *)(arena + o_lon),
                           (const double *)(arena + o_lat), (coslat && sinlat) ? (const double *)(arena + o_cl) : nullptr,
                           (coslat && sinlat) ? (const double *)(arena + o_sl) : nullptr, nullptr, nullptr, s);
        if (rc) goto done;
        rc = rwrt_ray_init(h, nsource, (const double *)(arena + o_slon), (const double *)(arena + o_slat), dmem, nzwn,
                           (const double *)(arena + o_zwn), freq, use_hx ? 1 : nt, (double *)(arena + o_out), s);
        if (rc) goto done;
        phase(1);                                   // inputs, BS.ready and the t = 0 state queued
        RunOpts o;
        o.nwin = nwin;
        o.hook = use_hx ? nullptr : host_copy_hook;
        o.hook_ctx = &ctx;
        o.hx = use_hx ? &hx : nullptr;
        rc = run_rk45(h, nsource, dmem, nzwn, nt, tstep, ttotal, rtol, atol, min_step_factor, cut_off, flags,
                      (double *)(arena + o_out), (int64_t *)(arena + o_acc), nullptr, (int8_t *)(arena + o_st),
                      (int64_t *)(arena + o_tot), s, o);
        if (rc) goto done;      // RWRT_HX_UNSUPPORTED included: the caller runs the device-side expansion instead
        if (use_hx) h->host_phases[2] = std::chrono::duration<double, std::milli>(hx.t0 - t_entry).count();   // pipeline origin (hx_begin): live rays known, t = 0 row on the host
        phase(3);                                   // every time window queued (the routing decisions wait for copies: most of the pipeline is over)
    }
    if (!use_hx && ctx.first && nout > 0)  // nothing was integrated (no live ray / nt == 1): plain copy of what there is
        CUX(cudaMemcpyAsync(out7_host, arena + o_out, nout * 8, cudaMemcpyDeviceToHost, s));
    if (accepted_host && N > 0) CUX(cudaMemcpyAsync(accepted_host, arena + o_acc, (size_t)N * 8, cudaMemcpyDeviceToHost, s));
    if (status_host && N > 0) CUX(cudaMemcpyAsync(status_host, arena + o_st, (size_t)N, cudaMemcpyDeviceToHost, s));
    if (total_accepted_host) CUX(cudaMemcpyAsync(total_accepted_host, arena + o_tot, 8, cudaMemcpyDeviceToHost, s));
    CUX(cudaStreamSynchronize(s));
    CUX(cudaStreamSynchronize(ctx.copy_stream));
    phase(4);                                       // device work and copies finished
done:
    if (use_hx) {
        int rc2 = hx_finish(&hx, rc == RWRT_OK);    // joins the workers: every window is in the caller's arrays
        if (rc == RWRT_OK && rc2) rc = fail(rc2, "host-side expansion failed");
    }
    phase(5);                                       // host threads joined
    if (ctx.copy_stream) cudaStreamSynchronize(ctx.copy_stream);
    if (ctx.ev[0]) cudaEventDestroy(ctx.ev[0]);
    if (ctx.ev[1]) cudaEventDestroy(ctx.ev[1]);
    phase(6);                                       // per-call objects released
    return rc;
#undef CUX
}

int rwrt_real2d_host(rwrt_handle h, const float *u, const float *v, const double *lon, const double *lat,
                     const double *coslat, const double *sinlat, int nsource, const double *src_lon,
                     const double *src_lat, const int32_t *src_member, int nzwn, const double *zwn, double freq, int nt,
                     double tstep, double ttotal, double rtol, double atol, double min_step_factor, double cut_off,
                     uint32_t flags, double *out7_host, int64_t *accepted_host, int8_t *status_host,
                     int64_t *total_accepted_host, void *stream)
{
    if (!h || !u || !v || !lon || !lat || !src_lon || !src_lat || !zwn || !out7_host)
        return fail(RWRT_ERR_ARG, "rwrt_real2d_host: NULL argument");
    if (nsource < 0 || nzwn < 0 || nt < 1) return fail(RWRT_ERR_ARG, "rwrt_real2d_host: bad sizes");
    // Default: packed rows of the live rays travel and host threads write the caller's arrays from them
    // (RWRT_HOST_EXPAND=0: the device expands and the reference-layout rows travel).  Runs without live rays / on the
    // per-lane kernel have no ray records: they take the second form.
    bool use_hx = nt >= 2 && (long long)nsource * nzwn > 0;
    if (const char *e = getenv("RWRT_HOST_EXPAND")) use_hx = use_hx && atoi(e) != 0;
    if (use_hx) {
        int rc = real2d_host_impl(h, u, v, lon, lat, coslat, sinlat, nsource, src_lon, src_lat, src_member, nzwn, zwn, freq, nt,
                                  tstep, ttotal, rtol, atol, min_step_factor, cut_off, flags, out7_host, accepted_host,
                                  status_host, total_accepted_host, stream, true);
        if (rc != RWRT_HX_UNSUPPORTED) return rc;
    }
    return real2d_host_impl(h, u, v, lon, lat, coslat, sinlat, nsource, src_lon, src_lat, src_member, nzwn, zwn, freq, nt, tstep,
                            ttotal, rtol, atol, min_step_factor, cut_off, flags, out7_host, accepted_host, status_host,
                            total_accepted_host, stream, false);
}

}  // extern "C"
