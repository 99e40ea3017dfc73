/*
 * rwrt.h -- C ABI of librwrt.so, the B200 (sm_100a) ray-integration engine.
 *
 * Drop-in boundary for the RK45 ray path of yinan-codes/Rossby-wave-ray-tracing.  The reference
 * is pure Python; its "FFI" for this path is the set of Python calls listed beside each entry
 * point (file:line into the reference).  A maintainer binds these symbols with ctypes
 * (INTEGRATION.md shows the stub); `rossby-wave-ray-tracing_b200/lib.py` is that binding.
 *
 * Conventions
 *   - plain C types only; every array argument is a DEVICE pointer unless the function name ends
 *     in `_host` (those take host pointers and do their own H2D / D2H on the given stream);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are
 *     stream-ordered and return without synchronising unless documented otherwise;
 *   - return value: 0 = ok, negative = error (rwrt_status); rwrt_last_error() gives a
 *     thread-local message.  No C++ exception crosses the boundary;
 *   - the caller owns every buffer it passes in; the handle owns only its private field copies
 *     and work lists, released by rwrt_destroy;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     RWRT_ERR_CUDA.
 *
 * Array layouts are the reference's own: grids are [i_lon][j_lat] (bs.py:245-247); the seven ray
 * result arrays are (nt, 3, nsource, nzwn) C-order, zwn fastest (wr.py:160-167); a "slot" is the
 * flattened (root, source, zwn) index.
 */
#ifndef RWRT_H
#define RWRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RWRT_ABI_VERSION 2

typedef struct rwrt_ctx *rwrt_handle;

typedef enum {
    RWRT_OK = 0,
    RWRT_ERR_ARG = -1,   /* bad argument (the facade raises the reference's ValueError)      */
    RWRT_ERR_CUDA = -2,  /* CUDA runtime error / no device                                    */
    RWRT_ERR_STATE = -3, /* call order: e.g. ray_init before bs_ready                         */
    RWRT_ERR_NOMEM = -4
} rwrt_status;

/* flags for rwrt_ray_run_rk45 */
#define RWRT_MODE_FAST 0u   /* FMA contraction allowed (default)                              */
#define RWRT_MODE_STRICT 1u /* every operation a separate IEEE double op, reference order    */
/* flag for rwrt_ray_sweep_grid: grid each time window's ray records with a second kernel instead of
 * accumulating inside the integrator (same bits; kept as the cross-check of the fused form)       */
#define RWRT_GRID_FROM_RECORDS 2u

/* per-slot status codes written by rwrt_ray_run_rk45 (new, additive: the reference only has NaN) */
#define RWRT_RAY_NO_ROOT 0   /* NaN root at t = 0: never propagates   */
#define RWRT_RAY_ALIVE 1     /* reached ttotal                                                 */
#define RWRT_RAY_POLE 2      /* |lat| >= pi/2 at an output time           (wr.py:838-843)      */
#define RWRT_RAY_SPEED 3     /* moved >= cut_off between two outputs      (wr.py:844-850)      */
#define RWRT_RAY_NAN 4       /* state became NaN while stepping: |l| >= 100 critical line or
                                pole inside a step (wr.py:508-514, rkf45.py:400-403, 446)     */

int rwrt_abi_version(void);
const char *rwrt_last_error(void);
int rwrt_device_count(int *count);

/* Library-owned pinned host memory (for result arrays that are copied back asynchronously). */
int rwrt_host_alloc(void **ptr, size_t bytes);
int rwrt_host_free(void *ptr);

/*
 * One handle per (device, background grid).  Replaces WR.__init__ -> BS.__init__ (wr.py:133-171,
 * bs.py:70-112).  nmember > 1 batches an ensemble of background states on the same grid.
 */
int rwrt_create(int device, int nlon, int nlat, int xcyclic, int nmember, rwrt_handle *out);
int rwrt_destroy(rwrt_handle h);

/*
 * BS.ready(xcyclic) (bs.py:318-407): finite-difference derivative fields, 9-point smoothing,
 * stacking.  u, v: (nmember, nlon, nlat) float32.  lon (nlon), lat (nlat) float64 radians: the
 * coordinate arrays the sampler indexes with (interpolation.py:78-82).  coslat/sinlat (nlat):
 * optional tables (NULL = computed on the device); the facade passes NumPy's values so that the
 * fields are bit-identical to the reference's.
 * Outputs (either may be NULL): fields18 (nmember, nlon + xcyclic, nlat, 18) in the reference's
 * order (bs.py:349-372); diag9 (nmember, 9, nlon, nlat): q, uxx, uyy, vxx, vyy, uxy, vxy, betam, KS.
 * The handle keeps its own packed copy of the 11 fields the ray equations consume.
 */
int rwrt_bs_ready(rwrt_handle h, const float *u, const float *v, const double *lon, const double *lat,
                  const double *coslat, const double *sinlat, double *fields18, double *diag9, void *stream);

/* BS.read_from_precaled_nc analogue (bs.py:409-454): adopt externally computed fields18. */
int rwrt_bs_set_fields(rwrt_handle h, const double *fields18, const double *lon, const double *lat, void *stream);

/*
 * BS.cal_bs_mercator_point(lon, lat, mode='numpy') (bs.py:781-887): bilinear sample of the fields
 * at n points + Mercator transform.  out: (12, n) = fmu, fmv, fmux, fmuy, fmvx, fmvy, fmqx, fmqy,
 * fmqxx, fmqxy, fmqyx, fmqyy (the rows the ray path uses).  member: NULL or (n) int32.
 */
int rwrt_sample_mercator(rwrt_handle h, int64_t n, const double *lon, const double *lat, const int32_t *member,
                         double *out, void *stream);

/* All 18 rows of BS.cal_bs_mercator_point (bs.py:856-887), sampled from a fields18 table as written by
 * rwrt_bs_ready (rows 12..17 are the third-derivative terms the ray path never consumes).  out: (18, n). */
int rwrt_sample_mercator18(rwrt_handle h, const double *fields18, int64_t n, const double *lon, const double *lat,
                           const int32_t *member, double *out, void *stream);

/* One right-hand side per point: WR.diffun_numpy (wr.py:492-556).  y, dy: (5, n). */
int rwrt_rhs(rwrt_handle h, int64_t n, const double *y, const int32_t *member, double *dy, uint32_t flags,
             void *stream);

/*
 * WR.ray_initial_numpy (wr.py:344-395) + cal_ky_numpy / change_roots_order (bs.py:942-1040):
 * writes row 0 of the seven planes of out7 = (7, nt, 3, nsource, nzwn).  src_lon/src_lat radians.
 * src_member: NULL or (nsource) int32 ensemble member of each source.
 */
int rwrt_ray_init(rwrt_handle h, int nsource, const double *src_lon, const double *src_lat,
                  const int32_t *src_member, int nzwn, const double *zwn, double freq, int nt, double *out7,
                  void *stream);

/*
 * WR.core_ray_run_rk45 (wr.py:767-887) driving RK45 (rkf45.py:222-514): integrates every slot from
 * row 0 of out7 to ttotal, writing rows 1..nt-1.  tstep/ttotal in seconds; cut_off as in
 * WR.__init__ (0.1).  accepted / attempts: (3*nsource*nzwn) int64, status: int8 -- all optional.
 * total_accepted (optional, device int64[1]) receives the sum of accepted steps (the metric).
 */
int rwrt_ray_run_rk45(rwrt_handle h, int nsource, const int32_t *src_member, int nzwn, int nt, double tstep,
                      double ttotal, double rtol, double atol, double min_step_factor, double cut_off,
                      uint32_t flags, double *out7, int64_t *accepted, int64_t *attempts, int8_t *status,
                      int64_t *total_accepted, void *stream);

/*
 * WR.core_ray_run_numpy (wr.py:702-765): the fixed-step RK4 path (inte_method = '', main_wr.py's default)
 * with rk4_step_numpy's freeze-on-failure semantics (wr.py:583-622).  Same out7 convention as above.
 */
int rwrt_ray_run_rk4(rwrt_handle h, int nsource, const int32_t *src_member, int nzwn, int nt, double tstep,
                     double cut_off, uint32_t flags, double *out7, int8_t *status, void *stream);

/*
 * Same as rwrt_ray_run_rk45, and additionally delivers the seven result planes to HOST memory
 * out7_host (7, nt, 3, nsource, nzwn), pageable or pinned: the integration runs in time windows and what a
 * finished window produced travels while the next one integrates -- by default as packed rows of the live rays that
 * host threads expand into out7_host (see rwrt_real2d_host; RWRT_HOST_EXPAND=0: the finished rows of the device
 * planes are copied as they are, fast only into pinned memory).  out7 is the device buffer (row 0 = initial state,
 * as above) and holds the complete planes afterwards.  Synchronises before returning.
 */
int rwrt_ray_run_rk45_d2h(rwrt_handle h, int nsource, const int32_t *src_member, int nzwn, int nt, double tstep,
                          double ttotal, double rtol, double atol, double min_step_factor, double cut_off,
                          uint32_t flags, double *out7, int64_t *accepted, int64_t *attempts, int8_t *status,
                          int64_t *total_accepted, double *out7_host, void *stream);

/*
 * Ray-density / wave-ray-flux gridding of trajectories (manual section 4; no reference code --
 * parity unpinned, defined by oracle/gridding.py).  The domain is three longitude circles
 * [-360, 720) deg x [-90, 90] deg with cells of (dlon_deg, dlat_deg).  For every output time it >= 1
 * and every live sample: count[cell] += 1 (int64), flux_x/flux_y[cell] += round(ug|vg * 2^20) (int64
 * fixed point, so that the sums are order independent and bit exact across GPUs).
 * count, fx, fy: (ncx, ncy) int64 with ncx = ceil(1080/dlon_deg), ncy = ceil(180/dlat_deg).
 */
int rwrt_grid_accumulate(rwrt_handle h, int64_t nslot, int nt, const double *out7, double dlon_deg, double dlat_deg,
                         int ncx, int ncy, int64_t *count, int64_t *fx, int64_t *fy, void *stream);

/*
 * Grid-only sweep: ray_initial + RK45 + gridding in one pass, with NO trajectory arrays.  What a global sweep
 * (BASELINE configs[4]: 10^7 ray slots x 721 output times) must avoid is the reference's up-front result allocation
 * of 7 * nt * N doubles (wr.py:157-167: 404 GB there); the quantities the gridding tool reads from the ray file
 * (wr.py:947-959: rlon, rlat, rmwn, rug, rvg) are consumed where they are produced instead.  By default the
 * integrator adds every output sample to count / fx / fy itself (three integer reductions per output row, same
 * definition and same bits as rwrt_grid_accumulate on the trajectories of rwrt_ray_run_rk45); with
 * RWRT_GRID_FROM_RECORDS the integration runs in time windows whose 64-byte ray records a second kernel grids.
 * src_lon/src_lat (nsource) radians, zwn (nzwn): device pointers as in rwrt_ray_init.  count, fx, fy (ncx, ncy) int64
 * are ACCUMULATED INTO (zero them first; one ncclAllReduce(sum) over them merges the ranks of a sharded sweep --
 * integer sums, so the result does not depend on the sharding).  accepted (N) / status (N) / total_accepted: optional.
 */
int rwrt_ray_sweep_grid(rwrt_handle h, int nsource, const double *src_lon, const double *src_lat,
                        const int32_t *src_member, int nzwn, const double *zwn, double freq, int nt, double tstep,
                        double ttotal, double rtol, double atol, double min_step_factor, double cut_off, uint32_t flags,
                        double dlon_deg, double dlat_deg, int ncx, int ncy, int64_t *count, int64_t *fx, int64_t *fy,
                        int64_t *accepted, int8_t *status, int64_t *total_accepted, void *stream);

/*
 * rwrt_ray_run_rk45 with the RAY RECORDS as the native result: the integrator's own output -- one 64-byte record
 * (lon, lat, k, l, A, u_g, v_g, pad) per live ray and output row, rec[ray][row - 1][8] for rows 1..nt-1 -- stays in
 * library memory and the transpose to the reference's seven planes (wr.py:160-167, 868-885) happens only on demand
 * (rwrt_records_expand, any row range) or never (rwrt_grid_accumulate_records).  out7_row0: (7, N) = row 0 of the
 * seven planes as rwrt_ray_init(..., nt = 1, ...) writes it (copied to the handle).  The records stay valid until the
 * next run or rwrt_trim on this handle.  Fails with RWRT_ERR_NOMEM when nray * (nt - 1) * 64 B does not fit.
 */
int rwrt_ray_run_rk45_records(rwrt_handle h, int nsource, const int32_t *src_member, int nzwn, int nt, double tstep,
                              double ttotal, double rtol, double atol, double min_step_factor, double cut_off,
                              uint32_t flags, const double *out7_row0, int64_t *accepted, int64_t *attempts,
                              int8_t *status, int64_t *total_accepted, void *stream);
/* device pointers into the handle's records (any output pointer may be NULL): nray live rays; rec (nray, rec_rows, 8);
 * rows_done[ray] = rows 1 .. rows_done-1 are valid (a ray that died at row r has rows_done = r: wr.py:838-855);
 * ray_slot[ray] = its slot; slot_to_ray[slot] = its ray or -1 for a slot without a real root. */
int rwrt_records_info(rwrt_handle h, int64_t *nray, int *rec_rows, const double **rec, const int32_t **rows_done,
                      const int32_t **ray_slot, const int32_t **slot_to_ray);
/* rows [row_begin, row_end) (1 <= row_begin) of the seven reference-layout planes from the records:
 * dst7 (7, row_end - row_begin, N) device. */
int rwrt_records_expand(rwrt_handle h, int row_begin, int row_end, double *dst7, void *stream);
/* rwrt_grid_accumulate over the handle's records (all rows); count / fx / fy are accumulated into. */
int rwrt_grid_accumulate_records(rwrt_handle h, double dlon_deg, double dlat_deg, int ncx, int ncy, int64_t *count,
                                 int64_t *fx, int64_t *fy, void *stream);

/*
 * Whole ray path with HOST buffers (the call main_wr.real2d_hnf makes, main_wr.py:66-89, minus
 * NetCDF): H2D of u, v, coordinates, sources; BS.ready; ray_initial; RK45; the seven result arrays
 * (wr.py:767-781) in out7_host.  Synchronises before returning.
 * The integration runs in time windows.  Default: the rows of a finished window travel as seven packed planes of
 * the LIVE rays (56 bytes per live ray and row) to pinned staging buffers kept on the handle while the next window
 * integrates, and host threads (RWRT_HOST_THREADS, default: all cores) write the caller's planes from them -- about
 * half of the reference-layout bytes are NaN / constant rows that never need to cross the link, and out7_host may
 * be pageable.  When out7_host is pinned and all staging buffers are still waiting for the host threads (the cores, not
 * the link, are the bottleneck), a window's device-expanded rows are copied straight into out7_host instead
 * (RWRT_HOST_ROUTE=hx never does that, =direct always).  RWRT_HOST_EXPAND=0: the device
 * expands the records and the reference-layout rows travel (fast only into pinned memory, rwrt_host_alloc).
 * Both forms give the same bits.
 */
int rwrt_real2d_host(rwrt_handle h, const float *u, const float *v, const double *lon, const double *lat,
                     const double *coslat, const double *sinlat, int nsource, const double *src_lon,
                     const double *src_lat, const int32_t *src_member, int nzwn, const double *zwn, double freq,
                     int nt, double tstep, double ttotal, double rtol, double atol, double min_step_factor,
                     double cut_off, uint32_t flags, double *out7_host, int64_t *accepted_host,
                     int8_t *status_host, int64_t *total_accepted_host, void *stream);

/*
 * Host half of rwrt_real2d_host on its own (pure host code, no device needed): writes rows [row_begin, row_end) of
 * slots [slot_begin, slot_end) of the seven planes out7_host (7, nt, nslot) from the packed rows of one time window.
 * packed: (7, row_end - row_begin, nray) doubles = planes (lon, lat, k, l, A, u_g, v_g) of the nray >= 1 live rays
 * (NaN in all planes once a ray is dead, wr.py:838-855); slot_to_ray[slot] = index of the slot's ray, < 0 for a slot
 * without a real root (its planes 0..4 repeat row 0 of out7_host, u_g and v_g are NaN: wr.py:768-781 and SURVEY
 * appendix C-7).  row_begin >= 1: row 0 is the t = 0 state and must already be in out7_host.
 */
int rwrt_host_expand_rows(double *out7_host, int nt, int64_t nslot, const int32_t *slot_to_ray, const double *packed,
                          int64_t nray, int row_begin, int row_end, int64_t slot_begin, int64_t slot_end);

/* Releases the handle's scratch memory (ray records -- 64 B per live ray and output row of a time window --, ray
 * state, work lists, pinned staging buffers of rwrt_real2d_host).  The next run allocates again.  The fields stay. */
int rwrt_trim(rwrt_handle h);

/* counters of the last rwrt_ray_run_rk45 on this handle (host side, after a sync):
 * [0] kernels launched, [1] live rays, [2] ms of the whole run (first kernel after the live-ray compaction to the last one)
 * by CUDA events on the launching stream; rwrt_last_run_times splits it */
int rwrt_last_run_stats(rwrt_handle h, double *stats3);

/* CUDA-event times of the last run on this handle (ms, after a sync): [0] whole run (as stats3[2]), [1] the integrate
 * kernel alone (sum over its time windows) -- the dominant kernel bench.py's roofline is quoted on --, [2] t = 0 work
 * before the first window (rk45_prepare), [3] the rest (expand_records / pack_records / grid_records between and after
 * the windows). */
int rwrt_last_run_times(rwrt_handle h, double *ms4);

/* the last rwrt_real2d_host on this handle: [0] bytes copied device -> host, [1] bytes host -> device, [2] time windows,
 * [3] host threads that wrote result planes (0: the device expanded all of them), [4] windows whose device-expanded rows
 * were copied straight into the caller's pinned planes */
int rwrt_last_host_stats(rwrt_handle h, double *stats5);

/* Timeline of the last host-buffer call (rwrt_real2d_host / rwrt_ray_run_rk45_d2h with host-side expansion), one row of
 * RWRT_HOST_WINDOW_FIELDS doubles per time window, all times in ms since the call's pipeline started (device events and
 * the host clock share the origin): [0] route (0 packed rows + host threads, 1 direct copy of device-expanded rows),
 * [1] bytes on the link, [2] device: window ready to travel, [3] host: copy queued, [4] copy could start (link free,
 * data ready, queued), [5] device: copy finished, [6] host threads: first chunk of the window started (-1: direct),
 * [7] host threads: window completely in the caller's arrays (-1: direct).  [5]-[4] with [1] gives the link rate of
 * a window, [7]-[6] the host threads' time on it, [6]-[5] how long it waited for them.  *nwindows = rows available. */
/* Host-clock ms since entry into the last rwrt_real2d_host at the end of each phase: [0] device arena allocated,
 * [1] inputs, BS.ready and t = 0 state queued, [2] origin of the window timeline (live rays compacted, t = 0 row on the
 * host, events created), [3] all windows queued, [4] device work and copies finished, [5] host threads joined,
 * [6] per-call objects released (= return); [7] ms spent starting the host threads. */
int rwrt_last_host_phases(rwrt_handle h, double *ms8);
#define RWRT_HOST_WINDOW_FIELDS 8
int rwrt_last_host_windows(rwrt_handle h, double *rows, int max_windows, int *nwindows);

/* Roofline denominator: measured FP64 FMA throughput of the device (TFLOP/s, register-resident DFMA
 * chains).  Diagnostic only; used by bench.py next to MEASURED_PEAKS.json, which has no FP64 figure. */
int rwrt_measure_fp64_peak(rwrt_handle h, double *tflops);

/*
 * SHSF.SHSF(data, truncation_level, sampling) (SHSF.py:10-33): the reference's optional pre-filter of the background
 * winds -- spherical-harmonic analysis of a Driscoll-Healy grid (pyshtools SHExpandDH), triangular truncation
 * (SHCoeffs.pad) and synthesis (MakeGridDH, whose default sampling = 1 makes the OUTPUT an n x n grid with
 * n = 2 * truncation + 2, whatever the input was).  pyshtools is an un-vendored dependency: parity unpinned, the
 * published algorithm is restated.  A plan holds the transform tables for one (nlat, truncation).
 *   grid: (nmember, nlat, in_sampling * nlat) float64 device, row 0 = north pole, longitude 0 first;
 *   out : (nmember, nlat_out, nlon_out) float64 device, nlat_out = 2 * truncation + 2, nlon_out = out_sampling * nlat_out.
 */
typedef struct rwrt_shsf_plan_s *rwrt_shsf_plan;
int rwrt_shsf_plan_create(int device, int nlat, int in_sampling, int truncation, int out_sampling, rwrt_shsf_plan *out);
int rwrt_shsf_out_shape(rwrt_shsf_plan p, int *nlat_out, int *nlon_out);
int rwrt_shsf_apply(rwrt_shsf_plan p, const double *grid, int nmember, double *out, void *stream);
int rwrt_shsf_plan_destroy(rwrt_shsf_plan p);

#ifdef __cplusplus
}
#endif
#endif /* RWRT_H */
