// rwrt_internal.h -- launcher declarations shared between the translation units of librwrt.so.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "rwrt_device.cuh"

namespace rwrt {

struct RunParams;

// rwrt_rk45_{fast,strict}.cu
void launch_rk45_setup_fast(const RunParams &p, int prepare_blocks, int fill_blocks, cudaStream_t s, cudaStream_t aux);
void launch_rk45_setup_strict(const RunParams &p, int prepare_blocks, int fill_blocks, cudaStream_t s, cudaStream_t aux);
void launch_rk45_window_fast(const RunParams &p, int integrate_blocks, cudaStream_t s);
void launch_rk45_window_strict(const RunParams &p, int integrate_blocks, cudaStream_t s);
void launch_expand_records_fast(const RunParams &p, cudaStream_t s);
void launch_expand_records_strict(const RunParams &p, cudaStream_t s);
void launch_rk4_fast(const RunParams &p, cudaStream_t s);
void launch_rk4_strict(const RunParams &p, cudaStream_t s);
int occupancy_rk45_fast(bool coop);
int threads_rk45_fast(bool coop);
int threads_rk45_strict(bool coop);
int occupancy_rk45_strict(bool coop);
void launch_rhs_fast(const Grid &g, long long n, const double *y, const int32_t *member, double *dy, cudaStream_t s);
void launch_rhs_strict(const Grid &g, long long n, const double *y, const int32_t *member, double *dy, cudaStream_t s);

// rwrt_bs.cu (compiled with -fmad=false: bit-exact with the reference's NumPy arithmetic)
struct BsArgs {
    int nlon, nlat, nmember, xcyclic;
    const float *u, *v;            // (nmember, nlon, nlat)
    const double *coslat, *sinlat; // (nlat)
    double *q, *qxx, *qyy, *qxy;   // scratch planes (nmember, nlon, nlat): q and the UNSMOOTHED second derivatives
    double *fpk;                   // packed (nmember, nlon_f, nlat, 12)
    double *fields18;              // optional (nmember, nlon_f, nlat, 18)
    double *diag9;                 // optional (nmember, 9, nlon, nlat)
};
void launch_bs_ready(const BsArgs &a, cudaStream_t s);
void launch_trig_tables(int nlat, const double *lat, double *coslat, double *sinlat, cudaStream_t s);
void launch_pack_from_fields18(int nmember, int nlon_f, int nlat, const double *fields18, double *fpk, cudaStream_t s);

// rwrt_init.cu (compiled with -fmad=false: the cubic / QR arithmetic follows LAPACK operation by operation)
void launch_ray_init(const Grid &g, int nsource, const double *src_lon, const double *src_lat, const int32_t *src_member,
                     int nzwn, const double *zwn, double freq, int nt, double *out7, cudaStream_t s);
void launch_sample_mercator(const Grid &g, long long n, const double *lon, const double *lat, const int32_t *member,
                            double *out, cudaStream_t s);

void launch_sample18(const Grid &g, const double *f18, long long n, const double *lon, const double *lat,
                     const int32_t *member, double *out, cudaStream_t s);

// rwrt_grid.cu
void launch_grid_accumulate(long long nslot, int nt, const double *out7, double dlon_deg, double dlat_deg, int ncx,
                            int ncy, long long *count, long long *fx, long long *fy, cudaStream_t s);

void launch_grid_records(const double *rec, int rec_rows, const int32_t *rows_done, long long nlive, int it_begin, int it_end,
                         double dlon_deg, double dlat_deg, int ncx, int ncy, long long *count, long long *fx, long long *fy,
                         int sm_count, cudaStream_t s);

int measure_fp64_peak(int sm_count, double *tflops, cudaStream_t s);

// rwrt_api.cu helpers
void launch_compact_live(long long N, size_t plane_stride, const double *out7, int nsource, int nzwn, int32_t *live, int32_t *inv,
                         unsigned long long *nlive_dev, int *block_counts, int nblocks, cudaStream_t s);

void launch_pack_records(const double *rec, int rec_rows, const int32_t *rows_done, long long nlive, int it_begin, int it_end,
                         double *packed, cudaStream_t s);

}  // namespace rwrt
