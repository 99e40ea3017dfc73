"""Driver entry with the reference's signature: `real2d_hnf(**parameters)` (`main_wr.py:5-100`).

Same dict keys, same call sequence (WR -> loadbs_ncfile -> ready -> output -> set_zwn ->
set_source_matrix -> ray_info -> ray_run -> output).  `mode='numpy'` runs on the GPU with
`inte_method='rk45'` (adaptive RK45, the hot path) or `''` (fixed-step RK4, the reference's default);
`mode='original'` raises — there is no CPU fallback.
"""
import numpy as np

parameters = {
    'freq': 0.,
    'mm': None,
    'nn': None,
    'SW_lon': 70., 'SW_lat': -4.,
    'dlon': 4, 'dlat': 2,
    'nnx': 21, 'nny': 15,
    'zwn': np.array([1., 2., 3., 4., 5., 6., 7.]),
    'nzwn': 7,
    'tstep': 2,
    'ttotal': 90.,
    'mode': 'numpy',
    'root_method': 'numpy',
    'inte_method': 'rk45',   # '' = fixed-step RK4 (the reference's default), 'rk45' = adaptive
    'xcyclic': True,
    'cal_dtype': 'float64',
    'read_dtype': 'float32',
    'inputuv': 'path/to/basic_flow.nc',
    'bsfile': 'path/to/basic_state_out.nc',
    'ncfile': 'path/to/rays_out.nc',
    'rtol': 1e-6,
    'atol': 1e-6,
    'MinStepFactor': 1e-3,
}


def real2d_hnf(nzwn, mm, nn, freq, zwn, inputuv, ncfile, bsfile, SW_lon, SW_lat, dlon, dlat, nnx, nny, mode, tstep,
               ttotal, xcyclic, root_method, read_dtype, cal_dtype, inte_method, atol, rtol, MinStepFactor,
               device=None, quiet=False):
    from .constants import day, hour
    from .wr import WR

    nsource = nnx * nny
    wr1 = WR(nzwn, nsource, tstep * hour, ttotal * day, freq, nx=mm, ny=nn, read_dtype=read_dtype,
             cal_dtype=cal_dtype, rtol=rtol, atol=atol, ncfile=inputuv, MinStepFactor=MinStepFactor, device=device)
    wr1.bs.loadbs_ncfile(inputuv)
    wr1.bs.ready(xcyclic=xcyclic)
    if bsfile:
        wr1.bs.output(bsfile)
    wr1.set_zwn(zwn)
    wr1.set_source_matrix(SW_lon, SW_lat, dlon, dlat, nnx, nny)
    if not quiet:
        wr1.ray_info()
    wr1.ray_run(mode=mode, root_method=root_method, inte_method=inte_method)
    if ncfile:
        wr1.output(ncfile)
    return wr1


if __name__ == '__main__':
    real2d_hnf(**parameters)
