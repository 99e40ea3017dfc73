"""File formats either side of the ray path (reference: `bs.py:202-262, 461-511`, `wr.py:173-212, 916-959`).

The reference uses netCDF4 (HDF5).  This image has no netCDF4, so: netCDF4 is used when it can be
imported; otherwise NetCDF-3 (64-bit offset) through `scipy.io.netcdf_file`; `.npz` is always accepted.
Variable names, dimension order and units follow the reference, including its quirks: the bsfile's
lon/lat are written in RADIANS under a degrees label (bs.py:473-477) and ray lon/lat in degrees.
"""
from __future__ import annotations

import numpy as np

from .constants import pi

LAT_NAMES = ("lat", "latitude", "Lat", "Latitude")     # bs.py:214
LON_NAMES = ("lon", "longitude", "Lon", "Longitude")   # bs.py:215


def _open_read(path):
    """Returns (dict name -> ndarray, closer)."""
    if str(path).endswith(".npz"):
        z = np.load(path)
        return {k: z[k] for k in z.files}, lambda: None
    try:
        import netCDF4
        ds = netCDF4.Dataset(path)
        return {k: np.array(ds.variables[k][:]) for k in ds.variables}, ds.close
    except ImportError:
        from scipy.io import netcdf_file
        ds = netcdf_file(path, "r", mmap=False)
        return {k: np.array(ds.variables[k][:]) for k in ds.variables}, ds.close


def grid_shape(path):
    """wr.py:173-212: (nlat, nlon) from the coordinate variables, else from u's last two dims."""
    v, close = _open_read(path)
    try:
        lat = next((v[n] for n in LAT_NAMES if n in v), None)
        lon = next((v[n] for n in LON_NAMES if n in v), None)
        if lat is None or lon is None:
            print('!!!WARNING: Using u.shape[-2] and u.shape[-1] as nlat and nlon!!!')
            return v["u"].shape[-2], v["u"].shape[-1]
        return lat.shape[0], lon.shape[0]
    finally:
        close()


def read_uv(path, nlon, nlat, read_dtype="float32", dx=None, dy=None):
    """bs.py:202-262: u, v as (nlon, nlat) `read_dtype`; lon/lat radians with float32 precision when the
    file has coordinates (they are read as float32 before the conversion, bs.py:222-236)."""
    v, close = _open_read(path)
    try:
        tu = np.array(v["u"], dtype=read_dtype)
        tv = np.array(v["v"], dtype=read_dtype)
        lat_data = next((np.array(v[n], dtype=read_dtype) for n in LAT_NAMES if n in v), None)
        lon_data = next((np.array(v[n], dtype=read_dtype) for n in LON_NAMES if n in v), None)
    finally:
        close()
    dx = np.array([2.0 * pi / nlon]) if dx is None else dx
    dy = np.array([pi / (nlat - 1)]) if dy is None else dy
    lat = (lat_data * pi / 180).astype("float64") if lat_data is not None else -pi * 0.5 + np.arange(nlat) * dy
    lon = (lon_data * pi / 180).astype("float64") if lon_data is not None else np.arange(nlon) * dx
    u = np.transpose(tu, (1, 0))
    vv = np.transpose(tv, (1, 0))
    if lat_data is None or lon_data is None:
        print('###WARNING: lon and lat not found. Make sure your lats are from 90S to 90N and lons are from 0E to 360E###')
    elif lat_data[0] > lat_data[-1]:           # bs.py:251-256: flip u, v (the coordinate array is left as read)
        u = np.transpose(tu[::-1, :], (1, 0))
        vv = np.transpose(tv[::-1, :], (1, 0))
    return {"u": np.ascontiguousarray(u), "v": np.ascontiguousarray(vv), "lon": lon, "lat": lat}


def _writer(path):
    if str(path).endswith(".npz"):
        return None
    try:
        import netCDF4
        return ("nc4", netCDF4.Dataset(path, "w", format="NETCDF4"))
    except ImportError:
        from scipy.io import netcdf_file
        return ("nc3", netcdf_file(path, "w", version=2))


def write_bsfile(path, lon, lat, fields):
    """bs.py:461-511.  fields: name -> (array (nlon, nlat), unit)."""
    w = _writer(path)
    if w is None:
        np.savez_compressed(path, lon=lon, lat=lat, **{k: np.asarray(a) for k, (a, _) in fields.items()})
        return
    kind, ds = w
    ds.createDimension("lon", len(lon))
    ds.createDimension("lat", len(lat))
    for name, data in (("lon", lon), ("lat", lat)):
        var = ds.createVariable(name, "f8", (name,))
        var[:] = data
        var.units = "degrees_east" if name == "lon" else "degrees_north"
    for name, (data, unit) in fields.items():
        kw = dict(zlib=True, complevel=4) if kind == "nc4" else {}
        var = ds.createVariable(name, "f8", ("lon", "lat"), **kw)
        var[:, :] = np.asarray(data, dtype="float64")
        # the reference's unit strings are kept (bs.py:499: "1/(m·s)"); NetCDF-3 attributes are ASCII-only
        var.units = unit if kind == "nc4" else unit.encode("ascii", "replace").decode("ascii").replace("?", ".")
    ds.close()


def write_rayfile(path, zwn, nsource, nt, planes, rad2deg):
    """wr.py:916-959.  planes: name -> (nt, 3, nsource, nzwn)."""
    out = {k: np.asarray(v) for k, v in planes.items()}
    out["rlon"] = out["rlon"] * rad2deg
    out["rlat"] = out["rlat"] * rad2deg
    w = _writer(path)
    if w is None:
        np.savez_compressed(path, zwn=zwn, source_index=np.arange(nsource), time_index=np.arange(nt), **out)
        return
    kind, ds = w
    ds.createDimension("zwn", len(zwn))
    ds.createDimension("source", nsource)
    ds.createDimension("root", 3)
    ds.createDimension("time", nt)
    ds.createVariable("zwn", "f8", ("zwn",))[:] = zwn
    ds.createVariable("source_index", "i4", ("source",))[:] = np.arange(nsource)
    ds.createVariable("time_index", "i4", ("time",))[:] = np.arange(nt)
    units = {"rlon": "degrees", "rlat": "degrees", "rzwn": "rad_per_meter*Rearth", "rug": "m s-1", "rvg": "m s-1"}
    for name in ("rlon", "rlat", "rzwn", "rmwn", "ramp", "rug", "rvg"):
        var = ds.createVariable(name, "f8", ("time", "root", "source", "zwn"))
        var[:] = out[name]
        if name in units:
            var.units = units[name]
    ds.close()
