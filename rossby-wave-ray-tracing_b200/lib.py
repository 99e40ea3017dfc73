"""ctypes binding of librwrt.so (include/rwrt.h) — the stub INTEGRATION.md shows.

There is no CPU fallback: `load()` raises if the library has not been built, and every compute
entry point raises `RwrtError` when no CUDA device is visible.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RWRT_LIB") or os.path.join(HERE, "librwrt.so")   # RWRT_LIB: A/B builds of the same ABI

MODE_FAST = 0
MODE_STRICT = 1
GRID_FROM_RECORDS = 2

RAY_NO_ROOT, RAY_ALIVE, RAY_POLE, RAY_SPEED, RAY_NAN = 0, 1, 2, 3, 4

# every symbol include/rwrt.h declares (tests/test_abi.py checks the library exports all of them)
SYMBOLS = (
    "rwrt_abi_version", "rwrt_last_error", "rwrt_device_count", "rwrt_host_alloc", "rwrt_host_free",
    "rwrt_create", "rwrt_destroy", "rwrt_trim", "rwrt_bs_ready", "rwrt_bs_set_fields", "rwrt_sample_mercator", "rwrt_sample_mercator18", "rwrt_rhs",
    "rwrt_ray_init", "rwrt_ray_run_rk45", "rwrt_ray_run_rk45_d2h", "rwrt_ray_run_rk4", "rwrt_grid_accumulate", "rwrt_real2d_host", "rwrt_host_expand_rows", "rwrt_last_host_stats", "rwrt_last_run_stats", "rwrt_measure_fp64_peak",
    "rwrt_ray_sweep_grid", "rwrt_ray_run_rk45_records", "rwrt_records_info", "rwrt_records_expand",
    "rwrt_grid_accumulate_records", "rwrt_last_run_times", "rwrt_last_host_windows", "rwrt_last_host_phases",
    "rwrt_shsf_plan_create", "rwrt_shsf_out_shape", "rwrt_shsf_apply", "rwrt_shsf_plan_destroy",
)


class RwrtError(RuntimeError):
    pass


_lib = None
vp = C.c_void_p
i32, i64, u32, f64 = C.c_int, C.c_int64, C.c_uint32, C.c_double


def load():
    """dlopen librwrt.so and declare the signatures.  Fails loudly when the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RwrtError(f"{LIB_PATH} not built: run `python rossby-wave-ray-tracing_b200/build.py` "
                        "(there is no CPU fallback for the ray path)")
    L = C.CDLL(LIB_PATH)
    L.rwrt_abi_version.restype = i32
    L.rwrt_last_error.restype = C.c_char_p
    L.rwrt_device_count.argtypes = [C.POINTER(i32)]
    L.rwrt_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.rwrt_host_free.argtypes = [vp]
    L.rwrt_create.argtypes = [i32, i32, i32, i32, i32, C.POINTER(vp)]
    L.rwrt_destroy.argtypes = [vp]
    L.rwrt_trim.argtypes = [vp]
    L.rwrt_bs_ready.argtypes = [vp] + [vp] * 8 + [vp]
    L.rwrt_bs_set_fields.argtypes = [vp, vp, vp, vp, vp]
    L.rwrt_sample_mercator.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    L.rwrt_sample_mercator18.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp]
    L.rwrt_rhs.argtypes = [vp, i64, vp, vp, vp, u32, vp]
    L.rwrt_ray_init.argtypes = [vp, i32, vp, vp, vp, i32, vp, f64, i32, vp, vp]
    L.rwrt_ray_run_rk45.argtypes = [vp, i32, vp, i32, i32, f64, f64, f64, f64, f64, f64, u32, vp, vp, vp, vp, vp, vp]
    L.rwrt_ray_run_rk45_d2h.argtypes = [vp, i32, vp, i32, i32, f64, f64, f64, f64, f64, f64, u32, vp, vp, vp, vp, vp, vp, vp]
    L.rwrt_ray_run_rk4.argtypes = [vp, i32, vp, i32, i32, f64, f64, u32, vp, vp, vp]
    L.rwrt_grid_accumulate.argtypes = [vp, i64, i32, vp, f64, f64, i32, i32, vp, vp, vp, vp]
    L.rwrt_ray_sweep_grid.argtypes = ([vp, i32, vp, vp, vp, i32, vp, f64, i32] + [f64] * 6 + [u32, f64, f64, i32, i32] +
                                      [vp] * 7)
    L.rwrt_ray_run_rk45_records.argtypes = [vp, i32, vp, i32, i32, f64, f64, f64, f64, f64, f64, u32, vp, vp, vp, vp, vp, vp]
    L.rwrt_records_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i32)] + [C.POINTER(vp)] * 4
    L.rwrt_records_expand.argtypes = [vp, i32, i32, vp, vp]
    L.rwrt_grid_accumulate_records.argtypes = [vp, f64, f64, i32, i32, vp, vp, vp, vp]
    L.rwrt_last_run_times.argtypes = [vp, C.POINTER(f64)]
    L.rwrt_last_host_phases.argtypes = [vp, C.POINTER(f64)]
    L.rwrt_last_host_windows.argtypes = [vp, C.POINTER(f64), i32, C.POINTER(i32)]
    L.rwrt_real2d_host.argtypes = ([vp] + [vp] * 6 + [i32, vp, vp, vp, i32, vp, f64, i32] + [f64] * 6 +
                                   [u32, vp, vp, vp, vp, vp])
    L.rwrt_host_expand_rows.argtypes = [vp, i32, i64, vp, vp, i64, i32, i32, i64, i64]
    L.rwrt_last_run_stats.argtypes = [vp, C.POINTER(f64)]
    L.rwrt_last_host_stats.argtypes = [vp, C.POINTER(f64)]
    L.rwrt_measure_fp64_peak.argtypes = [vp, C.POINTER(f64)]
    L.rwrt_shsf_plan_create.argtypes = [i32, i32, i32, i32, i32, C.POINTER(vp)]
    L.rwrt_shsf_out_shape.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    L.rwrt_shsf_apply.argtypes = [vp, vp, i32, vp, vp]
    L.rwrt_shsf_plan_destroy.argtypes = [vp]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ("rwrt_last_error",):
            fn.restype = i32
    _lib = L
    return L


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().rwrt_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise RwrtError(f"{what}: {msg} (status {rc})")


def device_count() -> int:
    n = i32(0)
    rc = load().rwrt_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def require_gpu():
    if device_count() < 1:
        raise RwrtError("no CUDA device visible: rwrt_b200 has no CPU fallback for the ray-integration path")


def ptr(t):
    """Device (or pinned host) pointer of a torch tensor / numpy array / None as c_void_p."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return vp(t.data_ptr())
    return vp(t.ctypes.data)


def stream_ptr(stream=None):
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return vp(s.cuda_stream)
