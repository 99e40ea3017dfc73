"""rwrt_b200 — B200-native (sm_100a) Li-Yang Rossby-wave ray-tracing engine.

Drop-in for the ray-integration path of yinan-codes/Rossby-wave-ray-tracing
(`main_wr.real2d_hnf` -> `WR.ray_run(mode='numpy', inte_method='rk45')`).
The compute path is hand-written CUDA behind the C ABI declared in `include/rwrt.h`
(`librwrt.so`, built from `csrc/`).  There is NO CPU fallback: every compute entry
point raises if the CUDA library is missing or no GPU is visible.
"""
from . import constants  # noqa: F401

__all__ = ["constants", "synthetic", "BS", "WR", "real2d_hnf"]


def __getattr__(name):  # lazy: keep `import rwrt_b200` cheap and torch-free
    import importlib
    if name in ("synthetic", "bs", "wr", "main_wr", "lib", "gridding", "dist", "build"):
        return importlib.import_module(f"{__name__}.{name}")
    if name == "BS":
        return importlib.import_module(f"{__name__}.bs").BS
    if name == "WR":
        return importlib.import_module(f"{__name__}.wr").WR
    if name == "real2d_hnf":
        return importlib.import_module(f"{__name__}.main_wr").real2d_hnf
    raise AttributeError(name)
