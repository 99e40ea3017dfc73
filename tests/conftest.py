"""pytest configuration: the `gpu` marker (tests that need a real B200) and import paths."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU-marked tests skip (instead of failing in lib.require_gpu) on a box without a CUDA device, so that a plain
    `pytest tests` on the build container is green; on the GPU box nothing is skipped."""
    import pytest
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    try:
        import rwrt_b200  # noqa: F401
        from rwrt_b200 import lib
        ndev = lib.device_count()
    except Exception:
        ndev = 0
    if ndev < 1:
        skip = pytest.mark.skip(reason="no CUDA device visible (GPU tests run on the B200 box with -m gpu)")
        for it in gpu_items:
            it.add_marker(skip)
