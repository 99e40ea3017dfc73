"""Import stub used ONLY by oracle/refshim.py when it imports the read-only
reference from /root/reference inside the build container (netCDF4 is not
installed there).  Test infrastructure, never on the product path."""


class Dataset:  # pragma: no cover - constructing it is always an error here
    def __init__(self, *a, **k):
        raise RuntimeError("netCDF4 stub: file I/O is not available in the oracle shim")
