"""CPU tests of the boundary: the C-ABI library builds, loads and exports every symbol that
include/rwrt.h declares; the facade mirrors the reference's interface and error behaviour and fails
loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import rwrt_b200  # noqa: F401
from rwrt_b200 import build as rbuild
from rwrt_b200 import lib
from rwrt_b200.wr import WR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    rbuild.build()
    return lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "rwrt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rwrt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(L):
    syms = header_symbols()
    assert set(syms) == set(lib.SYMBOLS)
    for s in syms:
        assert hasattr(L, s), s
    assert L.rwrt_abi_version() == 2      # 2: sweep / records entry points (round 2)


def test_no_compute_without_gpu(L):
    if lib.device_count() > 0:
        pytest.skip("GPU present")
    h = lib.vp()
    rc = L.rwrt_create(0, 144, 73, 1, 1, C.byref(h))
    assert rc == -2 and b"no CPU fallback" in L.rwrt_last_error()
    w = WR(3, 1, 7200.0, 86400.0, nx=144, ny=73)
    with pytest.raises(lib.RwrtError):
        w.bs.ready(xcyclic=True)


def test_product_does_not_import_oracle():
    """The oracle is the checker, never the product: no import / include / dlopen of anything under oracle/."""
    pkg = os.path.join(ROOT, "rossby-wave-ray-tracing_b200")
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|(#include\s+\"[^\"]*oracle)|(librwrt_oracle)|(rwrt_oracle\.c)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert not bad.search(txt), f"{f} reaches into oracle/"


def test_wr_host_logic_matches_reference_semantics():
    w = WR(3, 6, 2 * 3600.0, 30 * 86400.0, nx=144, ny=73)
    assert w.nt == 361                                   # wr.py:157
    assert w.cut_off[0] == 0.1 * 7200.0 / 3600.0         # wr.py:170
    with pytest.raises(ValueError):
        w.set_zwn([1.0, 2.0])                            # wr.py:216-217
    with pytest.raises(ValueError):
        w.set_source_array([1.0], [2.0])                 # wr.py:229-230
    with pytest.raises(ValueError):
        w.set_source_matrix(0.0, 0.0, 1, 1, 2, 2)        # wr.py:242-243
    with pytest.raises(ValueError):
        w.set_source_matrix(0.0, 80.0, 1, 10, 3, 2)      # wr.py:245-246
    w.set_source_matrix(350.0, -4.0, 4, 2, 3, 2)
    # the reference's double loop (wr.py:249-258), restated
    exp_lon, exp_lat = [], []
    for iy in range(2):
        for ix in range(3):
            exp_lon.append(((350.0 % 360.0 + ix * 4) % 360.0) * (np.pi / 180.0))
            exp_lat.append((-4.0 + iy * 2) * (np.pi / 180.0))
    assert np.array_equal(w.source_lon, np.array(exp_lon)) and np.array_equal(w.source_lat, np.array(exp_lat))
    assert w.rlon.shape == (361, 3, 6, 3) and np.isnan(w.rlon).all()   # wr.py:160-167
    with pytest.raises(ValueError):
        WR(3, 1)                                          # wr.py:139-141: nx/ny or ncfile needed


def test_ray_run_rejects_non_gpu_paths():
    w = WR(1, 1, 7200.0, 86400.0, nx=144, ny=73)
    for kw in (dict(mode="original"), dict(mode="original", inte_method="rk45"), dict(mode="numpy", debug=True)):
        with pytest.raises(NotImplementedError):
            w.ray_run(**kw)


def test_ncio_roundtrip(tmp_path):
    from rwrt_b200 import ncio, synthetic
    u, v = synthetic.jet_z()
    lat_deg = np.linspace(90, -90, 73).astype("float32")       # descending: must be flipped (bs.py:251-256)
    lon_deg = (np.arange(144) * 2.5).astype("float32")
    p = str(tmp_path / "uv.npz")
    np.savez(p, u=u.T[::-1], v=v.T[::-1], lat=lat_deg, lon=lon_deg)
    d = ncio.read_uv(p, 144, 73)
    assert np.array_equal(d["u"], u) and np.array_equal(d["v"], v)
    assert ncio.grid_shape(p) == (73, 144)
    # coordinates carry float32 precision (bs.py:222-236)
    assert np.array_equal(d["lon"], (lon_deg * np.pi / 180).astype("float64"))


def _expand_numpy(out, inv, packed, b, e):
    """Restatement of the packed rows -> plane expansion (wr.py:768-781 for the constant rows of no-root slots)."""
    for s, li in enumerate(inv):
        for it in range(b, e):
            if li < 0:
                out[:5, it, s] = out[:5, 0, s]
                out[5:, it, s] = np.nan
            else:
                out[:, it, s] = packed[:, it - b, li]


@pytest.mark.parametrize("isa", ["sse2", "avx2", "avx512"])
@pytest.mark.parametrize("nslot,nt,b,e", [(1, 2, 1, 2), (37, 9, 1, 9), (301, 40, 17, 36), (64, 20, 3, 19), (1030, 7, 2, 6),
                                          (9001, 4, 1, 3)])
def test_host_expand_rows_equals_restatement(L, monkeypatch, isa, nslot, nt, b, e):
    """rwrt_host_expand_rows (the host half of rwrt_real2d_host; pure host code) against a NumPy restatement: ragged
    slot counts (unaligned lines, every store width the CPU has), no-root slots, dead rays (NaN rows), sub-ranges of
    slots longer than one task chunk, untouched rows outside the window."""
    monkeypatch.setenv("RWRT_HX_ISA", isa)    # the library never picks a wider path than the CPU supports
    rng = np.random.default_rng(nslot * 1000 + nt)
    inv = np.full(nslot, -1, dtype=np.int32)
    live = rng.random(nslot) < 0.55
    live[0] = True
    nlive = int(live.sum())
    inv[live] = rng.permutation(nlive).astype(np.int32)      # any slot -> ray map
    packed = rng.standard_normal((7, e - b, nlive))
    packed[:, (e - b) // 2:, ::5] = np.nan                   # rays that died inside the window
    out = np.full((7, nt, nslot), -7.0)
    out[:, 0, :] = rng.standard_normal((7, nslot))
    out[3:5, 0, ~live] = np.nan
    want = out.copy()
    _expand_numpy(want, inv, packed, b, e)
    # in two slot ranges, like two worker threads
    cut = nslot // 3
    for s0, s1 in ((0, cut), (cut, nslot)):
        lib.check(L.rwrt_host_expand_rows(lib.ptr(out), nt, nslot, lib.ptr(inv), lib.ptr(packed), nlive, b, e, s0, s1))
    assert np.array_equal(out, want, equal_nan=True)
    # argument errors: row 0 is not produced here; rows must exist
    assert L.rwrt_host_expand_rows(lib.ptr(out), nt, nslot, lib.ptr(inv), lib.ptr(packed), nlive, 0, e, 0, nslot) != 0
    assert L.rwrt_host_expand_rows(lib.ptr(out), nt, nslot, lib.ptr(inv), lib.ptr(packed), nlive, 1, nt + 1, 0, nslot) != 0


def test_stage_loop_of_the_integrator_has_no_divergence_guards(L):
    """Code-generation guard (CPU, via cuobjdump; tools/sass_loop.py prints the same view): the stage loop of
    fast::rk45_integrate_coop must contain no WARPSYNC.ALL / BRA.DIV.  When ptxas cannot prove the warp converged there (an early `continue`, a `trap` or a second
    exit test in the refill section did that in round 2), it guards every __syncwarp / tcgen05 instruction of the loop and
    the kernel loses 9 % with bit-identical results -- invisible to every parity test."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    for out_mode in (1, 2, 3):
        fn = f"_ZN4rwrt4fast19rk45_integrate_coopILi{out_mode}EEEvNS_9RunParamsE"
        sass = subprocess.run([cuobjdump, "-sass", "-fun", fn, lib.LIB_PATH], capture_output=True, text=True).stdout
        rows = [(int(m.group(1), 16), m.group(2)) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", sass)]
        assert len(rows) > 2000, "integrator SASS not found"
        sttm = [i for i, (_, t) in enumerate(rows) if "STTM" in t]
        assert sttm, "the stage vectors no longer go to tensor memory?"
        lo = hi = None
        for i in range(sttm[0], min(sttm[0] + 80, len(rows))):          # the loop's back edge follows the STTM of stage s
            m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", rows[i][1])
            if m and int(m.group(1), 16) < rows[i][0]:
                lo, hi = int(m.group(1), 16), rows[i][0]
                break
        assert lo is not None, "stage loop not found"
        body = [t for a, t in rows if lo <= a <= hi]
        assert 500 < len(body) < 800, len(body)
        guards = [t for t in body if "WARPSYNC" in t or "BRA.DIV" in t]
        assert not guards, f"OUT={out_mode}: {len(guards)} divergence guards in the stage loop: {guards[:3]}"
        assert sum("DFMA" in t for t in body) > 100 and sum("LDGSTS" in t for t in body) >= 3
        # ... and the properties the late round-2 kernel versions were measured on (profiles/r2_integrator_experiments.txt),
        # each worth 1-5 % with bit-identical results:
        assert len(body) < 600, f"stage loop grew to {len(body)} instructions (v23: 527; an unrolled copy loop was 733 and 4.5 % slower)"
        assert sum("LDGSTS" in t for t in body) == 3, "the gather's copy loop must stay rolled (three cp.async per trip)"
        assert sum("LDS.128" in t for t in body) == 22 and sum("LDS.64" in t for t in body) <= 1, \
            "the staged cell is read with 22 LDS.128 (an LDS.64 conflicts two-way on the 400-byte rows)"
        assert not any("BRX" in t for t in body), "the stage combinations go through direct branches, not a jump table"
        assert not any("DSETP.MAX" in t or "DSETP.MIN" in t for t in body), "no emulated fp64 min / max in the stage loop"
