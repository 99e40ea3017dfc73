"""Re-pins the committed golden fixtures against the LIVE, unmodified reference whenever it is mounted
(/root/reference in the build container; absent on the GPU box -> skipped there).

`oracle/make_golden.py --check` regenerates a fixture from the reference into a scratch directory and
compares it with the committed file array by array, bit for bit.  The quick fixtures are re-checked here on
every CPU run; `python -m oracle.make_golden --check` (no names) re-checks all of them (~3 minutes).
"""
import pytest

from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="the reference is not mounted at /root/reference")


@pytest.mark.parametrize("name", ["ky_frac", "kat", "pointwise", "clim_g2"])
def test_committed_fixture_is_what_the_live_reference_produces(name):
    from oracle import make_golden
    bad = make_golden.check([name], verbose=False)
    assert not bad, f"tests/golden/{name}.npz differs from the live reference in {bad}"
