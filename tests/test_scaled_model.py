"""CPU tier: the kernels' arithmetic (tests/scaled_model.py, float32) meets the parity bar
against the float64 reference fixtures.  Tolerance = BASELINE.json north_star: 1e-4 relative,
measured as max|a-b| / max|ref| per tensor; scores of normalised inputs are ~0, so
scores are measured against max(|ref|, 1)."""
import numpy as np
import pytest

from scaled_model import asg_scaled

TOL = 1e-4


def _rel(a, ref, floor=1e-30):
    fin = np.isfinite(ref)
    assert (np.isfinite(a) == fin).all()
    assert (a[~fin] == ref[~fin]).all()
    if not fin.any():
        return 0.0
    return np.abs(a[fin] - ref[fin]).max() / max(np.abs(ref[fin]).max(), floor)


@pytest.mark.parametrize("dtype,tol", [(np.float32, TOL), (np.float64, 1e-10)])
def test_scaled_model_vs_reference_fixtures(golden_cases, dtype, tol):
    for name, c in golden_cases.items():
        r = asg_scaled(c["x"], c["y"], c["A"], c["il"], c["tl"], c["g"], -c["g"], dtype=dtype)
        for k in ("full", "aligned", "loss", "gx", "gA"):
            floor = 1.0 if k in ("full", "aligned", "loss") else 1e-30
            err = _rel(np.asarray(r[k], dtype=np.float64), c["ref_" + k], floor)
            assert err <= tol, (name, k, err)
