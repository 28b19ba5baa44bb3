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


def test_scaled_model_peaked_emissions_fp64():
    """emissions scaled by 10 and 20 (span of hundreds of nats inside a block of target positions).  In float64
    the formulation covers it (the shifts are flushed / saturated at the format's own range, 2^1022, not at fp32's 2^126)."""
    from oracle import asg_oracle as O
    from scaled_model import asg_scaled
    rng = np.random.default_rng(7)
    T, N, L, S = 60, 2, 12, 10
    for scale in (10.0, 20.0):
        x = scale * rng.standard_normal((T, N, L))
        A = 0.3 * rng.standard_normal((L, L))
        y = rng.integers(0, L, (N, S))
        il = np.array([60, 52]); tl = np.array([10, 7])
        g = np.array([1.0, 0.7])
        ref = O.asg_loss(x, y, A, il, tl, g=g)
        got = asg_scaled(x, y, A, il, tl, g, -g, np.float64)
        for k in ("loss", "gx", "gA"):
            err = np.abs(got[k] - ref[k]).max() / max(np.abs(ref[k]).max(), 1.0 if k == "loss" else 1e-30)
            assert err <= 1e-9, (scale, k, err)
