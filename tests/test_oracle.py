"""CPU tier: pins the numpy oracle (oracle/asg_oracle.py) against
(a) the golden vectors / known answers of the reference's own tests (tests/golden/reference_tests.json),
(b) fixtures produced by the compiled unmodified reference in float64 (tests/golden/ref_cases.npz),
(c) the compiled reference itself, live, when oracle/_ref is present."""
import math

import numpy as np
import pytest

from oracle import asg_oracle as O


def _btn(flat, B, T, L):
    return np.asarray(flat, dtype=np.float64).reshape(B, T, L).transpose(1, 0, 2)


def test_asg_4_golden(reference_tests):
    c = reference_tests["test_asg_4"]
    B, T, L, S = c["B"], c["T"], c["N_labels"], c["S"]
    x = _btn(c["inputs_BTN"], B, T, L)
    y = np.asarray(c["targets_BS"]).reshape(B, S)
    r = O.asg_loss(x, y, np.zeros((L, L)), np.asarray(c["input_lengths"]), np.asarray(c["target_lengths"]))
    assert np.abs(r["loss"] - np.asarray(c["expected_loss"])).sum() < c["tol"]["loss_abs_sum"]
    assert np.abs(r["gx"] - _btn(c["expected_input_grad_BTN"], B, T, L)).max() < c["tol"]["input_grad_max"]
    assert np.abs(r["gA"] - np.asarray(c["expected_trans_grad"]).reshape(L, L)).max() < c["tol"]["trans_grad_max"]


def test_fac_known_answers(reference_tests):
    c = reference_tests["test_fac_1"]
    x = _btn(c["inputs_BTN"], c["B"], c["T"], c["N_labels"])
    y = np.asarray(c["targets_BS"]).reshape(c["B"], c["S"])
    il, tl = np.full(2, 3), np.full(2, 2)
    s, *_ = O.fac_forward(x, y, np.zeros((2, 2)), il, tl)
    assert np.abs(s - np.asarray(c["expected_fac"])).sum() < 1e-10
    c = reference_tests["test_fac_2"]
    x = np.full((3, 1, 4), math.log(0.25))
    s, *_ = O.fac_forward(x, np.array([[0, 1]]), np.zeros((4, 4)), np.array([3]), np.array([2]))
    assert abs(s[0] - c["expected_fac"][0]) < 1e-10


def test_fcc_known_answers(reference_tests):
    c = reference_tests["test_fcc_1"]
    with np.errstate(divide="ignore"):
        x = np.log(_btn(c["probs_BTN"], c["B"], c["T"], c["N_labels"]))
    s, *_ = O.fcc_forward(x, np.zeros((2, 2)), np.full(2, 3))
    assert np.abs(s).sum() < 1e-10
    x = np.full((3, 2, 4), math.log(0.25))
    s, *_ = O.fcc_forward(x, np.zeros((4, 4)), np.full(2, 3))
    assert np.abs(s).sum() < 1e-10
    r = np.random.default_rng(0)
    p = r.uniform(size=(300, 3, 40))
    x = np.log(p / p.sum(-1, keepdims=True))
    s, *_ = O.fcc_forward(x, np.zeros((40, 40)), np.full(3, 300))
    assert np.abs(s).sum() < 1e-4


def test_asg_known_answers(reference_tests):
    c = reference_tests["test_asg_1"]
    with np.errstate(divide="ignore"):
        x = np.log(_btn(c["probs_BTN"], c["B"], c["T"], c["N_labels"]))
    y = np.asarray(c["targets_BS"]).reshape(2, 2)
    r = O.asg_loss(x, y, np.zeros((2, 2)), want_grad=False)
    assert np.abs(r["loss"] - np.asarray(c["expected_loss"])).sum() < 1e-10
    c = reference_tests["test_asg_2"]
    x = np.full((3, 1, 4), math.log(0.25))
    r = O.asg_loss(x, np.array([[0, 1]]), np.zeros((4, 4)), want_grad=False)
    assert abs(O.reduce_loss(r["loss"], "mean") - c["expected_loss_mean"]) < 1e-10
    # test_asg_3: S > T is truncated (asg.py:119-122)
    r3 = O.asg_loss(x, np.array([[0, 1, 1, 1]]), np.zeros((4, 4)), tl=np.array([4]), want_grad=False)
    r3b = O.asg_loss(x, np.array([[0, 1, 1]]), np.zeros((4, 4)), tl=np.array([3]), want_grad=False)
    assert np.isfinite(r3["loss"]).all() and abs(r3["loss"][0] - r3b["loss"][0]) < 1e-12


def _check_against(c, r, tol=1e-10):
    for k in ("loss", "full", "aligned", "gx", "gA"):
        ref = c["ref_" + k]
        fin = np.isfinite(ref)
        assert (np.isfinite(r[k]) == fin).all(), k
        assert (r[k][~fin] == ref[~fin]).all(), k
        if fin.any():
            scale = max(1.0, np.abs(ref[fin]).max())
            assert np.abs(r[k][fin] - ref[fin]).max() <= tol * scale, k


def test_oracle_matches_reference_fixtures(golden_cases):
    assert len(golden_cases) >= 16
    for name, c in golden_cases.items():
        r = O.asg_loss(c["x"], c["y"], c["A"], c["il"], c["tl"], c["g"])
        _check_against(c, r)


def test_invariants(golden_cases):
    """sum_l gx[t,n,l]=0 on valid frames of feasible utterances, gx=0 on padding."""
    c = golden_cases["mid40"]
    r = O.asg_loss(c["x"], c["y"], c["A"], c["il"], c["tl"], c["g"])
    T = c["x"].shape[0]
    for n in range(c["x"].shape[1]):
        Tn = int(c["il"][n])
        assert np.abs(r["gx"][:Tn, n].sum(-1)).max() < 1e-10
        assert (r["gx"][Tn:, n] == 0).all()
    assert abs(r["gA"].sum()) < 1e-8


def test_oracle_matches_live_reference():
    ref_driver = pytest.importorskip("oracle.ref_driver")
    try:
        ref_driver.surface()
    except ImportError:
        pytest.skip("oracle/_ref not built and /root/reference absent")
    r = np.random.default_rng(123)
    T, N, L, S = 25, 3, 11, 8
    x = r.standard_normal((T, N, L)); A = r.standard_normal((L, L)); y = r.integers(0, L, (N, S))
    tl = r.integers(1, S + 1, N); il = np.array([r.integers(tl[n], T + 1) for n in range(N)])
    g = r.uniform(0.5, 1.5, N)
    ref = ref_driver.ref_asg(x, y, A, il, tl, g)
    mine = O.asg_loss(x, y, A, il, tl, g)
    _check_against({"ref_" + k: v for k, v in ref.items()}, mine)


def test_viterbi_oracle_against_brute_force():
    """The Viterbi restatement (oracle/asg_oracle.py: viterbi_align) against an exhaustive search over every monotone alignment."""
    import itertools
    rng = np.random.default_rng(5)
    for T, S, L in [(5, 3, 4), (6, 1, 3), (4, 4, 5), (7, 4, 3)]:
        x = rng.standard_normal((T, 1, L))
        A = rng.standard_normal((L, L))
        y = rng.integers(0, L, (1, S))
        best, best_path = -np.inf, None
        for moves in itertools.combinations(range(1, T), S - 1):          # the frames at which the alignment advances
            pos = np.zeros(T, dtype=int)
            for m in moves:
                pos[m:] += 1
            lab = y[0][pos]
            sc = x[np.arange(T), 0, lab].sum() + sum(A[lab[t], lab[t - 1]] for t in range(1, T))
            if sc > best:
                best, best_path = sc, pos
        s, p = O.viterbi_align(x, y, A)
        assert abs(s[0] - best) < 1e-12 and (p[0] == best_path).all()
    # infeasible and empty cases
    s, p = O.viterbi_align(rng.standard_normal((6, 1, 4)), rng.integers(0, 4, (1, 5)), rng.standard_normal((4, 4)), np.array([3]), np.array([5]))
    assert np.isneginf(s[0]) and (p == -1).all()
