"""The product's LM state machine (varpro_b200/csrc/lm_step.cuh, compiled for the host by
tests/csrc/lm_harness.cpp) against the oracle's literal MINPACK lmder on the explicit Jacobian.

lm_step.cuh restates lmder on the q x q system (H = J^T J, g = J^T r, ||r||); fed with H, g computed
from the oracle's explicit J and r it must walk the same iterates as the oracle's lmder."""
import numpy as np
import pytest

import lm_harness as LH
import workloads as W


@pytest.mark.parametrize("name", ["c1", "mrhs20_2", "mrhs20_3", "lmfit", "lmfit_w", "oleary", "c2_16"])
def test_same_minimiser_and_evaluation_count(name):
    wl = {"c1": W.c1, "mrhs20_2": lambda: W.mrhs20(2), "mrhs20_3": lambda: W.mrhs20(3),
          "lmfit": lambda: W.lmfit_case(False), "lmfit_w": lambda: W.lmfit_case(True), "oleary": W.oleary,
          "c2_16": lambda: W.c2(S=16)}[name]()
    ref = W.make_oracle(wl)
    rep = ref.fit()
    h, trace = LH.fit_with_oracle_evals(W.make_oracle(wl), wl["alpha0"])
    assert h.termination in (2, 3, 4, 5, 6) and rep["successful"]
    a, b = np.sort(h.accepted()), np.sort(ref.params())
    assert np.max(np.abs(a - b) / np.abs(b)) <= 1e-8
    # zero-residual problems end in a rounding-noise regime where the count may differ by a few
    assert abs(h.nfev - rep["number_of_evaluations"]) <= 4


def test_iterates_match_lmder_on_noisy_data():
    """On data with a non-zero residual the iterate sequence is deterministic: same trial points."""
    wl = W.lmfit_case(False)
    h, trace = LH.fit_with_oracle_evals(W.make_oracle(wl), wl["alpha0"])
    ref = W.make_oracle(wl)
    rep = ref.fit()
    assert h.nfev == rep["number_of_evaluations"]
    assert np.allclose(h.accepted(), ref.params(), rtol=1e-12, atol=0)


def test_options_stepbound_patience_and_termination_codes():
    wl = W.c1()
    h, _ = LH.fit_with_oracle_evals(W.make_oracle(wl), wl["alpha0"], patience=1)  # maxfev = 3
    assert h.termination == 8 and h.nfev == 3  # LostPatience
    h, _ = LH.fit_with_oracle_evals(W.make_oracle(wl), wl["alpha0"], stepbound=1.0)
    assert h.termination in (2, 3, 4, 5, 6)
    # failed evaluation at the start (residuals() is None) -> User, like the crate; a NaN norm of a valid residual -> Numerical
    hh = LH.LmHarness([1.0, 2.0])
    assert hh.advance(np.nan, np.zeros(2), np.zeros((2, 2)), finite=False) is False and hh.termination == 0
    hh = LH.LmHarness([1.0, 2.0])
    assert hh.advance(np.nan, np.zeros(2), np.eye(2), finite=True) is False and hh.termination == 1
    # exactly zero residual -> ResidualsZero
    hh = LH.LmHarness([1.0, 2.0])
    assert hh.advance(0.0, np.zeros(2), np.eye(2), finite=True) is False and hh.termination == 2
    # gradient orthogonal at the start -> Orthogonal
    hh = LH.LmHarness([1.0, 2.0])
    assert hh.advance(1.0, np.zeros(2), np.eye(2), finite=True) is False and hh.termination == 3


def test_rank_deficient_normal_matrix_is_handled():
    """A singular J^T J (one parameter without influence) must not produce NaNs."""
    hh = LH.LmHarness([1.0, 2.0])
    H = np.array([[4.0, 0.0], [0.0, 0.0]])
    g = np.array([2.0, 0.0])
    assert hh.advance(10.0, g, H, finite=True)
    t = hh.trial()
    assert np.all(np.isfinite(t)) and t[1] == 2.0 and abs(t[0] - 0.5) < 1e-12


@pytest.mark.parametrize("name", ["c1", "lmfit_w", "oleary", "c2_16"])
def test_q_specialised_step_is_bitwise_the_generic_step(name):
    """lm_step.cuh runs q = 1..4 on register-resident, fully unrolled instantiations and q > 4 on the
    run-time-indexed one: same floating-point operations in the same order => identical state words
    after every evaluation."""
    wl = {"c1": W.c1, "lmfit_w": lambda: W.lmfit_case(True), "oleary": W.oleary, "c2_16": lambda: W.c2(S=16)}[name]()
    op = W.make_oracle(wl)
    a = LH.LmHarness(wl["alpha0"])
    b = LH.LmHarness(wl["alpha0"], generic=True)
    x = np.asarray(wl["alpha0"], dtype=np.float64)
    more, steps = True, 0
    while more:
        assert op.set_params(x)
        r, J = op.residuals(), op.jacobian()
        more = a.advance(r @ r, J.T @ r, J.T @ J)
        assert b.advance(r @ r, J.T @ r, J.T @ J) == more
        assert a.state_bytes() == b.state_bytes(), f"states differ after evaluation {steps + 1}"
        x = a.trial()
        steps += 1
    assert steps >= 5 and a.termination in (2, 3, 4, 5, 6)
