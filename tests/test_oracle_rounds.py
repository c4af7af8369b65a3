"""CPU checks of the two restatements that the round-based CUDA drivers are tested against (oracle/oracle.c):
orc_hcubature_rounds_batch (ib200_hcubature_batch, mode IB200_HCUB_ROUNDS) and orc_hcubature_two_level (the second level of
ib200_hcubature_single).  Neither scheme is in the reference -- it refines one box per step (HCubature.hcubature as called at
/root/reference/src/Integrals.jl:380-393) -- so both are pinned HERE against the reference-order restatement orc_hcubature and the closed
forms: same rule, same stopping test, (I, E) within the requested tolerance, evaluation counts within a known factor."""
import numpy as np
import pytest

import oracle as O
from integrals_b200.workloads import genz_params, gm_npoints

import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util_gpu import genz_exact, ORC_FAM          # noqa: E402  (closed forms; no GPU needed)


@pytest.mark.parametrize("family", ["genz_oscillatory", "genz_product_peak", "genz_gaussian"])
@pytest.mark.parametrize("dim", [2, 4])
def test_rounds_agree_with_reference_order(family, dim):
    nprob = 12
    par = genz_params(family, dim, nprob, scale=0.25)
    lb, ub = np.zeros(dim), np.ones(dim)
    reltol, abstol = 1e-6, 1e-8
    I1, E1, ne1, st1, nr = O.hcubature_rounds_batch(ORC_FAM[family], lb, ub, par, reltol=reltol, abstol=abstol, maxiters=2_000_000, nthreads=4)
    I0, E0, ne0, st0 = O.hcubature_batch(ORC_FAM[family], lb, ub, par, reltol=reltol, abstol=abstol, maxiters=2_000_000, nthreads=4)
    tol = np.maximum(abstol, reltol * np.abs(I0))
    assert np.all(st0 == 0) and np.all(st1 == 0)
    assert np.all(E1 <= tol * (1 + 1e-12)) and np.all(np.abs(I1 - I0) <= tol)
    exact = np.array([genz_exact(family, par[:, k], dim) for k in range(nprob)])
    assert np.all(np.abs(I1 - exact) <= 10 * tol)
    assert np.all((ne1 - gm_npoints(dim)) % (2 * gm_npoints(dim)) == 0)           # 1 + 2 per bisection, like the reference
    assert ne1.sum() <= 1.5 * ne0.sum() and np.all(nr <= 80)                       # more evaluations than worst-first, O(log boxes) rounds


def test_rounds_batch_is_the_single_problem_rounds_loop():
    par = genz_params("genz_gaussian", 3, 5, scale=0.5)
    lb, ub = np.zeros(3), np.ones(3)
    I, E, ne, st, nr = O.hcubature_rounds_batch("gauss", lb, ub, par, reltol=1e-7, abstol=0.0)
    for k in range(5):
        Ik, Ek, s = O.hcubature_rounds("gauss", lb, ub, par[:, k], reltol=1e-7, abstol=0.0, max_boxes=1 << 40)
        assert (Ik, Ek, s["rule_applications"] * 33, s["rounds"]) == (I[k], E[k], ne[k], nr[k])


def test_rounds_budget_initdiv_and_degenerate():
    lb, ub = np.zeros(3), np.ones(3)
    par = genz_params("genz_gaussian", 3, 3, scale=0.5)
    I, E, ne, st, _ = O.hcubature_rounds_batch("gauss", lb, ub, par, reltol=1e-12, abstol=0.0, maxiters=5000)
    assert np.all(st == 1) and np.all(ne >= 5000)                                  # the budget ends the solve, status MAXEVALS
    I3, E3, ne3, st3, _ = O.hcubature_rounds_batch("gauss", lb, ub, par, reltol=1e-7, abstol=0.0, initdiv=3)
    I0, *_ = O.hcubature_batch("gauss", lb, ub, par, reltol=1e-7, abstol=0.0, initdiv=3)
    assert np.all(st3 == 0) and np.all(np.abs(I3 - I0) <= 1e-7 * np.abs(I0)) and np.all(ne3 >= 27 * 33)
    I, E, ne, st, _ = O.hcubature_rounds_batch("gauss", np.array([0.0, 0.5]), np.array([1.0, 0.5]), par[:4, :1])
    assert I[0] == 0.0 and ne[0] == 17 and st[0] == 0                              # HCubature returns after the first box


@pytest.mark.parametrize("dim,reltol,switch", [(3, 1e-9, 64), (4, 1e-8, 128)])
def test_two_level_reaches_the_tolerance_in_bounded_extra_work(dim, reltol, switch):
    par = genz_params("genz_gaussian", dim, 3, scale=0.5)[:, 1]
    lb, ub = np.zeros(dim), np.ones(dim)
    I2, E2, s2 = O.hcubature_two_level("gauss", lb, ub, par, reltol=reltol, abstol=0.0, switch_boxes=switch)
    I1, E1, s1 = O.hcubature_rounds("gauss", lb, ub, par, reltol=reltol, abstol=0.0, max_boxes=1 << 40)
    exact = genz_exact("genz_gaussian", par, dim)
    assert s2["status"] == 0 and s2["level2_subproblems"] > 0 and E2 <= reltol * abs(I2)         # HCubature's stopping test holds for the whole integral
    assert abs(I2 - I1) <= reltol * abs(I1) and abs(I2 - exact) <= 10 * reltol * abs(exact)
    assert s2["rule_applications"] <= 2.0 * s1["rule_applications"]                            # the price of bounded memory
    # no hand-over needed -> the single-level rounds, bit for bit
    I3, E3, s3 = O.hcubature_two_level("gauss", lb, ub, par, reltol=1e-4, abstol=0.0, switch_boxes=1 << 30)
    I4, E4, s4 = O.hcubature_rounds("gauss", lb, ub, par, reltol=1e-4, abstol=0.0, max_boxes=1 << 40)
    assert (I3, E3, s3["rule_applications"], s3["level2_subproblems"]) == (I4, E4, s4["rule_applications"], 0)


def test_two_level_retirement_keeps_the_answer_in_bounded_arenas():
    """converged-box retirement (sub_arena): sub-problems that outgrow their arena drop the boxes that have converged and go on.  With a
    workable arena the whole solve still meets its tolerance with about the same work as the unbounded store; the retired error is part
    of E (never lost), so a solve that cannot finish says so through E and the status."""
    par = genz_params("genz_gaussian", 4, 3, scale=0.5)[:, 0]
    lb, ub = np.zeros(4), np.ones(4)
    reltol = 1e-9
    exact = genz_exact("genz_gaussian", par, 4)
    I0, E0, s0 = O.hcubature_two_level("gauss", lb, ub, par, reltol=reltol, abstol=0.0, max_boxes=1 << 40, switch_boxes=64)
    assert s0["status"] == 0 and s0["retired_boxes"] == 0
    I1, E1, s1 = O.hcubature_two_level("gauss", lb, ub, par, reltol=reltol, abstol=0.0, max_boxes=1 << 40, switch_boxes=64, sub_arena=2048)
    assert s1["status"] == 0 and s1["retired_boxes"] > 0 and E1 <= reltol * abs(I1)
    assert abs(I1 - I0) <= reltol * abs(I0) and abs(I1 - exact) <= 10 * reltol * abs(exact)
    assert s1["rule_applications"] <= 1.1 * s0["rule_applications"]
    # an arena far too small: some sub-problems end with a full arena -- reported (status 1), E carries what is left, I stays accurate
    I2, E2, s2 = O.hcubature_two_level("gauss", lb, ub, par, reltol=reltol, abstol=0.0, max_boxes=1 << 40, switch_boxes=64, sub_arena=128)
    assert s2["status"] == 1 and s2["retired_boxes"] > s1["retired_boxes"] and E2 > E1
    assert abs(I2 - exact) <= E2
