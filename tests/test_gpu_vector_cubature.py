"""GPU parity tests of the vector-valued HCubature path: ib200_hcubature_vec_batch, called through the
C ABI, against the CPU oracle orc_hcubature_vec on the same seeded inputs, against closed forms and against the scalar CUDA
path.  Tolerances: every component within max(abstol, reltol * norm(I)) of the oracle and of the closed form (the norm of the
whole vector sets the target, as in HCubature); error estimates within a factor 4; same status."""
import math

import numpy as np
import pytest

import integrals_b200 as ib
from util_gpu import F, ORC_FAM, genz_exact

pytestmark = pytest.mark.gpu
INF = math.inf


@pytest.fixture(scope="module")
def eng():
    return ib.Engine(0)


@pytest.fixture(scope="module")
def O():
    import oracle
    return oracle


def _genz_vec_params(rng, dim, ncomp, nprob, amax):
    a = rng.uniform(0.5, amax, (dim, ncomp, nprob)); u = rng.uniform(0.1, 0.9, (dim, ncomp, nprob))
    return np.asfortranarray(np.concatenate([a, u], axis=0))


def _norms(Iref, norm):
    return np.linalg.norm(Iref, axis=0) if norm == "2" else np.abs(Iref).max(axis=0)


@pytest.mark.parametrize("family,dim,ncomp,norm", [
    ("genz_gaussian", 2, 3, "2"), ("genz_gaussian", 3, 8, "inf"), ("genz_oscillatory", 2, 2, "2"), ("genz_product_peak", 3, 4, "2"),
    ("genz_corner_peak", 4, 2, "inf"), ("genz_c0", 2, 5, "2"), ("genz_discontinuous", 2, 2, "2"), ("genz_gaussian", 5, 2, "2"),
    ("genz_gaussian", 1, 4, "2"), ("genz_oscillatory", 6, 1, "inf")])
def test_hcubature_vector_valued_matches_oracle(eng, O, family, dim, ncomp, norm):
    """same rule, same subdivision order: results within the requested tolerance of the oracle, mostly identical evaluation
    counts (they differ only at rounding-level ties of the error keys / the split-axis choice)"""
    nprob, reltol, abstol, cap = 24, (1e-6 if dim <= 3 else 1e-5), 1e-10, 300000
    rng = np.random.default_rng(100 * dim + ncomp)
    par = _genz_vec_params(rng, dim, ncomp, nprob, {1: 6.0, 2: 5.0, 3: 3.5}.get(dim, 2.0))
    lb, ub = np.zeros(dim), np.ones(dim)
    I, E, ne, st = eng.hcubature_vec_batch(F[family], dim, lb, ub, par, reltol=reltol, abstol=abstol, maxevals=cap, norm=norm)
    Ir, Er, ner, str_ = O.hcubature_vec_batch(ORC_FAM[family], lb, ub, par, reltol=reltol, abstol=abstol, maxiters=cap, norm=norm,
                                              nthreads=O.max_threads())
    assert I.shape == (ncomp, nprob) and np.all(np.isfinite(I)) and np.all(np.isfinite(E))
    nI = _norms(Ir, norm)
    tol = np.maximum(abstol, reltol * nI)
    conv = (st == 0) & (str_ == 0)
    assert np.mean(st == str_) >= 0.9 and np.mean(conv) >= 0.5
    assert np.all(np.abs(I - Ir)[:, conv] <= tol[conv])
    assert np.all(np.abs(I - Ir) <= 4 * np.maximum(tol, np.maximum(E, Er)))           # at the evaluation cap: within the error estimates
    assert np.all((E[conv] <= 4 * Er[conv] + 1e-300) & (Er[conv] <= 4 * E[conv] + 1e-300))
    assert np.all(E[st == 0] <= tol[st == 0] * (1 + 1e-9))
    assert np.mean(ne == ner) >= 0.7
    same = ne == ner
    assert np.all(np.abs(I - Ir)[:, same] <= 1e-10 * np.maximum(nI[same], 1e-3))
    npts = 15 if dim == 1 else 1 + 4 * dim + 2 * dim * (dim - 1) + (1 << dim)
    assert np.all((ne - npts) % (2 * npts) == 0)
    if family in ("genz_gaussian", "genz_oscillatory", "genz_product_peak"):           # smooth: the estimate is reliable
        exact = np.array([[genz_exact(family, par[:, c, k], dim) for k in range(nprob)] for c in range(ncomp)])
        assert np.all(np.abs(I - exact)[:, conv] <= 10 * tol[conv])


def test_hcubature_vector_one_component_is_the_scalar_solver(eng):
    """ncomp = 1: the vector kernel evaluates Family::eval point by point, the scalar kernel the per-axis tables -- same
    rule and subdivision, values equal up to rounding"""
    dim, nprob = 3, 32
    rng = np.random.default_rng(3)
    par = _genz_vec_params(rng, dim, 1, nprob, 4.0)
    lb, ub = np.zeros(dim), np.ones(dim)
    for norm in ("2", "inf"):
        I, E, ne, st = eng.hcubature_vec_batch(F["genz_gaussian"], dim, lb, ub, par, reltol=1e-7, abstol=0.0, maxevals=200000, norm=norm)
        Is, Es, nes, sts = eng.hcubature_batch(F["genz_gaussian"], dim, lb, ub, par[:, 0, :], reltol=1e-7, abstol=0.0, maxevals=200000)
        assert np.array_equal(st, sts) and np.mean(ne == nes) >= 0.7
        assert np.all(np.abs(I[0] - Is) <= 1e-7 * np.abs(Is))
        same = ne == nes
        assert np.all(np.abs(I[0] - Is)[same] <= 1e-10 * np.abs(Is[same]))


def test_hcubature_vector_reference_known_answers_and_solve(eng, O):
    """test/interface_tests.jl:287-312 ("Standard Vector Integrands", HCubatureJL, dim 1..2, nout 1..2, reltol = abstol = 1e-5)
    through solve(); infinite domains under the three transformations; initdiv; per-problem bounds; argument errors"""
    for dim in (1, 2):
        lb, ub = np.ones(dim), 3 * np.ones(dim)
        for nout in (1, 2):
            par = np.zeros((3 * dim + 1, nout)); par[3 * dim] = np.arange(1, nout + 1)
            sol = ib.solve(ib.IntegralProblem(ib.VectorOf(ib.GaussPoly()), (lb, ub), par), ib.HCubatureB200(), engine=eng, reltol=1e-5, abstol=1e-5)
            assert sol.retcode == ib.ReturnCode.Success and sol.u.shape == (nout,)
            assert np.allclose(sol.u, 2.0 ** dim * np.arange(1, nout + 1), rtol=1e-13) and sol.stats["nevals"][0] == (15 if dim == 1 else 17)
            sol = ib.solve(ib.IntegralProblem(ib.VectorOf(ib.CosProduct()), (lb, ub), np.ones((dim, nout))), ib.HCubatureB200(), engine=eng,
                           reltol=1e-5, abstol=1e-5)
            assert np.all(np.abs(sol.u - (math.sin(3.0) - math.sin(1.0)) ** dim) <= 1e-5) and sol.resid <= 1e-5
    # Gaussians over R^2 and over a half plane: pi / (a1 a2) and half of it
    par = np.asfortranarray(np.array([[0.8, 1.3, 0.2, -0.4], [1.5, 0.6, 0.0, 0.3]]).T.reshape(4, 2, 1))
    exact = math.pi / np.array([0.8 * 1.3, 1.5 * 0.6])
    for tr, trn in enumerate(["if_inf", "tan", "cot"]):
        I, E, ne, st = eng.hcubature_vec_batch(F["genz_gaussian"], 2, [-INF, -INF], [INF, INF], par, reltol=1e-6, abstol=0.0, transform=tr,
                                               maxevals=400000)
        Ir, Er, ner, str_ = O.hcubature_vec_batch("gauss", [-INF, -INF], [INF, INF], par, tr=trn, reltol=1e-6, abstol=0.0, maxiters=400000)
        nrm = np.linalg.norm(exact)
        assert st[0] == 0 and str_[0] == 0 and np.all(np.abs(I[:, 0] - exact) <= 1e-5 * nrm) and np.all(np.abs(I - Ir) <= 1e-6 * nrm)
    ph = par.copy(); ph[2:, :, 0] = 0.0                         # centred: the half plane x1 >= 0 holds half of the mass
    I, E, ne, st = eng.hcubature_vec_batch(F["genz_gaussian"], 2, [0.0, -INF], [INF, INF], ph, reltol=1e-6, abstol=0.0, maxevals=400000)
    assert st[0] == 0 and np.all(np.abs(I[:, 0] - 0.5 * exact) <= 1e-5 * np.linalg.norm(exact))
    # initdiv and per-problem bounds against the oracle
    rng = np.random.default_rng(9)
    pg = _genz_vec_params(rng, 2, 3, 6, 4.0)
    lbp = np.asfortranarray(rng.uniform(-0.5, 0.0, (2, 6))); ubp = np.asfortranarray(rng.uniform(0.8, 1.5, (2, 6)))
    I, E, ne, st = eng.hcubature_vec_batch(F["genz_gaussian"], 2, lbp, ubp, pg, reltol=1e-7, abstol=0.0, initdiv=3)
    Ir, Er, ner, str_ = O.hcubature_vec_batch("gauss", lbp, ubp, pg, reltol=1e-7, abstol=0.0, initdiv=3)
    assert np.all(st == 0) and np.all(np.abs(I - Ir) <= 1e-7 * np.linalg.norm(Ir, axis=0)) and np.all(ne >= 9 * 17)
    # budget exhausted: status 1 and the exact evaluation count
    I, E, ne, st = eng.hcubature_vec_batch(F["genz_gaussian"], 3, np.zeros(3), np.ones(3), _genz_vec_params(rng, 3, 3, 2, 6.0), reltol=1e-14,
                                           abstol=0.0, maxevals=1000)
    assert np.all(st == 1) and np.all(ne == 33 + 66 * 15)
    with pytest.raises(ib.IB200Error, match="ncomp"):
        eng.hcubature_vec_batch(F["genz_gaussian"], 2, np.zeros(2), np.ones(2), np.ones((4, 9, 2)))
    with pytest.raises(ib.IB200Error, match="nparam"):
        eng.hcubature_vec_batch(F["genz_gaussian"], 2, np.zeros(2), np.ones(2), np.ones((3, 2, 2)))
    with pytest.raises(ib.IB200Error, match="tolerance"):
        eng.hcubature_vec_batch(F["genz_gaussian"], 2, np.zeros(2), np.ones(2), np.ones((4, 2, 2)), reltol=-1.0)
