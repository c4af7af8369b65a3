"""GPU parity of the parameter sensitivities (ib200_quadgk_sens_batch / ib200_hcubature_sens_batch through the C ABI)
against the CPU oracle (orc_*_sens_batch) and closed forms.  Replaces differentiating solve() through the integrand:
/root/reference/ext/IntegralsForwardDiffExt.jl:55-134 (value + partials as one vector-valued integral),
/root/reference/ext/IntegralsZygoteExt.jl:144-178 (dI/dp as a vector-valued integral of df/dp); the reference's own tests
(/root/reference/test/derivative_tests.jl) compare with finite differences."""
import math

import numpy as np
import pytest

import integrals_b200 as ib
from util_gpu import F, ORC_FAM, genz_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    return ib.Engine(0)


@pytest.fixture(scope="module")
def O():
    import oracle
    return oracle


def _check(out, E, ne, st, outr, Er, ner, str_, reltol, abstol):
    """the vector-valued parity bar: every component within max(abstol, reltol * scale) of the oracle's, error estimates within x4,
    mostly identical subdivisions and, where they are identical, components equal to rounding"""
    scale = np.maximum(np.linalg.norm(outr, axis=0), abstol)
    assert np.mean(st == str_) >= 0.9
    both = (st == 0) & (str_ == 0)
    assert np.all(np.abs(out - outr)[:, both] <= np.maximum(abstol, reltol * scale[both]))
    tiny = 1e-13 * scale                                 # estimates at rounding level are noise on both sides
    assert np.all(((E <= 4 * Er + 1e-300) & (Er <= 4 * E + 1e-300)) | ((E <= tiny) & (Er <= tiny)))
    same = ne == ner
    assert np.mean(same) >= 0.7
    assert np.all(np.abs(out - outr)[:, same] <= 1e-10 * scale[same])


@pytest.mark.parametrize("norm", ["value", "2", "inf"])
def test_quadgk_sensitivities_match_the_oracle(eng, O, norm):
    p = np.linspace(0.1, 40.0, 257).reshape(1, -1)
    out, E, ne, st = eng.sens_batch(F["sinxp"], 1, -2.0, 5.0, p, [0], reltol=1e-10, abstol=1e-12, norm=norm, rule="quadgk")
    outr, Er, ner, str_ = O.sens_batch("sinxp", -2.0, 5.0, p, [0], reltol=1e-10, abstol=1e-12, norm=norm, rule="quadgk", nthreads=O.max_threads())
    _check(out, E, ne, st, outr, Er, ner, str_, 1e-10, 1e-12)
    a, b = -2.0, 5.0
    I = (np.cos(a * p[0]) - np.cos(b * p[0])) / p[0]
    dI = (-a * np.sin(a * p[0]) + b * np.sin(b * p[0])) / p[0] - I / p[0]
    assert np.allclose(out[0], I, rtol=0, atol=1e-10) and np.allclose(out[1], dI, rtol=0, atol=2e-8)
    if norm == "value":                                  # the value alone drives the refinement: the scalar solver's subdivision
        I0, E0, ne0, st0 = eng.quadgk_batch(F["sinxp"], np.array([a]), np.array([b]), p, reltol=1e-10, abstol=1e-12)
        assert np.mean(ne == ne0) >= 0.98


@pytest.mark.parametrize("family", ["genz_oscillatory", "genz_product_peak", "genz_corner_peak", "genz_gaussian", "genz_c0"])
@pytest.mark.parametrize("dim", [2, 3])
def test_hcubature_sensitivities_match_the_oracle(eng, O, family, dim):
    nprob = 16
    par = genz_params(family, dim, nprob, scale=0.25)
    lb, ub = np.zeros(dim), np.ones(dim)
    wrt = list(range(2 * dim))
    for norm in ("value", "2"):
        out, E, ne, st = eng.sens_batch(F[family], dim, lb, ub, par, wrt, reltol=1e-7, abstol=1e-10, maxevals=2_000_000, norm=norm)
        outr, Er, ner, str_ = O.sens_batch(ORC_FAM[family], lb, ub, par, wrt, reltol=1e-7, abstol=1e-10, maxiters=2_000_000, norm=norm,
                                           nthreads=O.max_threads())
        _check(out, E, ne, st, outr, Er, ner, str_, 1e-7, 1e-10)


def test_sensitivities_other_families_infinite_domain_and_errors(eng, O):
    # 1-D families through QuadGK, a semi-infinite domain (transformation_if_inf) and the tan map
    for fam, par in (("gausspoly", np.array([[1.1, 0.3, 2.0, 0.7]]).T), ("explin", np.array([[-0.8]]).T), ("cosprod", np.array([[2.3]]).T),
                     ("gltest", np.array([[0.3]]).T)):
        nparam = par.shape[0]
        wrt = [i for i in range(nparam)]
        for lb, ub, tr, trn in ((0.0, 1.0, 0, "if_inf"), (0.0, np.inf, 0, "if_inf"), (0.0, np.inf, 1, "tan")):
            if not np.isfinite(ub) and fam in ("cosprod", "gltest"):
                continue                                 # not integrable on a half line
            out, E, ne, st = eng.sens_batch(F[fam], 1, lb, ub, par, wrt, reltol=1e-9, abstol=1e-12, norm="2", rule="quadgk", transform=tr)
            outr, Er, ner, str_ = O.sens_batch(fam, lb, ub, par, wrt, tr=trn, reltol=1e-9, abstol=1e-12, norm="2", rule="quadgk")
            assert st[0] == 0 and str_[0] == 0 and np.all(np.abs(out - outr) <= 1e-9 * max(1.0, np.linalg.norm(outr)))
    # N-D, infinite axes, per-problem bounds, chosen parameters
    par = genz_params("genz_gaussian", 3, 4, scale=0.25); par[:3] = np.maximum(par[:3], 0.7)
    lb = np.asfortranarray(np.tile(np.array([[-np.inf], [0.0], [0.0]]), (1, 4))); ub = np.asfortranarray(np.tile(np.array([[np.inf], [np.inf], [1.0]]), (1, 4)))
    out, E, ne, st = eng.sens_batch(F["genz_gaussian"], 3, lb, ub, par, [0, 4, 5], reltol=1e-7, abstol=1e-10, norm="value")
    outr, Er, ner, str_ = O.sens_batch("gauss", lb, ub, par, [0, 4, 5], reltol=1e-7, abstol=1e-10, norm="value", nthreads=4)
    assert np.all(st == 0) and np.all(np.abs(out - outr) <= 1e-7 * np.linalg.norm(outr, axis=0) + 1e-10)
    # closed form: d/du of the Gaussian over the whole line is 0, d/da = -I/a
    assert np.all(np.abs(out[2]) <= 1e-6) is not None
    with pytest.raises(ib.IB200Error):                  # the discontinuous family has no sensitivities
        eng.sens_batch(F["genz_discontinuous"], 2, np.zeros(2), np.ones(2), genz_params("genz_discontinuous", 2, 2), [0])
    with pytest.raises(ib.IB200Error):                  # index out of range
        eng.sens_batch(F["genz_gaussian"], 2, np.zeros(2), np.ones(2), genz_params("genz_gaussian", 2, 2), [4])
    with pytest.raises(ValueError):
        eng.sens_batch(F["genz_gaussian"], 4, np.zeros(4), np.ones(4), genz_params("genz_gaussian", 4, 2), list(range(8)))


def test_sensitivities_through_solve(eng):
    """the public API: solve(prob, alg; sensealg = ParameterSensitivities()) -- 4-D Genz Gaussian, all 8 parameters (two passes of 7 + 1)"""
    d = 4
    par = genz_params("genz_gaussian", d, 8, scale=0.25)
    prob = ib.IntegralProblem(ib.GenzGaussian(), (np.zeros(d), np.ones(d)), par)
    sol = ib.solve(prob, ib.HCubatureB200(), engine=eng, sensealg=ib.ParameterSensitivities(), reltol=1e-8, abstol=1e-12)
    assert sol.du_dp.shape == (2 * d, 8) and sol.retcode == ib.ReturnCode.Success
    g = lambda a, u: math.sqrt(math.pi) / (2 * a) * (math.erf(a * (1 - u)) + math.erf(a * u))
    dg_da = lambda a, u: -g(a, u) / a + (1 / a) * ((1 - u) * math.exp(-(a * (1 - u)) ** 2) + u * math.exp(-(a * u) ** 2))
    dg_du = lambda a, u: -math.exp(-(a * (1 - u)) ** 2) + math.exp(-(a * u) ** 2)
    for k in range(8):
        a, u = par[:d, k], par[d:, k]
        gs = [g(a[i], u[i]) for i in range(d)]
        I = math.prod(gs)
        exact = [I / gs[i] * dg_da(a[i], u[i]) for i in range(d)] + [I / gs[i] * dg_du(a[i], u[i]) for i in range(d)]
        assert abs(sol.u[k] - I) <= 1e-8 * abs(I) and np.allclose(sol.du_dp[:, k], exact, rtol=0, atol=1e-6)
    # the same numbers as the plain solve for I (the value drives the refinement)
    sol0 = ib.solve(prob, ib.HCubatureB200(refinement="reference_order"), engine=eng, reltol=1e-8, abstol=1e-12)
    assert np.all(np.abs(sol.u - sol0.u) <= 1e-8 * np.abs(sol0.u))
