"""Parameter sensitivities, CPU side: the gradients of the registered families written out in oracle/oracle.c
(orc_family_grad; the device copy is csrc/family.cuh Grad) against central finite differences of orc_family_eval, and dI/dp computed as a
vector-valued integral of [f, df/dp] (orc_*_sens_batch, the checker of ib200_*_sens_batch) against closed forms and against finite
differences of the integral.  The reference obtains the same numbers by differentiating solve() through the integrand
(/root/reference/ext/IntegralsForwardDiffExt.jl:55-134, /root/reference/ext/IntegralsZygoteExt.jl:144-178; its tests:
/root/reference/test/derivative_tests.jl compares against finite differences at rtol ~ 1e-3..1e-6 in the same way)."""
import math

import numpy as np
import pytest

import oracle as O

FAMS = {"sinxp": 1, "osc": 3, "prodpeak": 3, "corner": 3, "gauss": 3, "c0": 3, "cosprod": 3, "gausspoly": 2, "explin": 3, "gltest": 1}


def _params(fam, d, rng):
    if fam == "sinxp":
        return np.array([1.7])
    if fam == "gltest":
        return np.array([0.3])
    if fam in ("cosprod", "explin"):
        return rng.uniform(0.3, 1.5, d)
    if fam == "gausspoly":
        return np.concatenate([rng.uniform(0.5, 1.5, d), rng.uniform(0.2, 0.8, d), np.array([2.0, 1.0])[:d], [0.7]])
    return np.concatenate([rng.uniform(0.5, 2.0, d), rng.uniform(0.2, 0.8, d)])


@pytest.mark.parametrize("fam", list(FAMS))
def test_family_gradients_against_finite_differences(fam):
    d = FAMS[fam]
    rng = np.random.default_rng(7)
    par = _params(fam, d, rng)
    for _ in range(5):
        x = rng.uniform(0.05, 0.95, d)
        g = O.family_grad(fam, x, par)
        assert g is not None and abs(g[0] - O.family_eval(fam, x, par)) <= 4e-16 * abs(g[0]) + 1e-300
        for i in range(par.size):
            if fam == "gausspoly" and 2 * d <= i < 3 * d:
                assert g[1 + i] == 0.0                     # integer exponents
                continue
            h = 1e-6 * max(1.0, abs(par[i]))
            pp, pm = par.copy(), par.copy(); pp[i] += h; pm[i] -= h
            fd = (O.family_eval(fam, x, pp) - O.family_eval(fam, x, pm)) / (2 * h)
            assert abs(g[1 + i] - fd) <= 1e-7 * max(1.0, abs(fd)), (fam, i, g[1 + i], fd)


def test_discontinuous_family_has_no_sensitivities():
    assert O.family_grad("discont", np.array([0.3, 0.4]), np.array([1.0, 2.0, 0.5, 0.5])) is None


def test_quadgk_sensitivity_closed_form():
    """d/dp int_{-2}^{5} sin(x p) dx = [cos(a p) - cos(b p)]' / p ...: the README integrand (test/interface_tests.jl uses the same family)"""
    p = np.array([[0.7, 1.7, 3.1]])
    a, b = -2.0, 5.0
    out, E, ne, st = O.sens_batch("sinxp", a, b, p, [0], reltol=1e-10, abstol=1e-12, rule="quadgk")
    I = (np.cos(a * p[0]) - np.cos(b * p[0])) / p[0]
    dI = (-a * np.sin(a * p[0]) + b * np.sin(b * p[0])) / p[0] - I / p[0]
    assert np.all(st == 0) and np.allclose(out[0], I, rtol=0, atol=1e-11) and np.allclose(out[1], dI, rtol=0, atol=1e-9)
    # the value alone drives the refinement: the same subdivision (evaluations) as the scalar solve
    I0, E0, ne0, st0 = O.quadgk_batch("sinxp", a, b, p, reltol=1e-10, abstol=1e-12)
    assert np.array_equal(ne, ne0) and np.allclose(out[0], I0, rtol=1e-15, atol=0)


@pytest.mark.parametrize("norm", ["value", "2"])
def test_hcubature_sensitivity_gaussian_closed_form(norm):
    """3-D Genz Gaussian on [0,1]^3: I = prod_i g(a_i, u_i), g = sqrt(pi)/(2a) [erf(a(1-u)) + erf(a u)]"""
    d = 3
    par = np.array([[1.3, 0.8, 2.1, 0.3, 0.6, 0.45]]).T
    a, u = par[:d, 0], par[d:, 0]
    g = lambda a, u: math.sqrt(math.pi) / (2 * a) * (math.erf(a * (1 - u)) + math.erf(a * u))
    dg_da = lambda a, u: -g(a, u) / a + (1 / a) * ((1 - u) * math.exp(-(a * (1 - u)) ** 2) + u * math.exp(-(a * u) ** 2))
    dg_du = lambda a, u: -math.exp(-(a * (1 - u)) ** 2) + math.exp(-(a * u) ** 2)
    gs = [g(a[i], u[i]) for i in range(d)]
    I = math.prod(gs)
    exact = [I / gs[i] * dg_da(a[i], u[i]) for i in range(d)] + [I / gs[i] * dg_du(a[i], u[i]) for i in range(d)]
    out, E, ne, st = O.sens_batch("gauss", np.zeros(d), np.ones(d), par, list(range(2 * d)), reltol=1e-8, abstol=1e-10, norm=norm)
    assert st[0] == 0 and abs(out[0, 0] - I) <= 1e-8 * abs(I)
    assert np.allclose(out[1:, 0], exact, rtol=0, atol=2e-7)
    if norm == "value":                                  # identical subdivision to the scalar solve
        I0, E0, ne0, st0 = O.hcubature_batch("gauss", np.zeros(d), np.ones(d), par, reltol=1e-8, abstol=1e-10)
        assert ne[0] == ne0[0] and abs(out[0, 0] - I0[0]) <= 1e-15 * abs(I0[0])
    else:                                                # the derivatives take part in the error control: at least as much work
        outv, Ev, nev, stv = O.sens_batch("gauss", np.zeros(d), np.ones(d), par, list(range(2 * d)), reltol=1e-8, abstol=1e-10, norm="value")
        assert ne[0] >= nev[0]


def test_sensitivity_against_finite_differences_of_the_integral_infinite_domain():
    """dI/dp on a semi-infinite domain (transformation_if_inf): product peak in 2-D, central differences of the adaptive solve"""
    par = np.array([[1.2, 0.9, 0.3, 0.5]]).T
    lb, ub = np.array([0.0, -np.inf]), np.array([np.inf, 1.0])
    out, E, ne, st = O.sens_batch("prodpeak", lb, ub, par, [0, 1, 2, 3], reltol=1e-9, abstol=1e-12, norm="2")
    assert st[0] == 0
    for i in range(4):
        h = 1e-5
        pp, pm = par.copy(), par.copy(); pp[i] += h; pm[i] -= h
        Ip, *_ = O.hcubature_batch("prodpeak", lb, ub, pp, reltol=1e-11, abstol=1e-13)
        Im, *_ = O.hcubature_batch("prodpeak", lb, ub, pm, reltol=1e-11, abstol=1e-13)
        fd = (Ip[0] - Im[0]) / (2 * h)
        assert abs(out[1 + i, 0] - fd) <= 1e-5 * max(1.0, abs(fd)), (i, out[1 + i, 0], fd)
