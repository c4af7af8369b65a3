"""Pins the CPU oracle (oracle/oracle.c) against every known answer the reference's own tests
hold for the hot path: the reference has no golden files, every
expected value is a closed form written inline in test/*.jl -- reproduced here with the file:line.
Also checks the degree-of-exactness identities of the restated QuadGK / HCubature rules.
"""
import math

import numpy as np
import pytest

import oracle as O

INF = math.inf


# ---------------------------------------------------------------- README.md:30-40
def test_readme_example_quadgk():
    I, E, ne, st = O.solve_quadgk(lambda x: math.sin(1.7 * x), -2.0, 5.0, reltol=1e-10)
    exact = (math.cos(-3.4) - math.cos(8.5)) / 1.7
    assert st == 0
    assert I == pytest.approx(exact, rel=1e-14)
    # restated QuadGK: 3 segments, 75 evaluations, E = 5.69e-9
    assert ne == 75 and E == pytest.approx(5.688e-9, rel=1e-3)


# ---------------------------------------------------------------- test/sampled_tests.jl:3-30
@pytest.mark.parametrize("rule", ["trapezoid", "simpson"])
@pytest.mark.parametrize("gridkind", ["range", "rand_even", "rand_odd"])
def test_sampled_closed_forms(rule, gridkind):
    lb, ub, npoints = 0.4, 1.1, 1000
    rng = np.random.default_rng(7)
    if gridkind == "range":
        x = None
        grid = np.linspace(lb, ub, npoints)
        h = (ub - lb) / (npoints - 1)
    else:
        k = npoints if gridkind == "rand_even" else npoints + 1
        grid = np.concatenate([[lb], np.sort(rng.random(k) * (ub - lb) + lb), [ub]])
        x, h = grid, None
    exact = [(ub ** 6 - lb ** 6) / 6, math.sin(ub) - math.sin(lb)]
    for f, ex in zip([lambda v: v ** 5, np.cos], exact):
        y = f(grid)
        err = O.sampled_evalrule(rule, y, x=x, h=h) - ex
        assert abs(err) < 1e-4           # reference asserts the one-sided `error .< 10^-4`
        y2 = np.asfortranarray(np.stack([y, y], axis=0))   # f.([grid grid]') : 2 x n, dim = 2
        out = O.sampled_evalrule(rule, y2, x=x, h=h, dim=1)
        assert out.shape == (2,) and np.all(np.abs(out - ex) < 1e-4)


def test_sampled_ones_exact():  # test/sampled_tests.jl:35-43 (units dropped)
    out = O.sampled_evalrule("trapezoid", np.ones((2, 3), order="F"), x=np.array([0.0, 1.0, 2.0]), dim=1)
    assert np.array_equal(out, [2.0, 2.0])


def test_sampled_dim1_vs_dim2():  # test/sampled_tests.jl:69-82
    x = np.arange(0.0, 1.0001, 0.1)
    y = np.asfortranarray(np.sin(x)[:, None] * np.cos(x)[None, :])
    a = O.sampled_evalrule("trapezoid", y, h=0.1, dim=0)
    b = O.sampled_evalrule("trapezoid", np.asfortranarray(y.T), h=0.1, dim=1)
    assert np.allclose(a, b, rtol=0, atol=1e-15)


def test_sampled_weights_exactness():
    # Simpson-uniform (simpsons.jl:16-25) integrates cubics exactly; trapezoid integrates lines exactly
    n, h = 23, 0.37
    xs = np.arange(n) * h
    w = O.sampled_weights("simpson", n, h=h)
    for k in range(4):
        assert w @ xs ** k == pytest.approx(xs[-1] ** (k + 1) / (k + 1), rel=1e-13)
    # non-uniform Simpson (simpsons.jl:40-82) integrates quadratics exactly, odd and even lengths
    rng = np.random.default_rng(3)
    for n in (3, 4, 5, 6, 7, 8, 101, 102):
        x = np.sort(rng.random(n)) * 2.0 + 0.5
        w = O.sampled_weights("simpson", n, x=x)
        for k in range(3):
            assert w @ x ** k == pytest.approx((x[-1] ** (k + 1) - x[0] ** (k + 1)) / (k + 1), rel=1e-11)
        w = O.sampled_weights("trapezoid", n, x=x)
        for k in range(2):
            assert w @ x ** k == pytest.approx((x[-1] ** (k + 1) - x[0] ** (k + 1)) / (k + 1), rel=1e-12)
    # first-match-wins for short uniform Simpson grids (simpsons.jl:20-24)
    assert np.allclose(O.sampled_weights("simpson", 3, h=48.0), [17, 59, 17])
    assert np.allclose(O.sampled_weights("simpson", 5, h=48.0), [17, 59, 43, 59, 17])


def test_sampled_empty_raises():  # sampled.jl:55,72,106
    with pytest.raises(ValueError, match="No points"):
        O.sampled_evalrule("trapezoid", np.zeros((3, 0), order="F"), h=1.0, dim=1)


# ---------------------------------------------------------------- test/quadrule_tests.jl
def _trapz(n):  # quadrule_tests.jl:17-25
    x = np.linspace(-1, 1, n)
    h = x[1] - x[0]
    w = np.full(n, h); w[0] = w[-1] = h / 2
    return x, w


def _exact_cosprod(lb, ub, p):  # quadrule_tests.jl:7
    return math.prod((math.sin(p * u) - math.sin(p * l)) / p for l, u in zip(lb, ub))


def test_quadrule_1d_trapz():  # quadrule_tests.jl:27-36
    x, w = _trapz(1000)
    u = O.solve_quadrule(lambda v: math.cos(2.0 * v), -1.2, 3.5, x, w)
    assert u == pytest.approx(_exact_cosprod([-1.2], [3.5], 2.0), rel=1e-3)


def test_quadrule_2d_tensor():  # quadrule_tests.jl:41-55
    x, w = _trapz(100)
    lb, ub, p = [-1.2, -1.0], [3.5, 3.7], 1.2
    f = lambda v: math.cos(p * v[0]) * math.cos(p * v[1])
    u = O.solve_quadrule_tensor(f, lb, ub, x, w)
    assert u == pytest.approx(_exact_cosprod(lb, ub, p), rel=1e-3)
    # explicit node list in Iterators.product order (first index fastest) gives the same sum
    nodes = np.array([[x[i], x[j]] for j in range(100) for i in range(100)])
    wts = np.array([w[i] * w[j] for j in range(100) for i in range(100)])
    u2 = O.solve_quadrule(f, lb, ub, nodes, wts)
    assert u2 == pytest.approx(u, rel=1e-15)


def test_quadrule_lorentzian_inf(gl):  # quadrule_tests.jl:59-69
    x, w = gl(1000)
    assert O.solve_quadrule(lambda v: 1.0 / (v * v + 1.0), -INF, INF, x, w) == pytest.approx(math.pi, rel=1e-4)
    assert O.solve_quadrule(lambda v: 1.3 / (v * v + 1.69), -INF, INF, x, w) == pytest.approx(math.pi, rel=1e-4)


# ---------------------------------------------------------------- test/gaussian_quadrature_tests.jl
def test_gauss_legendre_known_values(gl):
    x, w = gl(250)                                      # GaussLegendre() default n = 250
    f = lambda v: 5 * v + math.sin(v) - 3.3 * math.exp(v)
    exact = -math.exp(3) * 3.3 + 3.3 / math.exp(5) - 40 + math.cos(5) - math.cos(3)   # :57
    assert O.solve_gauss_legendre(f, -5.0, 3.0, x, w) == pytest.approx(exact, rel=1.5e-8)
    assert O.solve_gauss_legendre(f, -5.0, 3.0, x, w, subintervals=7) == pytest.approx(exact, rel=1.5e-8)  # :60
    g = lambda v: math.exp(-v * v)
    for sub in (1, 17):                                 # :62-88
        assert O.solve_gauss_legendre(g, 0.0, INF, x, w, subintervals=sub) == pytest.approx(math.sqrt(math.pi) / 2, rel=1.5e-8)
        assert O.solve_gauss_legendre(g, -INF, INF, x, w, subintervals=sub) == pytest.approx(math.sqrt(math.pi), rel=1.5e-8)
        assert O.solve_gauss_legendre(g, -INF, 0.0, x, w, subintervals=sub) == pytest.approx(math.sqrt(math.pi) / 2, rel=1.5e-8)
    p = [0.3, 1.3, -0.5]                                # :91-99
    h = lambda v: 1 + v + v ** p[0] - math.cos(v * p[1]) + math.exp(v) * p[2]
    assert O.solve_gauss_legendre(h, 2.0, 6.3, x, w) == pytest.approx(-240.25235266303063249920743158729, rel=1.5e-8)
    x5, w5 = gl(500)
    assert O.solve_gauss_legendre(h, 2.0, 6.3, x5, w5, subintervals=17) == pytest.approx(-240.25235266303063249920743158729, rel=1.5e-8)


def test_gauss_legendre_commented_block(gl):  # gaussian_quadrature_tests.jl:4-20 (kept as documentation there)
    x, w = gl(250)
    f = lambda v: v ** 3 * math.sin(5 * v)
    exact = 2 / 625 * (69 * math.sin(5) - 95 * math.cos(5))
    assert O.solve_gauss_legendre(f, -1.0, 1.0, x, w) == pytest.approx(exact, rel=1e-12)
    assert O.solve_gauss_legendre(f, -1.0, 1.0, x, w, subintervals=2) == pytest.approx(exact, rel=1e-12)


# ---------------------------------------------------------------- test/inf_integral_tests.jl:42-163
_mvn = lambda x: math.exp(-(x[0] ** 2 + x[1] ** 2) / 0.8) / (2 * math.pi * 0.4)
_npdf = lambda x: math.exp(-x * x / 2) / math.sqrt(2 * math.pi)
_npdf_q = lambda x: _npdf(x[0]) * x[1] ** 2
_lor = lambda x: 1 / (x * x + 1)
_lor2 = lambda x: 1 / ((x - 2) ** 2 + 1)
INF_PROBLEMS = [
    (_mvn, [-INF, -INF], [INF, INF], 1.0), (_mvn, [INF, INF], [-INF, -INF], 1.0),
    (_mvn, [-INF, 0], [INF, INF], 0.5), (_mvn, [-INF, -INF], [INF, 0], 0.5),
    (_npdf_q, [-INF, 1], [INF, 4], 21.0), (_npdf_q, [0.0, 1], [INF, 4], 10.5), (_npdf_q, [-INF, 1], [0.0, 4], 10.5),
    (_npdf, -INF, INF, 1.0), (_npdf, INF, -INF, -1.0), (_npdf, 0.0, INF, 0.5), (_npdf, 0.0, -INF, -0.5),
    (_npdf, -INF, 0.0, 0.5), (_npdf, INF, 0.0, -0.5),
    (_lor, -INF, INF, math.pi), (_lor2, -INF, 2, math.pi / 2), (_lor2, 2, INF, math.pi / 2),
    (_lor2, INF, 2, -math.pi / 2), (_lor2, 2, -INF, -math.pi / 2),
    (lambda x: x * x, 1, 4, 21.0), (lambda x: x * x, 4, 1, -21.0),
]


@pytest.mark.parametrize("k", range(len(INF_PROBLEMS)))
def test_inf_integrals(k, gl):
    f, lb, ub, sol = INF_PROBLEMS[k]
    tol = 1e-3                                          # reltol = 0, abstol = 1e-3 (:3-4,182)
    if isinstance(lb, list):
        I, E, ne, st = O.solve_hcubature(f, lb, ub, reltol=0.0, abstol=tol)
        assert abs(I - sol) < tol
    else:
        lb, ub = float(lb), float(ub)
        I, *_ = O.solve_quadgk(f, lb, ub, reltol=0.0, abstol=tol)
        assert abs(I - sol) < tol
        I, *_ = O.solve_hcubature(f, [lb], [ub], reltol=0.0, abstol=tol)
        assert abs(I - sol) < tol
        x, w = gl(50)                                   # QuadratureRule(gausslegendre, n = 50)
        assert abs(O.solve_quadrule(f, lb, ub, x, w) - sol) < tol


# ---------------------------------------------------------------- test/alt_transformation_tests.jl
@pytest.mark.parametrize("tr", ["tan", "cot"])
def test_alt_transformations(tr):
    kw = dict(tr=tr, reltol=1e-6, abstol=1e-6)
    s = lambda *a, **k: O.solve_quadgk(*a, **k)[0]
    assert s(lambda x: math.exp(-x * x), -INF, INF, **kw) == pytest.approx(math.sqrt(math.pi), rel=1e-5)
    assert s(lambda x: 1 / (1 + x * x), -INF, INF, **kw) == pytest.approx(math.pi, rel=1e-5)
    if tr == "tan":
        assert s(lambda x: math.exp(-x), 0.0, INF, **kw) == pytest.approx(1.0, rel=1e-4)
        assert s(lambda x: math.exp(x), -INF, 0.0, **kw) == pytest.approx(1.0, rel=1e-4)
        assert s(lambda x: 1 / ((x - 2) ** 2 + 1), 2.0, INF, **kw) == pytest.approx(math.pi / 2, rel=1e-4)
        assert s(lambda x: math.exp(x), -INF, -1.0, **kw) == pytest.approx(math.exp(-1), rel=1e-4)
    else:
        kw2 = dict(tr=tr, reltol=1e-4, abstol=1e-4)
        assert s(lambda x: math.exp(-x), 0.0, INF, **kw2) == pytest.approx(1.0, rel=1e-3)
        assert s(lambda x: math.exp(x), -INF, 0.0, **kw2) == pytest.approx(1.0, rel=1e-3)
        assert s(lambda x: math.exp(-x * x), 0.0, INF, **kw2) == pytest.approx(math.sqrt(math.pi) / 2, rel=1e-3)
    # finite domains unchanged (:72-90); flipped limits (:92-101); HCubature 1-D (:103-111)
    assert s(lambda x: x * x, 0.0, 1.0, tr=tr, reltol=1e-10, abstol=1e-10) == pytest.approx(1 / 3, rel=1e-8)
    assert s(lambda x: math.exp(-x * x), INF, -INF, **kw) == pytest.approx(-math.sqrt(math.pi), rel=1e-5)
    I, *_ = O.solve_hcubature(lambda x: math.exp(-x * x), [-INF], [INF], tr=tr, reltol=1e-6, abstol=1e-6)
    assert I == pytest.approx(math.sqrt(math.pi), rel=1e-4)


# ---------------------------------------------------------------- test/interface_tests.jl:129-187
@pytest.mark.parametrize("dim", [1, 2])
def test_interface_integrands(dim):
    lb, ub = [1.0] * dim, [3.0] * dim
    exact1 = 2.0 ** dim
    exactc = (math.sin(3) - math.sin(1)) ** dim
    one = lambda x: 1.0
    cosp = (lambda x: math.cos(x)) if dim == 1 else (lambda x: math.cos(x[0]) * math.cos(x[1]))
    I, *_ = O.solve_hcubature(one, lb, ub, reltol=1e-5, abstol=1e-5)
    assert I == pytest.approx(exact1, rel=1e-2)
    I, *_ = O.solve_hcubature(cosp, lb, ub, reltol=1e-5, abstol=1e-5)
    assert I == pytest.approx(exactc, rel=1e-2)
    if dim == 1:
        assert O.solve_quadgk(one, 1.0, 3.0, reltol=1e-5, abstol=1e-5)[0] == pytest.approx(exact1, rel=1e-2)
        assert O.solve_quadgk(cosp, 1.0, 3.0, reltol=1e-5, abstol=1e-5)[0] == pytest.approx(exactc, rel=1e-2)


# ---------------------------------------------------------------- rule identities
def test_gk15_degree_of_exactness():
    for k in range(0, 23):      # Kronrod-15 exact through degree 22
        I, E = O.gk15(lambda x, k=k: x ** k, -0.3, 1.7)
        exact = (1.7 ** (k + 1) - (-0.3) ** (k + 1)) / (k + 1)
        assert I == pytest.approx(exact, rel=2e-13)
        if k <= 13:             # embedded Gauss-7 exact through degree 13 -> E ~ 0
            assert E <= 1e-12 * max(1.0, abs(exact))
    I, E = O.gk15(lambda x: x ** 14, -0.3, 1.7)
    assert E > 1e-9             # Gauss-7 is NOT exact at degree 14


def test_genz_malik_degree_and_counts():
    for n, npts in [(2, 17), (3, 33), (4, 57), (8, 401)]:
        cnt = [0]

        def one(x, cnt=cnt):
            cnt[0] += 1
            return 1.0
        I, E, k = O.cubrule(one, [0.0] * n, [1.0] * n)
        assert cnt[0] == npts and I == pytest.approx(1.0, rel=1e-14) and E < 1e-13
    a, b = [0.2, -1.0, 0.5], [1.3, 0.4, 2.0]
    rng = np.random.default_rng(5)
    for _ in range(12):
        e = rng.multinomial(7, [1 / 3] * 3)   # total degree 7 monomial
        f = lambda x, e=e: x[0] ** e[0] * x[1] ** e[1] * x[2] ** e[2]
        exact = math.prod((b[i] ** (e[i] + 1) - a[i] ** (e[i] + 1)) / (e[i] + 1) for i in range(3))
        I, E, k = O.cubrule(f, a, b)
        assert I == pytest.approx(exact, rel=1e-12, abs=1e-13)
    # degree-5 monomials: embedded rule exact too -> E ~ 0
    I, E, k = O.cubrule(lambda x: x[0] ** 3 * x[1] ** 2, a, b)
    assert E < 1e-12
    # split axis = largest fourth difference
    I, E, k = O.cubrule(lambda x: math.cos(7 * x[1]) + x[0] + x[2], a, b)
    assert k == 1


def test_hcubature_converges_to_closed_form():
    a = np.array([5.0, 3.0, 4.0]); u = np.array([0.3, 0.6, 0.5])
    par = np.concatenate([a, u])
    exact = math.prod(math.sqrt(math.pi) / (2 * a[i]) * (math.erf(a[i] * (1 - u[i])) + math.erf(a[i] * u[i])) for i in range(3))
    I, E, ne, st = O.hcubature_batch("gauss", [0.0] * 3, [1.0] * 3, par, reltol=1e-7, abstol=0.0)
    assert st[0] == 0 and abs(I[0] - exact) <= 1e-7 * exact and E[0] <= 1e-7 * abs(I[0])


def test_issue242_abstol_is_an_absolute_stop():  # interface_tests.jl:446-454 semantics: abstol honoured
    # |I| ~ 0.03 for p = 64: with the default abstol = 1e-8 the absolute test, not reltol, stops QuadGK
    I1, E1, n1, _ = O.solve_quadgk(lambda x: math.sin(64 * x), -2.0, 5.0, reltol=1e-10, abstol=1e-8)
    I2, E2, n2, _ = O.solve_quadgk(lambda x: math.sin(64 * x), -2.0, 5.0, reltol=1e-10, abstol=0.0)
    assert E1 <= 1e-8 and n2 > n1 and E2 <= 1e-10 * abs(I2)


def test_sampled_float32_restatement():   # test/interface_tests.jl:524-539 semantics: all arithmetic in Float32
    rng = np.random.default_rng(3)
    for n in (3, 11, 12, 1000):
        y = (rng.random(n) + 0.5).astype(np.float32)
        x = (np.sort(rng.random(n)) * 2 + 0.1).astype(np.float32)
        for rule in ("trapezoid", "simpson"):
            for kw in (dict(h=np.float32(0.01)), dict(x=x)):
                got, w = O.sampled_evalrule_f32(rule, y, **kw)
                assert isinstance(got, np.float32) and w.dtype == np.float32
                ref = O.sampled_evalrule(rule, y.astype(np.float64), **{k: (float(v) if k == "h" else v.astype(np.float64)) for k, v in kw.items()})
                w64 = O.sampled_weights(rule, n, **{k: (float(v) if k == "h" else v.astype(np.float64)) for k, v in kw.items()})
                # float weights/products/sequential sum vs the float64 rule on the same (float32-representable) inputs
                cond = 30 if (rule == "simpson" and "x" in kw) else 4        # non-uniform Simpson weights cancel in float
                assert abs(float(got) - ref) <= cond * n * 6e-8 * float(np.sum(np.abs(w64 * y)))


def test_quadgk_vector_valued_restatement():
    """orc_quadgk_vec (quadgk! with a norm, src/Integrals.jl:316-326): one component reproduces the scalar solver exactly; several
    components share the subdivision and each meets the closed form within the tolerance set by the norm of the whole vector"""
    p1 = np.array([[20.0]]).reshape(1, 1, 1)
    Iv, Ev, nev, stv = O.quadgk_vec_batch("sinxp", -2.0, 5.0, p1, reltol=1e-10, abstol=1e-12)
    Is, Es, nes, sts = O.quadgk_batch("sinxp", -2.0, 5.0, np.array([[20.0]]), reltol=1e-10, abstol=1e-12)
    assert Iv[0, 0] == Is[0] and Ev[0] == Es[0] and nev[0] == nes[0] and stv[0] == sts[0] == 0
    ps = np.array([1.7, 3.3, 0.5, 20.0])
    for norm in ("2", "inf"):
        I, E, ne, st = O.quadgk_vec_batch("sinxp", -2.0, 5.0, ps.reshape(1, 4, 1), reltol=1e-10, abstol=0.0, norm=norm)
        exact = (np.cos(-2 * ps) - np.cos(5 * ps)) / ps
        nI = np.linalg.norm(exact) if norm == "2" else np.abs(exact).max()
        assert st[0] == 0 and E[0] <= 1e-10 * nI * (1 + 1e-12) and np.all(np.abs(I[:, 0] - exact) <= 1e-10 * nI)
        assert ne[0] < nes[0]            # the easy components ride on the subdivision the hard one needs -- but the norm test stops earlier
    # Lorentzians over the real line through the automatic change of variables (test/inf_integral_tests.jl:85-150 integrands)
    par = np.array([[1.0, 0.0], [2.0, 0.5]]).T.reshape(2, 2, 1)
    I, E, ne, st = O.quadgk_vec_batch("prodpeak", -INF, INF, par, reltol=1e-9, abstol=0.0)
    assert st[0] == 0 and np.allclose(I[:, 0], [math.pi * 1.0, math.pi * 2.0], rtol=1e-8)


def test_kronrod_tables_and_rule_orders():
    """Gauss-Kronrod tables (integrals.jl_b200/kronrod.py, QuadGK.kronrod layout) and the oracle's rule of any order
    (QuadGKJL(order = n), src/algorithms.jl:51-56): order 7 equals the built-in QUADPACK constants bit for bit, the Kronrod
    rule integrates x^k exactly up to degree 3n+1, the embedded Gauss rule up to 2n-1, and every order meets the tolerance"""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from integrals_b200.kronrod import kronrod
    x, w, wg = kronrod(7)
    assert x[0] == -0.991455371120812639206854697526329 and w[7] == 0.209482141084727828012999174891714 and wg[3] == 0.417959183673469387755102040816327
    for n in (2, 3, 7, 10, 15, 21, 30):
        x, w, wg = kronrod(n)
        assert len(x) == n + 1 and len(w) == n + 1 and len(wg) == (n + 1) // 2 and x[-1] == 0.0 and np.all(np.diff(x) > 0)
        xs = np.concatenate([x, -x[:-1][::-1]]); ws = np.concatenate([w, w[:-1][::-1]])          # the full 2n+1 point rule
        for k in range(0, 3 * n + 2, 2):
            assert abs(np.sum(ws * xs ** k) - 2.0 / (k + 1)) < 5e-15
        odd = n % 2 == 1
        gx = np.concatenate([x[1::2], -x[1::2][::-1][(1 if odd else 0):]]); gw = np.concatenate([wg, wg[::-1][(1 if odd else 0):]])
        assert len(gx) == n
        for k in range(0, 2 * n, 2):
            assert abs(np.sum(gw * gx ** k) - 2.0 / (k + 1)) < 5e-15
    p = np.array([[1.7, 20.0, 50.0]])
    exact = (np.cos(-2 * p[0]) - np.cos(5 * p[0])) / p[0]
    I7, E7, ne7, st7 = O.quadgk_batch("sinxp", -2.0, 5.0, p, reltol=1e-10, abstol=1e-13)
    for n in (2, 3, 7, 10, 15, 21, 30):
        x, w, wg = kronrod(n)
        I, E, ne, st = O.quadgk_rule_batch("sinxp", -2.0, 5.0, p, x, w, wg, reltol=1e-10, abstol=1e-13)
        assert np.all(st == 0) and np.all(np.abs(I - exact) <= 1e-10 * np.abs(exact) + 1e-13) and np.all((ne - (2 * n + 1)) % (2 * (2 * n + 1)) == 0)
        if n == 7:
            assert np.array_equal(I, I7) and np.array_equal(E, E7) and np.array_equal(ne, ne7)


def test_vector_integrands_reference_known_answers():
    """test/interface_tests.jl:261-285 ("Standard Single Dimension Vector Integrands", QuadGKJL): f = collect(1.0:nout) on (1, 3)
    -> (ub - lb) * [1..nout]; through the registered family c * x^k * exp(-a^2 (x-u)^2) with a = 0, k = 0, c = 1..nout"""
    for nout in (1, 2):
        par = np.zeros((4, nout, 1), order="F")
        par[3, :, 0] = np.arange(1, nout + 1)
        I, E, ne, st = O.quadgk_vec_batch("gausspoly", 1.0, 3.0, par, reltol=1e-5, abstol=1e-5)
        assert st[0] == 0 and ne[0] == 15 and np.allclose(I[:, 0], 2.0 * np.arange(1, nout + 1), rtol=1e-12)
    # x^k components share one subdivision: exact for the (7,15) pair, one rule application
    par = np.zeros((4, 3, 1), order="F"); par[2, :, 0] = [0, 1, 2]; par[3, :, 0] = 1.0
    I, E, ne, st = O.quadgk_vec_batch("gausspoly", 1.0, 3.0, par, reltol=1e-10, abstol=0.0)
    assert np.allclose(I[:, 0], [2.0, 4.0, 26.0 / 3.0], rtol=1e-14) and ne[0] == 15


# ------------------------------------------------------------------------------------- vector-valued HCubature
def test_hcubature_vector_valued_restatement():
    """orc_hcubature_vec (HCubature with an array-valued integrand and `norm = alg.norm`, src/Integrals.jl:380-393;
    HCubatureJL.norm, src/algorithms.jl:66-67,108-115): one component reproduces the scalar solver bit for bit (both norms,
    Genz-Malik and the 1-D GK rule, initdiv > 1, infinite axes); several components share one subdivision and each meets its
    closed form within the tolerance that the norm of the whole vector sets"""
    rng = np.random.default_rng(7)
    for fam, d, kw in (("gauss", 3, {}), ("osc", 2, {}), ("prodpeak", 4, {}), ("c0", 2, {}), ("gauss", 1, {}), ("corner", 3, dict(initdiv=2)),
                       ("discont", 2, {})):
        par = np.asfortranarray(np.vstack([rng.uniform(1, 4, (d, 4)), rng.uniform(0.1, 0.9, (d, 4))]))
        I, E, ne, st = O.hcubature_batch(fam, np.zeros(d), np.ones(d), par, reltol=1e-7, abstol=0.0, maxiters=100000, **kw)
        for norm in ("2", "inf"):
            Iv, Ev, nev, stv = O.hcubature_vec_batch(fam, np.zeros(d), np.ones(d), par.reshape(2 * d, 1, 4, order="F"), reltol=1e-7, abstol=0.0,
                                                     maxiters=100000, norm=norm, **kw)
            assert np.array_equal(I, Iv[0]) and np.array_equal(E, Ev) and np.array_equal(ne, nev) and np.array_equal(st, stv)
    par = np.array([[0.8, 1.3], [0.2, -0.4]])                      # exp(-a^2 (x-u)^2) over R^2: pi / (a1 a2)
    I, E, ne, st = O.hcubature_batch("gauss", [-INF, -INF], [INF, INF], par.reshape(4, 1), reltol=1e-6, abstol=0.0)
    Iv, Ev, nev, stv = O.hcubature_vec_batch("gauss", [-INF, -INF], [INF, INF], par.reshape(4, 1, 1), reltol=1e-6, abstol=0.0)
    assert Iv[0, 0] == I[0] and Ev[0] == E[0] and nev[0] == ne[0] and abs(I[0] - math.pi / (0.8 * 1.3)) < 2e-6 * I[0]
    # three Gaussians of different width on [0,1]^3
    a = np.array([[1.0, 2.0, 3.0], [4.0, 2.0, 1.0], [6.0, 5.0, 4.0]]); u = np.array([[0.5, 0.4, 0.3], [0.2, 0.8, 0.5], [0.6, 0.5, 0.4]])
    par = np.asfortranarray(np.concatenate([a, u], axis=1).T.reshape(6, 3, 1))
    exact = np.array([math.prod(math.sqrt(math.pi) / (2 * ai) * (math.erf(ai * (1 - ui)) + math.erf(ai * ui)) for ai, ui in zip(a[c], u[c]))
                      for c in range(3)])
    for norm, nrm in (("2", np.linalg.norm(exact)), ("inf", np.abs(exact).max())):
        I, E, ne, st = O.hcubature_vec_batch("gauss", np.zeros(3), np.ones(3), par, reltol=1e-8, abstol=0.0, norm=norm)
        assert st[0] == 0 and E[0] <= 1e-8 * nrm * (1 + 1e-12) and np.all(np.abs(I[:, 0] - exact) <= 1e-8 * nrm)
        # the shared subdivision is at least as fine as what the hardest component needs alone under the same absolute target
        Is, Es, nes, _ = O.hcubature_batch("gauss", np.zeros(3), np.ones(3), par[:, 2, :], reltol=0.0, abstol=1e-8 * nrm)
        assert ne[0] >= nes[0]
    # maxevals and the evaluation count: 33 points per 3-D box, two boxes per refinement
    I, E, ne, st = O.hcubature_vec_batch("gauss", np.zeros(3), np.ones(3), par, reltol=1e-14, abstol=0.0, maxiters=1000)
    assert st[0] == 1 and ne[0] == 33 + 66 * 15


def test_hcubature_vector_integrands_reference_known_answers():
    """test/interface_tests.jl:287-312 ("Standard Vector Integrands", HCubatureJL, dim 1..2, nout 1..2, reltol = abstol = 1e-5):
    f = collect(1.0:nout) on [1,3]^dim -> prod(ub - lb) * [1..nout]; prod(cos.(x)) -> prod(sin(3) - sin(1))"""
    for dim in (1, 2):
        lb, ub = np.ones(dim), 3 * np.ones(dim)
        for nout in (1, 2):
            par = np.zeros((3 * dim + 1, nout, 1), order="F")
            par[3 * dim, :, 0] = np.arange(1, nout + 1)              # c * x^0 * exp(0)
            I, E, ne, st = O.hcubature_vec_batch("gausspoly", lb, ub, par, reltol=1e-5, abstol=1e-5)
            assert st[0] == 0 and ne[0] == (15 if dim == 1 else 17) and np.allclose(I[:, 0], 2.0 ** dim * np.arange(1, nout + 1), rtol=1e-13)
            I, E, ne, st = O.hcubature_vec_batch("cosprod", lb, ub, np.ones((dim, nout, 1)), reltol=1e-5, abstol=1e-5)
            assert st[0] == 0 and np.allclose(I[:, 0], (math.sin(3.0) - math.sin(1.0)) ** dim, rtol=1e-2)
            assert np.all(np.abs(I[:, 0] - (math.sin(3.0) - math.sin(1.0)) ** dim) <= 1e-5)


def test_round_synchronous_restatement_in_one_dimension():
    """orc_hcubature_rounds with HCubature's 1-D rule (GK(7,15) per segment): the checker of the 1-D BatchIntegralFunction bridge.
    README example: the rounds bisect exactly the segments QuadGK's worst-first loop bisects (75 evaluations, E = 5.69e-9);
    infinite domain; agreement with the reference-order solver within the tolerance"""
    I, E, st = O.solve_hcubature_rounds(lambda x: math.sin(1.7 * x), [-2.0], [5.0], reltol=1e-10, abstol=1e-8)
    Iq, Eq, neq, stq = O.solve_quadgk(lambda x: math.sin(1.7 * x), -2.0, 5.0, reltol=1e-10, abstol=1e-8)
    assert st["status"] == 0 and 15 * st["rule_applications"] == neq == 75 and abs(I - Iq) <= 1e-15 and abs(E - Eq) <= 1e-15
    assert abs(I - (math.cos(-3.4) - math.cos(8.5)) / 1.7) < 1e-14
    I, E, st = O.solve_hcubature_rounds(lambda x: math.exp(-x * x), [-INF], [INF], reltol=1e-9, abstol=0.0)
    Ih, Eh, neh, sth = O.solve_hcubature(lambda x: math.exp(-x * x), [-INF], [INF], reltol=1e-9, abstol=0.0)
    assert st["status"] == 0 and sth == 0 and abs(I - math.sqrt(math.pi)) <= 1e-9 * I and abs(I - Ih) <= 1e-9 * I and E <= 1e-9 * I
    assert 15 * st["rule_applications"] <= 2 * neh


def test_quadgk_published_example_of_the_dependency():
    """QuadGK.jl is not vendored in the reference (compat "2.11", Project.toml:65), so its own published example is the one
    reference-EXECUTED vector available for the adaptive path: the package README's quick start
        julia> integral, err = quadgk(x -> exp(-x^2), 0, 1, rtol=1e-8)
        (0.746824132812427, 7.887024366937112e-13)
    (quoted from the README; there is no network here to re-fetch it -- the agreement of all 16 digits of the error estimate,
    a difference of two 15-point sums, is what identifies the quotation).  orc_quadgk returns exactly these doubles: the rule's
    node / weight tables, the order of the Kronrod and Gauss sums and E = |K - G| are QuadGK's."""
    I, E, ne, st = O.solve_quadgk(lambda x: math.exp(-x * x), 0.0, 1.0, reltol=1e-8, abstol=0.0)
    assert (I, E, ne, st) == (0.746824132812427, 7.887024366937112e-13, 15, 0)
    # the same integral through the registered family (c x^k exp(-a^2 (x-u)^2), a = 1) and the batch driver
    Ib, Eb, neb, stb = O.quadgk_batch("gausspoly", 0.0, 1.0, np.array([[1.0], [0.0], [0.0], [1.0]]), reltol=1e-8, abstol=0.0)
    assert Ib[0] == pytest.approx(0.746824132812427, rel=1e-15) and Eb[0] == pytest.approx(7.887024366937112e-13, rel=1e-3) and neb[0] == 15


def test_genz_malik_rule_against_an_independent_implementation():
    """HCubature.jl is not vendored and publishes no output vector, so the Genz-Malik rule of the oracle (gm_rule: degree-7 value and
    the embedded degree-5 difference E = |I7 - I5|, Genz & Malik 1980 as HCubature.jl implements it) is pinned against an
    INDEPENDENT implementation of the same published rule: SciPy's scipy.integrate._rules.GenzMalikCubature (nodes and weights built
    from the paper by other authors).  Random boxes and Gaussians, d = 2..6: value and error estimate agree to rounding
    (1e-15 x volume; E itself is a difference of two sums of that size)."""
    rules = pytest.importorskip("scipy.integrate._rules")
    rng = np.random.default_rng(0)
    for d in (2, 3, 4, 5, 6):
        rule = rules.GenzMalikCubature(d)
        for _ in range(5):
            a = rng.uniform(-1, 0.5, d); b = a + rng.uniform(0.1, 1.5, d)
            al = rng.uniform(0.5, 3, d); u = rng.uniform(0, 1, d)
            Is = float(rule.estimate(lambda x: np.exp(-np.sum((al * (x - u)) ** 2, axis=-1)), a, b))
            Es = float(rule.estimate_error(lambda x: np.exp(-np.sum((al * (x - u)) ** 2, axis=-1)), a, b))
            I, E, kd = O.cubrule(lambda x: math.exp(-sum((al[k] * (x[k] - u[k])) ** 2 for k in range(d))), a, b)
            vol = float(np.prod(b - a))
            assert abs(I - Is) <= 1e-15 * vol and abs(E - Es) <= 1e-15 * vol and 0 <= kd < d
    # the (7,15) Kronrod value against SciPy's table (its error heuristic differs from QuadGK's |K - G| only in rounding here)
    gk = rules.GaussKronrodQuadrature(15)
    Ik, Ek = O.gk15(lambda x: math.exp(-x * x), 0.0, 1.0)
    Isk = float(gk.estimate(lambda x: np.exp(-x[..., 0] ** 2), np.array([0.0]), np.array([1.0])))
    assert abs(Ik - Isk) <= 4e-16


def test_sampled_rules_against_numpy_and_scipy():
    """independent cross-check of the sampled restatement (src/trapezoidal.jl:16-40, src/simpsons.jl:16-25): numpy's trapezoid on uniform
    and non-uniform grids; SciPy's composite Simpson 1/3 rule on NON-uniform grids with an odd number of points (src/simpsons.jl:40-82;
    the reference's uniform "Simpson" is the end-corrected 17-59-43-49 / 48 rule, which SciPy does not have); agreement to rounding
    relative to sum |w y|"""
    integrate = pytest.importorskip("scipy.integrate")
    trapz = getattr(np, "trapezoid", None) or np.trapz
    rng = np.random.default_rng(4)
    for n in (2, 3, 10, 101, 1000):
        y = np.asfortranarray(rng.standard_normal((5, n)) + 1.5)
        x = np.sort(rng.random(n)) * 3.0 - 1.0
        h = 0.37
        scale = np.abs(y).sum(axis=1) * h
        assert np.all(np.abs(O.sampled_evalrule("trapezoid", y, h=h, dim=1) - trapz(y, dx=h, axis=1)) <= 1e-14 * scale)
        assert np.all(np.abs(O.sampled_evalrule("trapezoid", y, x=x, dim=1) - trapz(y, x=x, axis=1)) <= 1e-14 * np.abs(y).sum(axis=1) * 3.0)
        if n % 2 == 1 and n >= 3:
            ws = np.abs(O.sampled_weights("simpson", n, x=x))
            assert np.all(np.abs(O.sampled_evalrule("simpson", y, x=x, dim=1) - integrate.simpson(y, x=x, axis=1)) <= 1e-13 * (np.abs(y) * ws).sum(axis=1))


def test_kronrod_tables_against_scipy():
    """the generated Gauss-Kronrod tables (kronrod.kronrod(n), QuadGK.kronrod's layout) against SciPy's published 15- and 21-point
    tables: identical Kronrod nodes and weights, Gauss weights to 1e-15"""
    rules = pytest.importorskip("scipy.integrate._rules")
    from integrals_b200.kronrod import kronrod
    for npts, order in ((15, 7), (21, 10)):
        r = rules.GaussKronrodQuadrature(npts)
        nodes, weights = (np.asarray(v).ravel() for v in r.nodes_and_weights)
        idx = np.argsort(nodes)
        x, w, wg = kronrod(order)
        assert np.array_equal(nodes[idx][:order + 1], x) and np.array_equal(weights[idx][:order + 1], w)
        ln, lw = (np.asarray(v).ravel() for v in r.lower_nodes_and_weights)
        assert np.allclose(lw[np.argsort(ln)][:len(wg)], wg, rtol=0, atol=2e-15)
