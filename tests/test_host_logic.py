"""Host-side mirror of the reference's solve interface (integrals.jl_b200/interface.py) without a GPU: argument
merging, the kwarg whitelist, automatic ChangeOfVariables, error behaviour, shapes handed to the engine.  A recording
stub stands in for the engine, so no kernel runs here (GPU behaviour: tests/test_gpu_parity.py)."""
import math

import numpy as np
import pytest

import integrals_b200 as ib


class StubEngine:
    def __init__(self):
        self.calls = []

    def _rec(self, name, *a, **k):
        self.calls.append((name, a, k))

    def quadgk_batch(self, fam, lb, ub, p, **k):
        self._rec("quadgk", fam, lb, ub, p, **k); n = p.shape[1]
        return np.zeros(n), np.zeros(n), np.zeros(n, np.int64), np.zeros(n, np.int32)

    def hcubature_batch(self, fam, dim, lb, ub, p, **k):
        self._rec("hcubature", fam, dim, lb, ub, p, **k); n = p.shape[1]
        return np.zeros(n), np.zeros(n), np.zeros(n, np.int64), np.zeros(n, np.int32)

    def hcubature_single(self, fam, dim, lb, ub, p, **k):
        self._rec("hcubature_single", fam, dim, lb, ub, p, **k)
        return 0.0, 0.0, dict(status=0)

    def fixed_rule_tensor_batch(self, fam, dim, lb, ub, p, x, w, **k):
        self._rec("tensor", fam, dim, lb, ub, p, x, w, **k); return np.zeros(p.shape[1])

    def fixed_rule_nodes_batch(self, fam, dim, lb, ub, p, x, w, **k):
        self._rec("nodes", fam, dim, lb, ub, p, x, w, **k); return np.zeros(p.shape[1])

    def gauss_legendre_batch(self, fam, lb, ub, p, x, w, **k):
        self._rec("gl", fam, lb, ub, p, x, w, **k); return np.zeros(p.shape[1])

    def sampled(self, rule, y, inner, n, outer, h=0.0, x=None, out=None):
        self._rec("sampled", rule, inner, n, outer, h=h, x=x); return np.zeros(max(inner * outer, 1))


def test_kwarg_whitelist():      # src/Integrals.jl:142-179
    prob = ib.IntegralProblem(ib.SinXP(), (-2.0, 5.0), 1.7)
    with pytest.raises(ib.CommonKwargError, match="Unrecognized keyword arguments"):
        ib.solve(prob, ib.QuadGKB200(), engine=StubEngine(), rtol=1e-3)
    with pytest.raises(ib.CommonKwargError):
        ib.solve(ib.IntegralProblem(ib.SinXP(), (-2.0, 5.0), 1.7, tol=1.0), ib.QuadGKB200(), engine=StubEngine())
    with pytest.raises(ValueError, match="No integration algorithm"):      # src/common.jl:163-177
        ib.solve(prob, None, engine=StubEngine())


def test_defaults_and_kwarg_merge():   # defaults src/Integrals.jl:254-255; problem kwargs under solve kwargs, src/common.jl:87
    e = StubEngine()
    prob = ib.IntegralProblem(ib.SinXP(), (-2.0, 5.0), 1.7, reltol=1e-5)
    sol = ib.solve(prob, ib.QuadGKB200(), engine=e)
    k = e.calls[-1][2]
    assert k["reltol"] == 1e-5 and k["abstol"] == 1e-8 and k["maxevals"] == ib.MAXEVALS_UNBOUNDED and k["transform"] == 0
    assert isinstance(sol.u, float) and sol.retcode == ib.ReturnCode.Success and sol.chi is None
    ib.solve(prob, ib.QuadGKB200(), engine=e, reltol=1e-9, maxiters=1000)
    k = e.calls[-1][2]
    assert k["reltol"] == 1e-9 and k["maxevals"] == 1000


def test_automatic_change_of_variables_and_explicit():   # src/common.jl:97-103, src/algorithms_meta.jl:67-71
    e = StubEngine()
    prob = ib.IntegralProblem(ib.GenzGaussian(), ([-math.inf, 0.0], [math.inf, 1.0]), [1.0, 2.0, 0.0, 0.5])
    cache = ib.init(prob, ib.HCubatureB200(), engine=e)
    assert isinstance(cache.alg, ib.ChangeOfVariables) and cache.alg.fu2gv is ib.transformation_if_inf
    cache.solve()
    name, a, k = e.calls[-1]
    assert name == "hcubature" and a[1] == 2 and k["transform"] == 0 and np.isinf(a[2][0])   # ORIGINAL bounds go to the device
    ib.solve(prob, ib.ChangeOfVariables(ib.transformation_tan_inf, ib.HCubatureB200()), engine=e)
    assert e.calls[-1][2]["transform"] == 1
    ib.solve(prob, ib.ChangeOfVariables(ib.transformation_cot_inf, ib.HCubatureB200(initdiv=2)), engine=e)
    assert e.calls[-1][2]["transform"] == 2 and e.calls[-1][2]["initdiv"] == 2
    with pytest.raises(TypeError):
        ib.ChangeOfVariables(lambda f, d: (f, d), ib.HCubatureB200())


def test_dimension_errors():     # src/Integrals.jl:263-266; ext/IntegralsFastGaussQuadratureExt.jl:41-43
    e = StubEngine()
    prob2 = ib.IntegralProblem(ib.GenzGaussian(), (np.zeros(2), np.ones(2)), [1.0, 2.0, 0.0, 0.5])
    with pytest.raises(ValueError, match="one-dimensional"):
        ib.solve(prob2, ib.QuadGKB200(), engine=e)
    with pytest.raises(ValueError, match="one-dimensional"):
        ib.solve(prob2, ib.GaussLegendreB200(n=8), engine=e)
    with pytest.raises(TypeError, match="registered"):
        ib.IntegralProblem(lambda x, p: x, (0.0, 1.0), 1.0)
    with pytest.raises(ValueError):
        ib.QuadratureRuleB200(lambda n: None, n=0)             # src/algorithms.jl:293-295
    with pytest.raises(ValueError, match="at least 1 subinterval"):
        ib.GaussLegendreB200(n=4, subintervals=0)


def test_batches_and_cache_reuse():   # caching interface: docs/src/tutorials/caching_interface.md:32-47
    e = StubEngine()
    P = np.asfortranarray(np.random.default_rng(0).random((1, 11)))
    cache = ib.init(ib.IntegralProblem(ib.SinXP(), (-2.0, 5.0), P), ib.QuadGKB200(), engine=e, reltol=1e-10)
    sol = cache.solve()
    assert sol.u.shape == (11,) and sol.resid.shape == (11,)
    cache.p = P[:, :3]                                          # cache.p = ...; solve!(cache)
    assert cache.solve().u.shape == (3,)
    cache.domain = (0.0, 1.0)
    cache.solve()
    assert float(e.calls[-1][1][1][0]) == 0.0 and float(e.calls[-1][1][2][0]) == 1.0


def test_single_domain_routing():
    e = StubEngine()
    prob = ib.IntegralProblem(ib.GenzGaussian(), (np.zeros(8), np.ones(8)), np.ones(16))
    ib.solve(prob, ib.HCubatureB200(single_domain=True), engine=e, reltol=1e-9, abstol=0.0, maxiters=401 * 1000)
    name, a, k = e.calls[-1]
    assert name == "hcubature_single" and k["max_boxes"] == 1000 and k["abstol"] == 0.0
    with pytest.raises(ValueError, match="one problem"):
        ib.solve(ib.IntegralProblem(ib.GenzGaussian(), (np.zeros(2), np.ones(2)), np.ones((4, 3))), ib.HCubatureB200(single_domain=True), engine=e)


def test_sampled_problem_shapes_and_errors():   # src/common.jl:319-343, src/sampled.jl:135-168
    e = StubEngine()
    y = np.zeros((4, 10, 3))
    sol = ib.solve(ib.SampledIntegralProblem(y, ib.Range(0.0, 1.0, 10), dim=1), ib.TrapezoidalB200(), engine=e)
    name, a, k = e.calls[-1]
    assert a[1:] == (4, 10, 3) and k["h"] == pytest.approx(1 / 9) and k["x"] is None and sol.u.shape == (4, 3)
    sol = ib.solve(ib.SampledIntegralProblem(y, np.linspace(0, 1, 3) ** 2, dim=-1), ib.SimpsonsB200(), engine=e)
    assert e.calls[-1][1][1:] == (40, 3, 1) and e.calls[-1][2]["x"] is not None and sol.u.shape == (4, 10) and sol.resid is None
    with pytest.raises(ValueError, match="same length"):
        ib.SampledIntegralProblem(y, np.zeros(7), dim=1)
    with pytest.raises(ValueError, match="no keyword arguments"):
        ib.solve(ib.SampledIntegralProblem(y, ib.Range(0, 1, 3)), ib.TrapezoidalB200(), engine=e, reltol=1e-3)
    cache = ib.init(ib.SampledIntegralProblem(y, ib.Range(0, 1, 3)), ib.SimpsonsB200(), engine=e)
    assert cache.isfresh is True                                 # init: isfresh = true, src/common.jl:330
    cache.solve()
    assert cache.isfresh is False
    cache.x = ib.Range(0, 2, 3)                                  # src/common.jl:276-281: marks the weights stale
    assert cache.isfresh is True
    cache.solve()
    assert cache.isfresh is False and e.calls[-1][2]["h"] == 1.0


def test_sampled_float32_and_complex_routing():   # test/interface_tests.jl:501-539
    class E32(StubEngine):
        def sampled_f32(self, rule, y, inner, n, outer, h=0.0, x=None, out=None):
            self._rec("sampled_f32", rule, inner, n, outer, h=h, x=x, ydtype=y.dtype); return np.zeros(max(inner * outer, 1), np.float32)
    e = E32()
    y32 = np.sin(np.linspace(0, 1, 11)).astype(np.float32)
    sol = ib.solve(ib.SampledIntegralProblem(y32, ib.Range(0.0, 1.0, 11, dtype=np.float32)), ib.TrapezoidalB200(), engine=e)
    assert e.calls[-1][0] == "sampled_f32" and isinstance(sol.u, np.float32) and e.calls[-1][2]["h"] == pytest.approx(0.1, rel=1e-6)
    sol = ib.solve(ib.SampledIntegralProblem(np.zeros((3, 11), np.float32), np.linspace(0, 1, 11, dtype=np.float32), dim=1), ib.SimpsonsB200(), engine=e)
    assert e.calls[-1][0] == "sampled_f32" and sol.u.dtype == np.float32 and sol.u.shape == (3,)
    # Float32 samples on a Float64 grid promote to Float64, as Float32 * Float64 does in the reference
    ib.solve(ib.SampledIntegralProblem(y32, ib.Range(0.0, 1.0, 11)), ib.TrapezoidalB200(), engine=e)
    assert e.calls[-1][0] == "sampled"
    # complex samples: real weights -> the float64 kernel on [2*inner, n, outer]
    yc = np.exp(1j * np.linspace(0, 1, 11))
    sol = ib.solve(ib.SampledIntegralProblem(yc, ib.Range(0.0, 1.0, 11)), ib.TrapezoidalB200(), engine=e)
    assert e.calls[-1][0] == "sampled" and e.calls[-1][1][1:] == (2, 11, 1) and isinstance(sol.u, complex)
    sol = ib.solve(ib.SampledIntegralProblem(np.zeros((4, 11, 3), complex), ib.Range(0.0, 1.0, 11), dim=1), ib.SimpsonsB200(), engine=e)
    assert e.calls[-1][1][1:] == (8, 11, 3) and sol.u.shape == (4, 3) and np.iscomplexobj(sol.u)


def test_vector_valued_routing():    # src/Integrals.jl:316-326 (array-valued integrand, QuadGKJL.norm)
    class EV(StubEngine):
        def quadgk_vec_batch(self, fam, lb, ub, p, **k):
            self._rec("quadgk_vec", fam, lb, ub, p, **k)
            p = np.asarray(p); nc = p.shape[1]; n = p.shape[2] if p.ndim == 3 else 1
            return np.zeros((nc, n), order="F"), np.zeros(n), np.zeros(n, np.int64), np.zeros(n, np.int32)
    e = EV()
    f = ib.VectorOf(ib.SinXP())
    sol = ib.solve(ib.IntegralProblem(f, (-2.0, 5.0), np.array([[1.7, 3.3, 0.5]])), ib.QuadGKB200(norm="inf"), engine=e, reltol=1e-9)
    name, a, k = e.calls[-1]
    assert name == "quadgk_vec" and k["norm"] == "inf" and k["reltol"] == 1e-9 and sol.u.shape == (3,) and isinstance(sol.resid, float)
    sol = ib.solve(ib.IntegralProblem(f, (-2.0, 5.0), np.zeros((1, 3, 7))), ib.QuadGKB200(), engine=e)
    assert sol.u.shape == (3, 7) and sol.resid.shape == (7,) and e.calls[-1][2]["norm"] == "2"
    with pytest.raises(TypeError, match="QuadGKB200"):
        ib.solve(ib.IntegralProblem(f, (-2.0, 5.0), np.zeros((1, 3))), ib.GaussLegendreB200(n=4), engine=e)
    # HCubatureJL.norm (src/algorithms.jl:66-67,108-115; forwarded at src/Integrals.jl:380-393): N-D vector integrands
    class EH(StubEngine):
        def hcubature_vec_batch(self, fam, dim, lb, ub, p, **k):
            self._rec("hcubature_vec", fam, dim, lb, ub, p, **k)
            p = np.asarray(p); nc = p.shape[1]; n = p.shape[2] if p.ndim == 3 else 1
            return np.zeros((nc, n), order="F"), np.zeros(n), np.zeros(n, np.int64), np.zeros(n, np.int32)
    e = EH()
    g = ib.VectorOf(ib.CosProduct())
    sol = ib.solve(ib.IntegralProblem(g, (np.ones(2), 3 * np.ones(2)), np.ones((2, 2))), ib.HCubatureB200(norm="inf", initdiv=2), engine=e,
                   reltol=1e-5, abstol=1e-5, maxiters=10**6)
    name, a, k = e.calls[-1]
    assert name == "hcubature_vec" and a[1] == 2 and k["norm"] == "inf" and k["initdiv"] == 2 and k["maxevals"] == 10**6
    assert k["reltol"] == 1e-5 and k["abstol"] == 1e-5 and k["transform"] == 0 and sol.u.shape == (2,) and isinstance(sol.resid, float)
    sol = ib.solve(ib.IntegralProblem(g, (np.ones(3), 3 * np.ones(3)), np.ones((3, 4, 5))), ib.HCubatureB200(), engine=e)
    assert sol.u.shape == (4, 5) and sol.resid.shape == (5,) and e.calls[-1][2]["norm"] == "2" and sol.retcode == ib.ReturnCode.Success
    sol = ib.solve(ib.IntegralProblem(ib.VectorOf(ib.SinXP()), (1.0, 3.0), np.ones((1, 2))), ib.HCubatureB200(), engine=e)   # hquadrature
    assert e.calls[-1][1][1] == 1 and sol.u.shape == (2,)
    with pytest.raises(ValueError, match="not defined in 2 dimensions"):
        ib.solve(ib.IntegralProblem(ib.VectorOf(ib.SinXP()), (np.ones(2), 3 * np.ones(2)), np.ones((1, 2))), ib.HCubatureB200(), engine=e)
    with pytest.raises(ValueError, match="single_domain"):
        ib.solve(ib.IntegralProblem(g, (np.ones(2), 3 * np.ones(2)), np.ones((2, 2))), ib.HCubatureB200(single_domain=True), engine=e)
    with pytest.raises(ValueError):
        ib.HCubatureB200(norm="1")
    with pytest.raises(ValueError):
        ib.QuadGKB200(norm="1")
    with pytest.raises(TypeError):
        ib.VectorOf(ib.VectorOf(ib.SinXP()))


def test_init_named_keywords_are_accepted():   # src/common.jl:79-85: sensealg, do_inf_transformation, verbose are not "common kwargs"
    e = StubEngine()
    prob = ib.IntegralProblem(ib.SinXP(), (-2.0, 5.0), 1.7)
    sol = ib.solve(prob, ib.QuadGKB200(), engine=e, reltol=1e-6, verbose=False, sensealg=None, do_inf_transformation=True)
    assert sol.retcode == ib.ReturnCode.Success and e.calls[-1][2]["reltol"] == 1e-6


def test_batch_integral_function_routing():    # src/Integrals.jl:274-315 (BatchIntegralFunction); SURVEY 8(f) rank 3
    class EB(StubEngine):
        def integrate_batch(self, f, dim, lb, ub, **k):
            self._rec("integrate_batch", f, dim, lb, ub, **k)
            pts = np.zeros((dim, 4)) if dim > 1 else np.zeros(4)
            self.values = f(pts)
            return 1.5, 1e-9, dict(status=0, batches=1)
    e = EB()
    seen = {}

    def f(pts, p):
        seen["shape"], seen["p"] = pts.shape, p
        return np.ones(pts.shape[-1])
    sol = ib.solve(ib.IntegralProblem(ib.BatchIntegralFunction(f), (np.zeros(3), np.ones(3)), 2.5), ib.HCubatureB200(store_boxes=4096), engine=e,
                   reltol=1e-6, abstol=1e-7, maxiters=33 * 1000 + 5)
    name, a, k = e.calls[-1]
    assert name == "integrate_batch" and a[1] == 3 and k["reltol"] == 1e-6 and k["abstol"] == 1e-7 and k["max_boxes"] == 1000
    assert k["store_boxes"] == 4096 and k["transform"] == 0 and k["device"] is False and seen == {"shape": (3, 4), "p": 2.5}
    assert sol.u == 1.5 and sol.resid == 1e-9 and sol.retcode == ib.ReturnCode.Success and sol.stats["batches"] == 1
    sol = ib.solve(ib.IntegralProblem(ib.BatchIntegralFunction(f, device=True), (-2.0, np.inf)), ib.ChangeOfVariables(ib.transformation_tan_inf, ib.QuadGKB200()),
                   engine=e, maxiters=100)
    name, a, k = e.calls[-1]
    assert a[1] == 1 and k["max_boxes"] == 6 and k["transform"] == 1 and k["device"] is True and seen["shape"] == (4,) and seen["p"] is None
    with pytest.raises(ValueError, match="one-dimensional"):
        ib.solve(ib.IntegralProblem(ib.BatchIntegralFunction(f), (np.zeros(2), np.ones(2))), ib.QuadGKB200(), engine=e)
    with pytest.raises(ValueError, match="initdiv"):
        ib.solve(ib.IntegralProblem(ib.BatchIntegralFunction(f), (np.zeros(2), np.ones(2))), ib.HCubatureB200(initdiv=2), engine=e)
    with pytest.raises(TypeError, match="QuadGKB200"):
        ib.solve(ib.IntegralProblem(ib.BatchIntegralFunction(f), (0.0, 1.0)), ib.GaussLegendreB200(n=4), engine=e)
    with pytest.raises(ib.CommonKwargError):
        ib.solve(ib.IntegralProblem(ib.BatchIntegralFunction(f), (0.0, 1.0)), ib.QuadGKB200(), engine=e, tol=1e-3)
    with pytest.raises(TypeError):
        ib.BatchIntegralFunction(3.0)
    with pytest.raises(TypeError, match="BatchIntegralFunction"):
        ib.IntegralProblem(lambda x, p: x, (0.0, 1.0))


def test_sampled_nd_layout_property():
    """evalrule over an arbitrary axis of an N-D array (src/sampled.jl:99-124): the host layer hands the engine Julia's column-major
    [inner, n, outer] view; with the engine replaced by the oracle's kernel-equivalent on that flat view, solve() must equal numpy's
    trapezoid along the same axis for every shape / axis (hypothesis)"""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")
    import oracle as O
    trapz = getattr(np, "trapezoid", None) or np.trapz

    class OracleSampled(StubEngine):
        def sampled(self, rule, y, inner, n, outer, h=0.0, x=None, out=None):
            flat = np.ravel(np.asarray(y), order="F").reshape((inner, n, outer), order="F")      # the memory the C ABI would see
            u = O.sampled_evalrule(rule, flat, h=None if x is not None else h, x=x, dim=1)
            return np.ravel(np.atleast_1d(u), order="F")

    @hyp.settings(max_examples=60, deadline=None)
    @hyp.given(st.lists(st.integers(1, 5), min_size=1, max_size=4), st.integers(0, 3), st.integers(2, 9), st.booleans())
    def check(shape, dim, n, uniform):
        dim = dim % len(shape)
        shape = list(shape); shape[dim] = n
        rng = np.random.default_rng(sum(shape) + dim)
        y = rng.standard_normal(shape)
        x = ib.Range(0.0, 2.0, n) if uniform else np.sort(rng.random(n)) + np.arange(n)
        sol = ib.solve(ib.SampledIntegralProblem(y, x, dim=dim), ib.TrapezoidalB200(), engine=OracleSampled())
        ref = trapz(y, x=x.toarray() if uniform else x, axis=dim)
        assert np.shape(sol.u) == np.shape(ref) and np.allclose(sol.u, ref, rtol=1e-12, atol=1e-12)
    check()


def test_integrand_markers_evaluate_like_the_oracle_family():
    """every registered marker is also callable, f(x, p), with the definition the device code and the oracle share (csrc/family.cuh,
    orc_family_eval): the pointwise values agree to rounding for all 11 families, so the same problem object can be handed to a
    generic (closure-based) solver for parity checks -- INTEGRATION.md does the same with callable structs in Julia"""
    import oracle as O
    from util_gpu import ORC_FAM
    rng = np.random.default_rng(12)
    for cls in ib.ALL_INTEGRANDS:
        f = cls()
        for dim in (1, 2, 3, 5):
            if not f.supports_dim(dim):
                continue
            for _ in range(20):
                x = rng.uniform(0.05, 0.95, dim)
                if f.name == "gausspoly":
                    p = np.concatenate([rng.uniform(0.5, 3, dim), rng.uniform(0, 1, dim), rng.integers(0, 4, dim).astype(float), [rng.uniform(0.5, 2)]])
                else:
                    p = rng.uniform(0.2, 3.0, f.nparam(dim))
                got, ref = f(x if dim > 1 else x[0], p), O.family_eval(ORC_FAM[f.name], x, p)
                assert abs(got - ref) <= 1e-13 * max(1.0, abs(ref)), (f.name, dim, got, ref)


def test_retcode_from_engine_status():
    """status -> retcode: Success like the reference (src/Integrals.jl:338,396) when the tolerance or the caller's maxiters
    ended the solve; MaxIters when the engine's own box / segment store filled up first; QuadGK's DomainError for a non-finite estimate"""
    class E(StubEngine):
        def __init__(self, st, ne):
            super().__init__(); self.st, self.ne = st, ne

        def quadgk_batch(self, fam, lb, ub, p, **k):
            n = p.shape[1]
            return np.zeros(n), np.zeros(n), np.full(n, self.ne, np.int64), np.full(n, self.st, np.int32)
        hcubature_batch = lambda self, fam, dim, lb, ub, p, **k: self.quadgk_batch(fam, lb, ub, p, **k)
    prob = ib.IntegralProblem(ib.SinXP(), (-2.0, 5.0), 1.7)
    assert ib.solve(prob, ib.QuadGKB200(), engine=E(0, 75)).retcode == ib.ReturnCode.Success
    assert ib.solve(prob, ib.QuadGKB200(), engine=E(1, 1005), maxiters=1000).retcode == ib.ReturnCode.Success      # the caller's budget
    assert ib.solve(prob, ib.QuadGKB200(), engine=E(1, 250000)).retcode == ib.ReturnCode.MaxIters                  # the engine's capacity
    assert ib.solve(prob, ib.QuadGKB200(), engine=E(3, 300)).retcode == ib.ReturnCode.Success                      # endpoint round-off: QuadGK returns
    with pytest.raises(ib.DomainError):
        ib.solve(prob, ib.QuadGKB200(), engine=E(2, 15))
    p4 = ib.IntegralProblem(ib.GenzGaussian(), (np.zeros(4), np.ones(4)), np.ones((8, 3)))
    assert ib.solve(p4, ib.HCubatureB200(), engine=E(2, 57)).retcode == ib.ReturnCode.Success                      # HCubature stops and returns
    assert ib.solve(p4, ib.HCubatureB200(), engine=E(1, 57 * 1000)).retcode == ib.ReturnCode.MaxIters


def test_hcubature_refinement_modes():
    e = StubEngine()
    p4 = ib.IntegralProblem(ib.GenzGaussian(), (np.zeros(4), np.ones(4)), np.ones((8, 3)))
    ib.solve(p4, ib.HCubatureB200(), engine=e); assert e.calls[-1][2]["mode"] == "rounds"
    ib.solve(p4, ib.HCubatureB200(refinement="reference_order"), engine=e); assert e.calls[-1][2]["mode"] == "reference_order"
    p1 = ib.IntegralProblem(ib.GenzGaussian(), (np.zeros(1), np.ones(1)), np.ones((2, 3)))
    ib.solve(p1, ib.HCubatureB200(), engine=e); assert e.calls[-1][2]["mode"] == "reference_order"     # rounds are built for dim >= 2
    with pytest.raises(ValueError):
        ib.HCubatureB200(refinement="fastest")


def test_parameter_sensitivities_routing():
    """sensealg = ParameterSensitivities (the reference passes `sensealg` through init, src/common.jl:79-133, and differentiates solve in
    ext/IntegralsForwardDiffExt.jl / ext/IntegralsZygoteExt.jl): one engine call per group of 7 parameters, shapes of sol.u / sol.du_dp"""
    class ES(StubEngine):
        def sens_batch(self, fam, dim, lb, ub, p, wrt, **k):
            self._rec("sens", fam, dim, lb, ub, p, list(wrt), **k); n = p.shape[1]
            out = np.zeros((1 + len(wrt), n), order="F"); out[1:] = np.asarray(wrt, float)[:, None]
            return out, np.zeros(n), np.zeros(n, np.int64), np.zeros(n, np.int32)
    e = ES()
    par = np.full((8, 5), 1.0)                           # 4-D Genz Gaussian: 8 parameters, 5 problems
    prob = ib.IntegralProblem(ib.GenzGaussian(), (np.zeros(4), np.ones(4)), par)
    sol = ib.solve(prob, ib.HCubatureB200(initdiv=2), engine=e, sensealg=ib.ParameterSensitivities(), reltol=1e-7)
    sens = [c for c in e.calls if c[0] == "sens"]
    assert [c[1][5] for c in sens] == [[0, 1, 2, 3, 4, 5, 6], [7]]                  # 8 parameters: groups of 7 + 1
    assert all(c[2]["norm"] == "value" and c[2]["rule"] == "hcubature" and c[2]["initdiv"] == 2 and c[2]["reltol"] == 1e-7 for c in sens)
    assert sol.u.shape == (5,) and sol.du_dp.shape == (8, 5) and np.array_equal(sol.du_dp[:, 0], np.arange(8.0))
    # one problem, chosen parameters, a norm over value and derivatives; QuadGK in 1-D
    e.calls.clear()
    sol = ib.solve(ib.IntegralProblem(ib.SinXP(), (-2.0, 5.0), np.array([1.7])), ib.QuadGKB200(), engine=e,
                   sensealg=ib.ParameterSensitivities(wrt=[0], control="2"))
    assert e.calls[-1][2]["rule"] == "quadgk" and e.calls[-1][2]["norm"] == "2" and isinstance(sol.u, float) and sol.du_dp.shape == (1,)
    with pytest.raises(ValueError, match="wrt"):
        ib.solve(ib.IntegralProblem(ib.SinXP(), (-2.0, 5.0), np.array([1.7])), ib.QuadGKB200(), engine=e, sensealg=ib.ParameterSensitivities(wrt=[3]))
    with pytest.raises(TypeError, match="QuadGKB200 and HCubatureB200"):
        ib.solve(ib.IntegralProblem(ib.SinXP(), (-2.0, 5.0), np.array([1.7])), ib.GaussLegendreB200(n=4), engine=e, sensealg=ib.ParameterSensitivities())
    # any other sensealg (the reference's AD back-ends) is accepted and ignored
    sol = ib.solve(ib.IntegralProblem(ib.SinXP(), (-2.0, 5.0), np.array([1.7])), ib.QuadGKB200(), engine=e, sensealg="ReCallVJP(ZygoteVJP())")
    assert sol.du_dp is None and e.calls[-1][0] == "quadgk"
