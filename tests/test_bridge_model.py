"""CPU checks of the BatchIntegralFunction bridge's round scheme and host loop on the Python restatement of
its device kernels (tests/bridge_model.py): the scheme refines exactly like the C oracle orc_hcubature_rounds with the same
caller-evaluated integrand; Engine.integrate_batch drives the protocol correctly; solve() routes BatchIntegralFunction problems.
The CUDA kernels themselves are checked by tests/test_gpu_zz_batch_bridge.py."""
import math

import numpy as np
import pytest

import integrals_b200 as ib
import oracle as O
from bridge_model import BridgeModel
from integrals_b200 import _lib

INF = math.inf


class ModelEngine(BridgeModel):
    """the model behind Engine's method names, with Engine's own host loop"""
    device = 0
    integrate_batch = _lib.Engine.integrate_batch

    def synchronize(self):
        pass


def _rational(dim):
    a = [3.0, 5.0, 2.0, 4.0][:dim]; u = [0.3, 0.6, 0.5, 0.2][:dim]

    def batch(pts):
        x = pts.reshape(1, -1) if dim == 1 else pts
        s = np.ones(x.shape[1])
        for k in range(dim):
            s = s + a[k] * (x[k] - u[k]) * (x[k] - u[k])
        return 1.0 / s

    def point(x):
        x = np.atleast_1d(x)
        s = 1.0
        for k in range(dim):
            s = s + a[k] * (x[k] - u[k]) * (x[k] - u[k])
        return 1.0 / s
    return batch, point


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_round_scheme_with_caller_evaluated_integrand_matches_the_oracle(dim):
    batch, point = _rational(dim)
    kw = dict(reltol=1e-7, abstol=0.0, max_boxes=5000)
    I, E, st = ModelEngine().integrate_batch(batch, dim, [0.0] * dim, [1.0] * dim, **kw)
    Ir, Er, sr = O.solve_hcubature_rounds(point, [0.0] * dim, [1.0] * dim, **kw)
    assert (st["rounds"], st["boxes"], st["rule_applications"], st["status"]) == (sr["rounds"], sr["boxes"], sr["rule_applications"], sr["status"])
    assert abs(I - Ir) <= 1e-13 * abs(Ir) and abs(E - Er) <= 1e-13 * abs(Ir) and 1 <= st["batches"] <= st["rounds"]


def test_round_scheme_infinite_domains_and_budget():
    f = lambda x: 1.0 / (1.0 + x * x)
    for tr, trn in enumerate(["if_inf", "tan", "cot"]):
        I, E, st = ModelEngine().integrate_batch(f, 1, [-INF], [INF], reltol=1e-9, abstol=0.0, transform=tr)
        Ir, Er, sr = O.solve_hcubature_rounds(f, [-INF], [INF], tr=trn, reltol=1e-9, abstol=0.0)
        assert (st["rounds"], st["boxes"]) == (sr["rounds"], sr["boxes"]) and abs(I - Ir) <= 1e-13 * math.pi and abs(I - math.pi) <= 1e-9 * math.pi
        # rounds that bisect nothing (threshold above every stored error) are run inside the engine, not handed to the caller
        assert st["batches"] <= st["rounds"] and (tr != 0 or st["batches"] < st["rounds"])
    batch, point = _rational(2)
    I, E, st = ModelEngine().integrate_batch(batch, 2, [0.0, 0.0], [1.0, 1.0], reltol=1e-13, abstol=0.0, max_boxes=100)
    Ir, Er, sr = O.solve_hcubature_rounds(point, [0.0, 0.0], [1.0, 1.0], reltol=1e-13, abstol=0.0, max_boxes=100)
    assert st["status"] == 1 == sr["status"] and st["rule_applications"] == sr["rule_applications"] >= 100 and abs(I - Ir) <= 1e-13 * abs(Ir)
    I, E, st = ModelEngine().integrate_batch(batch, 2, [0.0, 0.0], [1.0, 1.0], reltol=1e-13, abstol=0.0, max_boxes=10 ** 6, store_boxes=32)
    assert st["status"] == 1 and st["boxes"] <= 32


def test_host_loop_and_solve_routing_on_the_model():
    e = ModelEngine()
    seen = []

    def f(pts, p):
        seen.append(pts.shape)
        return p * np.cos(pts[0]) * np.cos(pts[1])
    sol = ib.solve(ib.IntegralProblem(ib.BatchIntegralFunction(f), (np.ones(2), 3 * np.ones(2)), 2.0), ib.HCubatureB200(), engine=e,
                   reltol=1e-8, abstol=0.0)
    assert abs(sol.u - 2.0 * (math.sin(3.0) - math.sin(1.0)) ** 2) <= 1e-8 * abs(sol.u) and sol.resid <= 1e-8 * abs(sol.u)
    assert all(s[0] == 2 and s[1] % 17 == 0 for s in seen) and seen[0] == (2, 17) and sol.stats["batches"] == len(seen)
    sol = ib.solve(ib.IntegralProblem(ib.BatchIntegralFunction(lambda x, p: np.sin(p * x)), (-2.0, 5.0), 1.7), ib.QuadGKB200(), engine=e, reltol=1e-10)
    assert abs(sol.u - (math.cos(-3.4) - math.cos(8.5)) / 1.7) < 1e-14 and sol.stats["nevals"] == 75         # README.md:30-40
    with pytest.raises(ValueError, match="must return"):
        e.integrate_batch(lambda x: np.ones(3), 1, [0.0], [1.0])
    # max_batch (BatchIntegralFunction's max_batch): same points, same result, f sees at most max_batch of them per call
    sizes = []

    def g(pts, p):
        sizes.append(pts.shape[-1])
        return np.cos(pts[0]) * np.cos(pts[1]) * np.cos(pts[2])
    dom = (np.ones(3), 3 * np.ones(3))
    whole = ib.solve(ib.IntegralProblem(ib.BatchIntegralFunction(g), dom), ib.HCubatureB200(), engine=e, reltol=1e-7, abstol=0.0)
    n_whole = list(sizes); sizes.clear()
    part = ib.solve(ib.IntegralProblem(ib.BatchIntegralFunction(g, max_batch=100), dom), ib.HCubatureB200(), engine=e, reltol=1e-7, abstol=0.0)
    assert part.u == whole.u and part.resid == whole.resid and max(sizes) <= 100 and sum(sizes) == sum(n_whole) == whole.stats["nevals"]
    assert max(n_whole) > 100
    h = lambda x, p: np.sin(1.7 * x)
    s1 = ib.solve(ib.IntegralProblem(ib.BatchIntegralFunction(h, max_batch=7), (-2.0, 5.0)), ib.QuadGKB200(), engine=e, reltol=1e-10)
    assert abs(s1.u - (math.cos(-3.4) - math.cos(8.5)) / 1.7) < 1e-14
    with pytest.raises(ValueError):
        ib.BatchIntegralFunction(h, max_batch=0)
    # the reference's in-place style f(y, pts, p) (src/Integrals.jl:277-297)

    def h_iip(y, x, p):
        y[:] = np.sin(p * x)
    s2 = ib.solve(ib.IntegralProblem(ib.BatchIntegralFunction(h_iip, inplace=True), (-2.0, 5.0), 1.7), ib.QuadGKB200(), engine=e, reltol=1e-10)
    assert s2.u == s1.u and s2.resid == s1.resid

    def g_iip(y, pts, p):
        y[:] = np.cos(pts[0]) * np.cos(pts[1]) * np.cos(pts[2])
    part2 = ib.solve(ib.IntegralProblem(ib.BatchIntegralFunction(g_iip, inplace=True, max_batch=64), dom), ib.HCubatureB200(), engine=e, reltol=1e-7, abstol=0.0)
    assert part2.u == whole.u


def test_pointwise_integral_function_on_the_model():
    """IntegralFunction(f): the reference's pointwise closure f(x, p), evaluated by the caller at the engine's points; the reference's
    scalar known answers (test/interface_tests.jl:129-187: f = 1 and prod(cos.(x)) on [1,3]^dim, reltol = abstol = 1e-5; README.md:30-40)"""
    e = ModelEngine()
    for dim in (1, 2, 3):
        alg = ib.QuadGKB200() if dim == 1 else ib.HCubatureB200()
        dom = (1.0, 3.0) if dim == 1 else (np.ones(dim), 3 * np.ones(dim))
        sol = ib.solve(ib.IntegralProblem(ib.IntegralFunction(lambda x, p: 1.0), dom), alg, engine=e, reltol=1e-5, abstol=1e-5)
        assert abs(sol.u - 2.0 ** dim) <= 1e-12
        f = (lambda x, p: math.cos(x)) if dim == 1 else (lambda x, p: float(np.prod(np.cos(x))))
        sol = ib.solve(ib.IntegralProblem(ib.IntegralFunction(f), dom), alg, engine=e, reltol=1e-5, abstol=1e-5)
        assert abs(sol.u - (math.sin(3.0) - math.sin(1.0)) ** dim) <= 1e-5
    sol = ib.solve(ib.IntegralProblem(ib.IntegralFunction(lambda x, p: math.sin(x * p)), (-2.0, 5.0), 1.7), ib.QuadGKB200(), engine=e, reltol=1e-10)
    assert abs(sol.u - (math.cos(-3.4) - math.cos(8.5)) / 1.7) < 1e-14 and sol.stats["nevals"] == 75
    # the same answer as the reference-order CPU restatement with the same closure
    Iq, Eq, neq, stq = O.solve_quadgk(lambda x: math.sin(x * 1.7), -2.0, 5.0, reltol=1e-10, abstol=1e-8)
    assert abs(sol.u - Iq) <= 1e-15 and abs(sol.resid - Eq) <= 1e-15
    with pytest.raises(TypeError, match="registered"):
        ib.IntegralProblem(lambda x, p: x, (0.0, 1.0))          # plain callables stay rejected: the opt-in is explicit
