"""GPU parity tests of the BatchIntegralFunction bridge (csrc/batch_bridge.cu): the engine owns the
points and the adaptive refinement of one integral on the device, the CALLER evaluates the integrand on each round's batch
(reference: BatchIntegralFunction under QuadGKJL, src/Integrals.jl:274-315).  Checker: orc_hcubature_rounds with the same
integrand -- same rounds, boxes and rule applications, (I, E) equal to summation-order rounding -- closed forms, and the
reference's batched-integrand known answers (test/interface_tests.jl:189-258).
(The file name sorts this module after the other GPU modules.)"""
import math

import numpy as np
import pytest

import integrals_b200 as ib

pytestmark = pytest.mark.gpu
INF = math.inf


@pytest.fixture(scope="module")
def eng():
    return ib.Engine(0)


@pytest.fixture(scope="module")
def O():
    import oracle
    return oracle


def _rational(dim):
    """1 / (1 + sum_k a_k (x_k - u_k)^2) written with IEEE basic operations only, in a fixed order: the vectorised (numpy) and
    the pointwise (oracle callback) evaluations return identical doubles"""
    a = [3.0, 5.0, 2.0, 4.0, 1.5][:dim]; u = [0.3, 0.6, 0.5, 0.2, 0.7][:dim]

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


@pytest.mark.parametrize("dim", [1, 2, 3, 5])
def test_bridge_refines_like_the_oracle(eng, O, dim):
    batch, point = _rational(dim)
    lb, ub = [0.0] * dim, [1.0] * dim
    kw = dict(reltol=1e-7 if dim < 5 else 1e-5, abstol=0.0, max_boxes=20000)
    I, E, st = eng.integrate_batch(batch, dim, lb, ub, **kw)
    Ir, Er, sr = O.solve_hcubature_rounds(point, lb, ub, **kw)
    assert st["status"] == 0 and sr["status"] == 0
    assert (st["rounds"], st["boxes"], st["rule_applications"]) == (sr["rounds"], sr["boxes"], sr["rule_applications"])
    assert abs(I - Ir) <= 1e-13 * abs(Ir) and abs(E - Er) <= 1e-13 * abs(Ir) and E <= kw["reltol"] * abs(I)
    npts = 15 if dim == 1 else 1 + 4 * dim + 2 * dim * (dim - 1) + (1 << dim)
    assert st["nevals"] == npts * st["rule_applications"] and 1 <= st["batches"] <= st["rounds"]
    # the reference-order solver (one box per step) agrees within the tolerance
    Ih, Eh, neh, sth = O.solve_hcubature(point, lb, ub, reltol=kw["reltol"], abstol=0.0)
    assert sth == 0 and abs(I - Ih) <= kw["reltol"] * abs(Ih)


def test_bridge_infinite_domains_readme_example_and_budget(eng, O):
    # README.md:30-40 through the 1-D bridge: QuadGK's answer, error estimate and 75 evaluations
    I, E, st = eng.integrate_batch(lambda x: np.sin(1.7 * x), 1, [-2.0], [5.0], reltol=1e-10, abstol=1e-8)
    assert abs(I - (math.cos(-3.4) - math.cos(8.5)) / 1.7) < 1e-14 and E == pytest.approx(5.688e-9, rel=1e-3) and st["nevals"] == 75
    # Lorentzian over R under the three transformations; 2-D product on R x [0, inf)
    for tr, trn in enumerate(["if_inf", "tan", "cot"]):
        I, E, st = eng.integrate_batch(lambda x: 1.0 / (1.0 + x * x), 1, [-INF], [INF], reltol=1e-9, abstol=0.0, transform=tr)
        Ir, Er, sr = O.solve_hcubature_rounds(lambda x: 1.0 / (1.0 + x * x), [-INF], [INF], tr=trn, reltol=1e-9, abstol=0.0)
        assert st["status"] == 0 and abs(I - math.pi) <= 1e-9 * math.pi and abs(I - Ir) <= 1e-13 * math.pi
        assert (st["rounds"], st["boxes"]) == (sr["rounds"], sr["boxes"])
        f2 = lambda x: 1.0 / ((1.0 + x[0] * x[0]) * (1.0 + x[1]) * (1.0 + x[1]))
        I, E, st = eng.integrate_batch(f2, 2, [-INF, 0.0], [INF, INF], reltol=1e-6, abstol=0.0, transform=tr)
        Ir, Er, sr = O.solve_hcubature_rounds(f2, [-INF, 0.0], [INF, INF], tr=trn, reltol=1e-6, abstol=0.0)
        assert st["status"] == 0 and abs(I - math.pi) <= 1e-5 * math.pi and abs(I - Ir) <= 1e-12 * math.pi
        assert (st["rounds"], st["boxes"], st["rule_applications"]) == (sr["rounds"], sr["boxes"], sr["rule_applications"])
    # budget of rule applications: status 1, same refinement as the oracle at the same budget
    batch, point = _rational(3)
    I, E, st = eng.integrate_batch(batch, 3, [0.0] * 3, [1.0] * 3, reltol=1e-13, abstol=0.0, max_boxes=300)
    Ir, Er, sr = O.solve_hcubature_rounds(point, [0.0] * 3, [1.0] * 3, reltol=1e-13, abstol=0.0, max_boxes=300)
    assert st["status"] == 1 and sr["status"] == 1 and st["rule_applications"] == sr["rule_applications"] >= 300
    assert abs(I - Ir) <= 1e-13 * abs(Ir)
    # a full store ends the solve the same way
    I, E, st = eng.integrate_batch(batch, 3, [0.0] * 3, [1.0] * 3, reltol=1e-13, abstol=0.0, max_boxes=10 ** 6, store_boxes=64)
    assert st["status"] == 1 and st["boxes"] <= 64 and abs(I - Ir) <= 1e-4 * abs(Ir)


def test_bridge_through_solve_and_call_order(eng):
    """solve(IntegralProblem(BatchIntegralFunction(f), domain, p), alg): the reference's batched-integrand known answers
    (test/interface_tests.jl:189-258: f = 1 and prod(cos.(x)) on [1,3]^dim, reltol = abstol = 1e-5); protocol errors"""
    for dim in (1, 2, 3):
        alg = ib.QuadGKB200() if dim == 1 else ib.HCubatureB200()
        dom = (1.0, 3.0) if dim == 1 else (np.ones(dim), 3 * np.ones(dim))
        ones = ib.BatchIntegralFunction(lambda pts, p: np.ones(pts.shape[-1]))
        sol = ib.solve(ib.IntegralProblem(ones, dom), alg, engine=eng, reltol=1e-5, abstol=1e-5)
        assert sol.retcode == ib.ReturnCode.Success and abs(sol.u - 2.0 ** dim) <= 1e-12 and sol.stats["batches"] == 1
        cosp = ib.BatchIntegralFunction(lambda pts, p: np.cos(p * pts) if pts.ndim == 1 else np.prod(np.cos(p * pts), axis=0))
        sol = ib.solve(ib.IntegralProblem(cosp, dom, 1.0), alg, engine=eng, reltol=1e-5, abstol=1e-5)
        assert abs(sol.u - (math.sin(3.0) - math.sin(1.0)) ** dim) <= 1e-5 and sol.resid <= 1e-5
    with pytest.raises(ValueError, match="one-dimensional"):
        ib.solve(ib.IntegralProblem(ones, (np.ones(2), 3 * np.ones(2))), ib.QuadGKB200(), engine=eng)
    with pytest.raises(TypeError, match="BatchIntegralFunction"):
        ib.solve(ib.IntegralProblem(ones, (1.0, 3.0)), ib.GaussLegendreB200(n=4), engine=eng)
    with pytest.raises(ValueError, match="must return"):
        eng.integrate_batch(lambda x: np.ones(3), 1, [0.0], [1.0])
    eng.batch_open(2, [0.0, 0.0], [1.0, 1.0])
    n = eng.batch_next()
    assert n == 17
    with pytest.raises(ib.IB200Error, match="no points handed out"):
        eng.batch_values(np.zeros(n), n)
    with pytest.raises(ib.IB200Error, match="ib200_batch_next"):
        eng.batch_points(np.zeros((2, n + 1), order="F"), n + 1)
    with pytest.raises(ib.IB200Error, match="has not finished"):
        eng.batch_result()
    x = np.empty((2, n), order="F")
    eng.batch_points(x, n)
    assert np.all((x > 0.0) & (x < 1.0)) and np.allclose(x[:, 0], 0.5)          # the first point of the rule is the centre
    with pytest.raises(ib.IB200Error, match="no round pending"):
        eng.batch_next()
    eng.batch_values(np.ones(n), n)                                             # a constant: exact, finished after one round
    assert eng.batch_next() == 0
    I, E, st = eng.batch_result()
    assert abs(I - 1.0) <= 1e-14 and st["rounds"] == 1 and st["status"] == 0


def test_bridge_device_evaluated_integrand(eng):
    """device=True: the points stay on the GPU (torch tensor of shape (n, dim)), the integrand is a torch expression, the values
    are read in place; same refinement as the host-evaluated run of the same arithmetic"""
    import torch
    dim = 3
    batch, _ = _rational(dim)
    a = torch.tensor([3.0, 5.0, 2.0], dtype=torch.float64, device="cuda"); u = torch.tensor([0.3, 0.6, 0.5], dtype=torch.float64, device="cuda")

    def dev(x):
        assert x.is_cuda and x.shape[1] == dim
        s = torch.ones(x.shape[0], dtype=torch.float64, device=x.device)
        for k in range(dim):
            s = s + a[k] * (x[:, k] - u[k]) * (x[:, k] - u[k])
        return 1.0 / s
    Id, Ed, sd = eng.integrate_batch(dev, dim, [0.0] * dim, [1.0] * dim, reltol=1e-7, abstol=0.0, device=True)
    Ih, Eh, sh = eng.integrate_batch(batch, dim, [0.0] * dim, [1.0] * dim, reltol=1e-7, abstol=0.0)
    assert sd["status"] == 0 and (sd["rounds"], sd["boxes"]) == (sh["rounds"], sh["boxes"]) and abs(Id - Ih) <= 1e-13 * abs(Ih)
    g = ib.BatchIntegralFunction(lambda x, p: torch.exp(-((p * (x - 0.5)) ** 2).sum(dim=1)), device=True)    # Gaussian, closed form
    sol = ib.solve(ib.IntegralProblem(g, (np.zeros(2), np.ones(2)), 3.0), ib.HCubatureB200(), engine=eng, reltol=1e-8, abstol=0.0)
    exact = (math.sqrt(math.pi) / 3.0 * math.erf(1.5)) ** 2
    assert abs(sol.u - exact) <= 1e-7 * exact and sol.resid <= 1e-8 * exact
