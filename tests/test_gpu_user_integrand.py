"""GPU parity of device-resident user integrands (ib200_integrand_compile / ib200_integrate_compiled through the C ABI): an integrand given
as CUDA source, compiled with NVRTC and evaluated on the device inside the bridge's rounds, against the CPU oracle running the same
rounds with the same integrand as a Python callable (orc_solve_hcubature_rounds) and against closed forms.  Stands in for the
reference's closure f(x, p), /root/reference/src/Integrals.jl:327-335, 380-393 (README example: test/interface_tests.jl)."""
import math

import numpy as np
import pytest

import integrals_b200 as ib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    return ib.Engine(0)


@pytest.fixture(scope="module")
def O():
    import oracle
    return oracle


def test_compiled_integrand_matches_the_rounds_oracle(eng, O):
    # IEEE basic operations only: the device and the CPU callback return identical doubles -> identical rounds, boxes and rule applications
    h = eng.compile_integrand("return 1.0 / (1.0 + p[0] * x[0] * x[0] + p[1] * x[1] * x[1] * x[2]);", 3)
    par = np.array([3.0, 5.0])
    lb, ub = np.zeros(3), np.array([1.0, 2.0, 0.5])
    I, E, s = eng.integrate_compiled(h, 3, lb, ub, par, reltol=1e-8, abstol=0.0, max_boxes=1 << 18)
    Ir, Er, sr = O.solve_hcubature_rounds(lambda x: 1.0 / (1.0 + par[0] * x[0] * x[0] + par[1] * x[1] * x[1] * x[2]), lb, ub, reltol=1e-8, abstol=0.0,
                                          max_boxes=1 << 18)
    assert s["status"] == 0 and sr["status"] == 0
    assert (s["rounds"], s["boxes"], s["rule_applications"]) == (sr["rounds"], sr["boxes"], sr["rule_applications"])
    assert abs(I - Ir) <= 1e-13 * abs(Ir) and abs(E - Er) <= 1e-13 * abs(Ir)
    assert eng.kernel_launches > 0


def test_compiled_integrand_closed_forms_transforms_and_solve(eng):
    # README example (sin(x p) over [-2, 5]) through the public API, 1-D
    f = ib.CudaIntegralFunction("return sin(x[0] * p[0]);")
    sol = ib.solve(ib.IntegralProblem(f, (-2.0, 5.0), np.array([1.7])), ib.QuadGKB200(), engine=eng, reltol=1e-10, abstol=1e-12)
    assert abs(sol.u - (math.cos(-3.4) - math.cos(8.5)) / 1.7) <= 1e-12 and sol.retcode == ib.ReturnCode.Success
    # Gaussian over the whole plane (transformation_if_inf and the tan map): pi / (a b)
    g = ib.CudaIntegralFunction("return exp(-(p[0] * x[0]) * (p[0] * x[0]) - (p[1] * x[1]) * (p[1] * x[1]));")
    for alg in (ib.HCubatureB200(), ib.ChangeOfVariables(ib.transformation_tan_inf, ib.HCubatureB200())):
        sol = ib.solve(ib.IntegralProblem(g, (np.array([-np.inf, -np.inf]), np.array([np.inf, np.inf])), np.array([1.3, 0.7])), alg, engine=eng,
                       reltol=1e-8, abstol=1e-12)
        assert abs(sol.u - math.pi / (1.3 * 0.7)) <= 1e-7 * sol.u and sol.resid <= 1e-8 * sol.u * 1.0001
    # an integrand no registered family expresses: the volume of the unit ball octant (discontinuous), no parameters
    ball = ib.CudaIntegralFunction("return x[0] * x[0] + x[1] * x[1] + x[2] * x[2] <= 1.0 ? 1.0 : 0.0;")
    sol = ib.solve(ib.IntegralProblem(ball, (np.zeros(3), np.ones(3))), ib.HCubatureB200(), engine=eng, reltol=1e-3, abstol=0.0, maxiters=33 * 400000)
    assert abs(sol.u - math.pi / 6) <= 5e-3
    # a source that does not compile surfaces the compiler's message
    with pytest.raises(ib.IB200Error, match="does not compile"):
        ib.solve(ib.IntegralProblem(ib.CudaIntegralFunction("return y[0];"), (0.0, 1.0)), ib.QuadGKB200(), engine=eng)
