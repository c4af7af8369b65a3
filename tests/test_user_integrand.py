"""Device-resident user integrands, CPU side: the CUDA source of an integrand is compiled by NVRTC for sm_100a without a GPU
(ib200_integrand_compile through the C ABI), compile errors come back with the compiler's log, and the host mirror routes a
CudaIntegralFunction to the engine.  Stands in for the reference's closure f(x, p) (/root/reference/src/Integrals.jl:327-335, 380-393)."""
import ctypes as C

import numpy as np
import pytest

import integrals_b200 as ib


def _compile(body, dim):
    L = ib.load_library()
    h = C.c_int32(-1)
    rc = L.ib200_integrand_compile(None, body.encode(), dim, C.byref(h))
    return rc, h.value, L.ib200_last_error(None).decode()


def test_compile_without_a_gpu_and_error_log():
    rc, h, msg = _compile("return exp(-p[0] * x[0] * x[0]) * cos(p[1] * x[1]);", 2)
    assert rc == 0 and h >= 0
    rc2, h2, msg2 = _compile("const double r = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]); return r < 1.0 ? 1.0 : 0.0;", 3)
    assert rc2 == 0 and h2 != h
    rc3, _, msg3 = _compile("return undefined_function(x[0]);", 1)
    assert rc3 != 0 and "undefined_function" in msg3 and "does not compile" in msg3
    rc4, _, msg4 = _compile("return 1.0;", 9)
    assert rc4 != 0 and "dim" in msg4
    L = ib.load_library()
    assert L.ib200_integrand_release(None, h) == 0 and L.ib200_integrand_release(None, h) != 0      # released once


def test_cuda_integral_function_routing():
    class E:
        def __init__(self):
            self.calls = []

        def compile_integrand(self, body, dim):
            self.calls.append(("compile", body, dim)); return 7

        def integrate_compiled(self, handle, dim, lb, ub, p, **k):
            self.calls.append(("integrate", handle, dim, lb, ub, np.asarray(p), k)); return 1.5, 1e-9, dict(status=0, nevals=17)
    e = E()
    f = ib.CudaIntegralFunction("return p[0] * x[0] + x[1];")
    prob = ib.IntegralProblem(f, (np.zeros(2), np.ones(2)), np.array([2.0]))
    sol = ib.solve(prob, ib.HCubatureB200(), engine=e, reltol=1e-6, maxiters=17 * 100)
    assert sol.u == 1.5 and sol.resid == 1e-9 and sol.retcode == ib.ReturnCode.Success
    assert e.calls[0] == ("compile", "return p[0] * x[0] + x[1];", 2)
    name, handle, dim, lb, ub, p, k = e.calls[1]
    assert (name, handle, dim) == ("integrate", 7, 2) and k["reltol"] == 1e-6 and k["max_boxes"] == 100 and k["transform"] == 0 and p.tolist() == [2.0]
    ib.solve(prob, ib.HCubatureB200(), engine=e)                       # compiled once per dimension
    assert [c[0] for c in e.calls] == ["compile", "integrate", "integrate"]
    with pytest.raises(ValueError, match="one-dimensional"):
        ib.solve(prob, ib.QuadGKB200(), engine=e)
    with pytest.raises(TypeError, match="QuadGKB200"):
        ib.solve(prob, ib.GaussLegendreB200(n=4), engine=e)
    with pytest.raises(TypeError, match="return statement"):
        ib.CudaIntegralFunction("x[0]")
