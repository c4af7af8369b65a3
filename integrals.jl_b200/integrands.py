"""Registered device-side integrand family (markers).

Hand-written CUDA cannot execute arbitrary host closures, so `IntegralProblem.f` must be one of
these marker objects (the Julia wrapper uses callable structs with the same names).
`prob.p` carries the parameters: a vector of `nparam(dim)` numbers for one problem, or a matrix
`nparam x nprob` (one problem per column, Julia column-major) for a batch.

Each marker is also callable, `f(x, p)`, with the same definition as the device code
(csrc/family.cuh) so a caller can evaluate the integrand pointwise; `solve` never calls it --
there is no CPU solve path.
"""
import math

import numpy as np

from ._lib import FAMILY_IDS


class RegisteredIntegrand:
    name = None

    @property
    def family(self):
        return FAMILY_IDS[self.name]

    def nparam(self, dim):
        raise NotImplementedError

    def supports_dim(self, dim):
        return 1 <= dim <= 8

    def __repr__(self):
        return f"{type(self).__name__}()"


class _AU(RegisteredIntegrand):
    """Genz families: p = [a_1..a_d, u_1..u_d]."""

    def nparam(self, dim):
        return 2 * dim

    @staticmethod
    def _au(x, p):
        x = np.atleast_1d(np.asarray(x, dtype=float)); p = np.asarray(p, dtype=float)
        d = x.size
        return x, p[:d], p[d:2 * d]


class SinXP(RegisteredIntegrand):
    """sin(x*p), 1-D (README.md:30-40 of the reference)."""
    name = "sinxp"

    def nparam(self, dim):
        return 1

    def supports_dim(self, dim):
        return dim == 1

    def __call__(self, x, p):
        return math.sin(float(np.ravel(x)[0]) * float(np.ravel(p)[0]))


class GenzOscillatory(_AU):
    name = "genz_oscillatory"

    def __call__(self, x, p):
        x, a, u = self._au(x, p)
        return math.cos(2 * math.pi * u[0] + float(a @ x))


class GenzProductPeak(_AU):
    name = "genz_product_peak"

    def __call__(self, x, p):
        x, a, u = self._au(x, p)
        return 1.0 / float(np.prod(a ** -2.0 + (x - u) ** 2))


class GenzCornerPeak(_AU):
    name = "genz_corner_peak"

    def __call__(self, x, p):
        x, a, u = self._au(x, p)
        return (1.0 + float(a @ x)) ** (-(x.size + 1))


class GenzGaussian(_AU):
    name = "genz_gaussian"

    def __call__(self, x, p):
        x, a, u = self._au(x, p)
        return math.exp(-float(np.sum((a * (x - u)) ** 2)))


class GenzC0(_AU):
    name = "genz_c0"

    def __call__(self, x, p):
        x, a, u = self._au(x, p)
        return math.exp(-float(np.sum(a * np.abs(x - u))))


class GenzDiscontinuous(_AU):
    name = "genz_discontinuous"

    def __call__(self, x, p):
        x, a, u = self._au(x, p)
        if x[0] > u[0] or (x.size > 1 and x[1] > u[1]):
            return 0.0
        return math.exp(float(a @ x))


class CosProduct(RegisteredIntegrand):
    """prod_i cos(p_i x_i) (test/quadrule_tests.jl:6); p has one entry per axis."""
    name = "cosprod"

    def nparam(self, dim):
        return dim

    def __call__(self, x, p):
        x = np.atleast_1d(np.asarray(x, dtype=float)); p = np.broadcast_to(np.asarray(p, dtype=float).ravel(), x.shape)
        return float(np.prod(np.cos(p * x)))


class GaussPoly(RegisteredIntegrand):
    """c * prod_i x_i^k_i * exp(-sum_i a_i^2 (x_i-u_i)^2); p = [a, u, k, c]."""
    name = "gausspoly"

    def nparam(self, dim):
        return 3 * dim + 1

    def __call__(self, x, p):
        x = np.atleast_1d(np.asarray(x, dtype=float)); p = np.asarray(p, dtype=float); d = x.size
        a, u, k, c = p[:d], p[d:2 * d], p[2 * d:3 * d], p[3 * d]
        return float(c * np.prod(x ** k.astype(int)) * math.exp(-float(np.sum((a * (x - u)) ** 2))))


class ExpLinear(RegisteredIntegrand):
    """exp(sum_i a_i x_i)."""
    name = "explin"

    def nparam(self, dim):
        return dim

    def __call__(self, x, p):
        x = np.atleast_1d(np.asarray(x, dtype=float)); p = np.asarray(p, dtype=float).ravel()
        return math.exp(float(p @ x))


class GLTest(RegisteredIntegrand):
    """5x + sin(x) - p*exp(x), 1-D (test/gaussian_quadrature_tests.jl:48)."""
    name = "gltest"

    def nparam(self, dim):
        return 1

    def supports_dim(self, dim):
        return dim == 1

    def __call__(self, x, p):
        x = float(np.ravel(x)[0]); p = float(np.ravel(p)[0])
        return 5 * x + math.sin(x) - p * math.exp(x)


ALL_INTEGRANDS = [SinXP, GenzOscillatory, GenzProductPeak, GenzCornerPeak, GenzGaussian, GenzC0, GenzDiscontinuous,
                  CosProduct, GaussPoly, ExpLinear, GLTest]


class VectorOf(RegisteredIntegrand):
    """Vector-valued integrand built from one registered family: component c is `base` with parameter column
    p[:, c].  `prob.p` is nparam x ncomp (one problem, `sol.u` a vector of ncomp) or nparam x ncomp x nprob (a batch,
    `sol.u` ncomp x nprob).  The reference's counterpart is an array-valued integrand with an `integrand_prototype`
    (src/Integrals.jl:226-239, 316-326, 380-393): all components share nodes and subdivision, the error is a norm over them.
    Integrated by QuadGKB200 (1-D) and HCubatureB200 (1..8-D)."""

    def __init__(self, base):
        if not isinstance(base, RegisteredIntegrand) or isinstance(base, VectorOf):
            raise TypeError("VectorOf wraps one registered scalar integrand")
        self.base = base
        self.name = base.name

    def nparam(self, dim):
        return self.base.nparam(dim)

    def supports_dim(self, dim):
        return self.base.supports_dim(dim)

    def __call__(self, x, p):
        p = np.asarray(p, dtype=float)
        return np.array([self.base(x, p[:, c]) for c in range(p.shape[1])])

    def __repr__(self):
        return f"VectorOf({self.base!r})"
