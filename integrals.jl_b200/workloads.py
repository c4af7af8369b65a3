"""Synthetic inputs of the BASELINE.json configurations, shared by bench.py and the tests.

C1  QuadGK   int_{-2}^{5} sin(x p) dx, 65,536 p values, reltol 1e-10 (abstol = reference default 1e-8)
C2  sampled  trapezoid + Simpson over a 4096 x 1,048,576 fp64 array along dim 2
C3  fixed    tensor Gauss-Legendre n = 64 in 3-D, 2^20 Genz oscillatory parameter sets
C4  HCubature Genz suite (6 families), 4-D, 2^18 parameter sets per family, reltol 1e-6
C5  one 8-D Genz Gaussian with two infinite axes, reltol 1e-9 (region-partitioned; hcubature_single)
"""
import numpy as np

from ._lib import FAMILY_IDS

GENZ = ["genz_oscillatory", "genz_product_peak", "genz_corner_peak", "genz_gaussian", "genz_c0", "genz_discontinuous"]
# Genz standard difficulties (h, e): sum a_i = h / d^e
GENZ_DIFFICULTY = {"genz_oscillatory": (110.0, 1.5), "genz_product_peak": (600.0, 2.0), "genz_corner_peak": (600.0, 2.0),
                   "genz_gaussian": (100.0, 1.0), "genz_c0": (150.0, 2.0), "genz_discontinuous": (100.0, 2.0)}
SEED = 20261019

# nominal FP64 operation counts per integrand evaluation (DFMA = 2): family part only
FAMILY_FLOPS = {"sinxp": lambda d: 26, "genz_oscillatory": lambda d: 2 * d + 1 + 25, "genz_product_peak": lambda d: 15 * d,
                "genz_corner_peak": lambda d: 2 * d + 1 + 10 + d, "genz_gaussian": lambda d: 4 * d + 30,
                "genz_c0": lambda d: 3 * d + 30, "genz_discontinuous": lambda d: 2 * d + 30}
TRANSFORM_FLOPS_FINITE = 4          # per coordinate: u = lb + (1+t)*den, jac product


def genz_params(family, dim, nprob, seed=SEED, scale=1.0):
    """[a; u] per column (2*dim x nprob, column-major): u ~ U[0,1), a = r * (h/d^e) / sum r."""
    rng = np.random.default_rng(seed + FAMILY_IDS[family])
    h, e = GENZ_DIFFICULTY[family]
    u = rng.random((dim, nprob))
    r = rng.random((dim, nprob))
    a = r * (scale * h / dim ** e) / r.sum(axis=0, keepdims=True)
    return np.asfortranarray(np.concatenate([a, u], axis=0))


def gm_npoints(d):
    return 15 if d == 1 else 1 + 4 * d + 2 * d * (d - 1) + (1 << d)


def hcubature_flops_per_eval(family, d):
    """family + finite transform + Genz-Malik node generation (2d) + accumulation (1) + 30 per box."""
    return FAMILY_FLOPS[family](d) + TRANSFORM_FLOPS_FINITE * d + 2 * d + 1 + 30.0 / gm_npoints(d)


def quadgk_flops_per_eval(family="sinxp"):
    """family + finite transform + GK node generation (2) + K and G accumulation (4)."""
    return FAMILY_FLOPS[family](1) + TRANSFORM_FLOPS_FINITE + 2 + 4


def fixed_tensor_flops_per_eval(family, d):
    """family + finite transform + tensor node generation with per-axis precompute (2) + accumulation (2)."""
    return FAMILY_FLOPS[family](d) + TRANSFORM_FLOPS_FINITE * d + 2 + 2


def c1_params(nprob=65536):
    return (0.1 + 63.9 * np.arange(nprob) / (nprob - 1)).reshape(1, nprob, order="F")


C1 = dict(lb=-2.0, ub=5.0, reltol=1e-10, abstol=1e-8, nprob=65536)
C2 = dict(m=4096, n=1 << 20)
C3 = dict(family="genz_oscillatory", dim=3, n1d=64, nprob=1 << 20)
C4 = dict(dim=4, nprob_per_family=1 << 18, reltol=1e-6, abstol=1e-8, maxiters=4_000_000, initdiv=1)
C5 = dict(dim=8, reltol=1e-9, abstol=0.0, sum_a=12.5, seed=1)


def c5_problem():
    """-> lb, ub, params [a; u], exact.  8-D Genz Gaussian exp(-sum a_i^2 (x_i-u_i)^2), sum a_i = 12.5, seed 1; axes 1-2 on
    (-inf, inf) (transformation_if_inf), axes 3-8 on [0, 1].  Closed form: prod sqrt(pi)/a_i over the
    infinite axes times prod sqrt(pi)/(2 a_i) (erf(a_i (1-u_i)) + erf(a_i u_i)) over the finite ones."""
    import math
    rng = np.random.default_rng(C5["seed"])
    d = C5["dim"]
    u = rng.random(d); r = rng.random(d)
    a = r * C5["sum_a"] / r.sum()
    lb = np.array([-np.inf, -np.inf] + [0.0] * (d - 2)); ub = np.array([np.inf, np.inf] + [1.0] * (d - 2))
    exact = math.prod(math.sqrt(math.pi) / ai for ai in a[:2]) * math.prod(
        math.sqrt(math.pi) / (2 * ai) * (math.erf(ai * (1 - ui)) + math.erf(ai * ui)) for ai, ui in zip(a[2:], u[2:]))
    return lb, ub, np.concatenate([a, u]), exact
