"""Shared helpers for the GPU parity tests (the CUDA path through the C ABI vs the CPU oracle)."""
import math

import numpy as np

import integrals_b200 as ib

F = ib.FAMILY_IDS
ORC_FAM = {"sinxp": "sinxp", "genz_oscillatory": "osc", "genz_product_peak": "prodpeak", "genz_corner_peak": "corner",
           "genz_gaussian": "gauss", "genz_c0": "c0", "genz_discontinuous": "discont", "cosprod": "cosprod",
           "gausspoly": "gausspoly", "explin": "explin", "gltest": "gltest"}
from integrals_b200.workloads import GENZ, GENZ_DIFFICULTY, genz_params  # noqa: E402,F401


def genz_exact(family, par, dim):
    """closed forms over [0,1]^dim in float64 (mpmath for the corner peak)."""
    a, u = par[:dim], par[dim:]
    if family == "genz_oscillatory":
        return math.cos(2 * math.pi * u[0] + 0.5 * a.sum()) * math.prod(2 * math.sin(ai / 2) / ai for ai in a)
    if family == "genz_product_peak":
        return math.prod(ai * (math.atan(ai * (1 - ui)) + math.atan(ai * ui)) for ai, ui in zip(a, u))
    if family == "genz_gaussian":
        return math.prod(math.sqrt(math.pi) / (2 * ai) * (math.erf(ai * (1 - ui)) + math.erf(ai * ui)) for ai, ui in zip(a, u))
    if family == "genz_c0":
        return math.prod((2 - math.exp(-ai * ui) - math.exp(-ai * (1 - ui))) / ai for ai, ui in zip(a, u))
    if family == "genz_discontinuous":
        return math.prod((math.exp(ai * (ui if i < 2 else 1.0)) - 1) / ai for i, (ai, ui) in enumerate(zip(a, u)))
    if family == "genz_corner_peak":
        import mpmath as mp
        mp.mp.dps = 40
        s = mp.mpf(0)
        for v in range(1 << dim):
            bits = [(v >> i) & 1 for i in range(dim)]
            s += (-1) ** sum(bits) / (1 + sum(mp.mpf(float(a[i])) * bits[i] for i in range(dim)))
        return float(s / (mp.factorial(dim) * mp.fprod([mp.mpf(float(x)) for x in a])))
    raise KeyError(family)
