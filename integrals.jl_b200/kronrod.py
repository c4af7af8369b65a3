"""Gauss-Kronrod (n, 2n+1) rules in QuadGK.jl's layout (`QuadGK.kronrod(n)`): x = the n+1 non-positive Kronrod nodes in
increasing order (x[-1] = 0), w = their Kronrod weights, wg = the Gauss weights of the embedded n-point rule (nodes x[1], x[3], ...).
Tables for n = 1..30 are precomputed with mpmath at 60 digits by scripts/make_kronrod_tables.py; the Julia binding passes
`QuadGK.kronrod(order)` straight to the C ABI instead."""
import json
import os

import numpy as np

_TAB = None


def kronrod(n):
    global _TAB
    if _TAB is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kronrod_tables.json")) as fh:
            _TAB = json.load(fh)
    if str(int(n)) not in _TAB:
        raise ValueError(f"Gauss-Kronrod tables are shipped for orders 1..30, got {n}")
    r = _TAB[str(int(n))]
    return np.array(r["x"]), np.array(r["w"]), np.array(r["wg"])
