"""ctypes binding of the C oracle (TEST INFRASTRUCTURE ONLY -- see oracle.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FAM = dict(sinxp=0, osc=1, prodpeak=2, corner=3, gauss=4, c0=5, discont=6,
           cosprod=7, gausspoly=8, explin=9, gltest=10)
TR = dict(if_inf=0, tan=1, cot=2)

_INTEGRAND = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_int, C.c_void_p)
_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)


def build(force=False):
    """Compile oracle/liboracle.so with the committed Makefile (gcc; a second or two)."""
    so = os.path.join(_HERE, "liboracle.so")
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("oracle.c", "oracle.h", "Makefile"))
    if force or not os.path.exists(so) or os.path.getmtime(so) < src_m:
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


BASELINE_FLAGS = "-O3 -march=native -ffp-contract=fast -fopenmp"


def build_baseline():
    """Compile the CPU-baseline build of the same source (-O3 -march=native; `make baseline`) ON THIS MACHINE, into a directory keyed
    by the CPU's feature flags (so a library built for another machine is never loaded), and return its path."""
    import hashlib
    try:
        flags = next(ln for ln in open("/proc/cpuinfo") if ln.startswith("flags"))
    except (OSError, StopIteration):
        flags = "unknown"
    key = hashlib.sha1(flags.encode()).hexdigest()[:12]
    so = os.path.join(_HERE, "_baseline", key, "liboracle_baseline.so")
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("oracle.c", "oracle.h", "Makefile"))
    if not os.path.exists(so) or os.path.getmtime(so) < src_m:
        subprocess.check_call(["make", "-C", _HERE, "-s", "baseline", f"BASELINE=_baseline/{key}/liboracle_baseline.so"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


def use_baseline_build():
    """Switch this module to the -O3 -march=native build (bench.py's CPU legs only; the tests keep the parity build)."""
    global _LIB
    _LIB = None
    lib(build_baseline())
    return BASELINE_FLAGS


def lib(path=None):
    global _LIB
    if _LIB is None:
        L = C.CDLL(path or build())
        L.orc_family_eval.restype = C.c_double
        L.orc_family_eval.argtypes = [C.c_int, C.c_int, _dp, _dp]
        L.orc_family_nparam.restype = C.c_int
        L.orc_trap_uniform_w.restype = C.c_double
        L.orc_trap_uniform_w.argtypes = [C.c_int64, C.c_double, C.c_int64]
        L.orc_simpson_uniform_w.restype = C.c_double
        L.orc_simpson_uniform_w.argtypes = [C.c_int64, C.c_double, C.c_int64]
        L.orc_trap_nonuniform_w.restype = C.c_double
        L.orc_trap_nonuniform_w.argtypes = [_dp, C.c_int64, C.c_int64]
        L.orc_simpson_nonuniform_w.restype = C.c_double
        L.orc_simpson_nonuniform_w.argtypes = [_dp, C.c_int64, C.c_int64]
        L.orc_sampled_weights.argtypes = [C.c_int, C.c_int64, C.c_double, _dp, _dp]
        L.orc_sampled_evalrule.argtypes = [C.c_int, _dp, C.c_int64, C.c_int64, C.c_int64, C.c_double, _dp, _dp, C.c_int]
        L.orc_evalrule.restype = C.c_double
        L.orc_evalrule.argtypes = [_INTEGRAND, C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, C.c_int64]
        L.orc_solve_quadrule.restype = C.c_double
        L.orc_solve_quadrule.argtypes = [_INTEGRAND, C.c_void_p, C.c_int, _dp, _dp, C.c_int, _dp, _dp, C.c_int64]
        L.orc_solve_quadrule_tensor.restype = C.c_double
        L.orc_solve_quadrule_tensor.argtypes = [_INTEGRAND, C.c_void_p, C.c_int, _dp, _dp, C.c_int, _dp, _dp, C.c_int]
        L.orc_solve_gauss_legendre.restype = C.c_double
        L.orc_solve_gauss_legendre.argtypes = [_INTEGRAND, C.c_void_p, C.c_double, C.c_double, C.c_int, _dp, _dp, C.c_int, C.c_int]
        L.orc_solve_quadgk.argtypes = [_INTEGRAND, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double,
                                       C.c_int64, _dp, _dp, _i64p]
        L.orc_solve_hcubature.argtypes = [_INTEGRAND, C.c_void_p, C.c_int, _dp, _dp, C.c_int, C.c_double, C.c_double,
                                          C.c_int64, C.c_int, _dp, _dp, _i64p]
        L.orc_quadgk.argtypes = [_INTEGRAND, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64,
                                 _dp, _dp, _i64p, _i64p]
        L.orc_hcubature.argtypes = [_INTEGRAND, C.c_void_p, C.c_int, _dp, _dp, C.c_double, C.c_double, C.c_int64,
                                    C.c_int, _dp, _dp, _i64p, _i64p]
        L.orc_cubrule.argtypes = [_INTEGRAND, C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, C.POINTER(C.c_int)]
        L.orc_gk15.argtypes = [_INTEGRAND, C.c_void_p, C.c_double, C.c_double, _dp, _dp]
        L.orc_u2t.argtypes = [C.c_double, C.c_double, _dp, _dp]
        L.orc_t2ujac.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, _dp, _dp]
        L.orc_quadgk_batch.argtypes = [C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int64, C.c_double,
                                       C.c_double, C.c_int64, _dp, _dp, _i64p, _i32p, C.c_int]
        L.orc_hcubature_batch.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int64,
                                          C.c_double, C.c_double, C.c_int64, C.c_int, _dp, _dp, _i64p, _i32p, C.c_int]
        L.orc_fixed_tensor_batch.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int64,
                                             _dp, _dp, C.c_int, _dp, C.c_int]
        L.orc_fixed_nodes_batch.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int64,
                                            _dp, _dp, C.c_int64, _dp, C.c_int]
        L.orc_gauss_legendre_batch.argtypes = [C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int64,
                                               _dp, _dp, C.c_int, C.c_int, _dp, C.c_int]
        L.orc_max_threads.restype = C.c_int
        L.orc_hcubature_rounds_family.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int, _dp, C.c_double, C.c_double,
                                                  C.c_int64, C.c_int, _dp, _dp, _i64p, _i64p, _i64p]
        _LIB = L
    return _LIB


def _arr(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


def _wrap(f):
    """Python callable f(x: ndarray[dim]) -> float as a C integrand (1-D: x is a float)."""
    def cb(xp, dim, _user):
        x = np.ctypeslib.as_array(xp, shape=(dim,))
        return float(f(x[0] if dim == 1 else x.copy()))
    return _INTEGRAND(cb)


def max_threads():
    return lib().orc_max_threads()


# ---- registered family ---------------------------------------------------------------------
def family_nparam(fam, dim):
    return lib().orc_family_nparam(FAM[fam] if isinstance(fam, str) else fam, dim)


def family_eval(fam, x, par):
    x = _arr(np.atleast_1d(x)); par = _arr(np.atleast_1d(par))
    return lib().orc_family_eval(FAM[fam] if isinstance(fam, str) else fam, x.size, _p(x), _p(par))


# ---- transforms ------------------------------------------------------------------------------
def u2t(lb, ub):
    a = C.c_double(); b = C.c_double()
    lib().orc_u2t(lb, ub, C.byref(a), C.byref(b))
    return a.value, b.value


def t2ujac(tr, t, lb, ub):
    a = C.c_double(); b = C.c_double()
    lib().orc_t2ujac(TR[tr] if isinstance(tr, str) else tr, t, lb, ub, C.byref(a), C.byref(b))
    return a.value, b.value


# ---- sampled ---------------------------------------------------------------------------------
def sampled_weights(rule, n, h=None, x=None):
    """rule: 'trapezoid'|'simpson'.  x None -> uniform (n, h) like an AbstractRange grid."""
    w = np.empty(n)
    xx = None if x is None else _arr(x)
    rc = lib().orc_sampled_weights(0 if rule.startswith("trap") else 1, n, 0.0 if h is None else h,
                                   None if xx is None else _p(xx), _p(w))
    if rc:
        raise ValueError(f"orc_sampled_weights rc={rc}")
    return w


def sampled_evalrule(rule, y, x=None, h=None, dim=-1, nthreads=1):
    """y: numpy array in *Julia column-major semantics* given as a Fortran-ordered array;
    integrates along axis `dim` (0-based, default last).  Returns array without that axis."""
    y = np.asfortranarray(y, dtype=np.float64)
    dim = dim % y.ndim
    n = y.shape[dim]
    inner = int(np.prod(y.shape[:dim], dtype=np.int64))
    outer = int(np.prod(y.shape[dim + 1:], dtype=np.int64))
    out_shape = y.shape[:dim] + y.shape[dim + 1:]
    out = np.empty(max(inner * outer, 1))
    xx = None if x is None else _arr(x)
    rc = lib().orc_sampled_evalrule(0 if rule.startswith("trap") else 1, y.ctypes.data_as(_dp), inner, n, outer,
                                    0.0 if h is None else h, None if xx is None else _p(xx), _p(out), nthreads)
    if rc == -3:
        raise ValueError("No points to integrate")
    if rc:
        raise ValueError(f"orc_sampled_evalrule rc={rc}")
    out = out[:inner * outer].reshape(out_shape, order="F")
    return out if out.ndim else float(out)


def sampled_evalrule_f32(rule, y, x=None, h=None, dim=-1):
    """Float32 samples on a Float32 grid, every operation in float, sequential sum (the reference's generic code).
    -> (result float32 array without axis `dim`, the float32 weights)."""
    y = np.asfortranarray(y, dtype=np.float32)
    dim = dim % y.ndim
    n = y.shape[dim]
    inner = int(np.prod(y.shape[:dim], dtype=np.int64)); outer = int(np.prod(y.shape[dim + 1:], dtype=np.int64))
    out_shape = y.shape[:dim] + y.shape[dim + 1:]
    out = np.empty(max(inner * outer, 1), dtype=np.float32); w = np.empty(n, dtype=np.float32)
    xx = None if x is None else np.ascontiguousarray(x, dtype=np.float32)
    fp = C.POINTER(C.c_float)
    L = lib()
    L.orc_sampled_evalrule_f32.argtypes = [C.c_int, fp, C.c_int64, C.c_int64, C.c_int64, C.c_float, fp, fp, fp]
    rc = L.orc_sampled_evalrule_f32(0 if rule.startswith("trap") else 1, y.ctypes.data_as(fp), inner, n, outer,
                                    np.float32(0.0 if h is None else h), None if xx is None else xx.ctypes.data_as(fp),
                                    out.ctypes.data_as(fp), w.ctypes.data_as(fp))
    if rc:
        raise ValueError(f"orc_sampled_evalrule_f32 rc={rc}")
    out = out[:inner * outer].reshape(out_shape, order="F")
    return (out if out.ndim else np.float32(out)), w


# ---- generic-callback solves (pin the oracle against the reference's own tests) --------------
def _bounds(lb, ub):
    lb = _arr(np.atleast_1d(lb)); ub = _arr(np.atleast_1d(ub))
    return lb, ub, lb.size


def solve_quadgk(f, lb, ub, tr="if_inf", reltol=1e-8, abstol=1e-8, maxiters=2**62):
    cb = _wrap(f)
    I = C.c_double(); E = C.c_double(); ne = C.c_int64()
    st = lib().orc_solve_quadgk(cb, None, lb, ub, TR[tr], reltol, abstol, maxiters, C.byref(I), C.byref(E), C.byref(ne))
    return I.value, E.value, ne.value, st


def solve_hcubature(f, lb, ub, tr="if_inf", reltol=1e-8, abstol=1e-8, maxiters=2**62, initdiv=1):
    cb = _wrap(f)
    lb, ub, d = _bounds(lb, ub)
    I = C.c_double(); E = C.c_double(); ne = C.c_int64()
    st = lib().orc_solve_hcubature(cb, None, d, _p(lb), _p(ub), TR[tr], reltol, abstol, maxiters, initdiv,
                                   C.byref(I), C.byref(E), C.byref(ne))
    return I.value, E.value, ne.value, st


def solve_quadrule(f, lb, ub, nodes, weights, tr="if_inf"):
    cb = _wrap(f)
    lb, ub, d = _bounds(lb, ub)
    nodes = _arr(nodes).reshape(-1, d); weights = _arr(weights).ravel()
    return lib().orc_solve_quadrule(cb, None, d, _p(lb), _p(ub), TR[tr], _p(nodes), _p(weights), weights.size)


def solve_quadrule_tensor(f, lb, ub, x1d, w1d, tr="if_inf"):
    cb = _wrap(f)
    lb, ub, d = _bounds(lb, ub)
    x1d = _arr(x1d); w1d = _arr(w1d)
    return lib().orc_solve_quadrule_tensor(cb, None, d, _p(lb), _p(ub), TR[tr], _p(x1d), _p(w1d), x1d.size)


def solve_gauss_legendre(f, lb, ub, nodes, weights, subintervals=1, tr="if_inf"):
    cb = _wrap(f)
    nodes = _arr(nodes); weights = _arr(weights)
    return lib().orc_solve_gauss_legendre(cb, None, lb, ub, TR[tr], _p(nodes), _p(weights), nodes.size, subintervals)


def quadgk(f, a, b, rtol=0.0, atol=0.0, maxevals=10**7):
    """Bare QuadGK.quadgk (no change of variables)."""
    cb = _wrap(f)
    I = C.c_double(); E = C.c_double(); ne = C.c_int64(); ns = C.c_int64()
    st = lib().orc_quadgk(cb, None, a, b, rtol, atol, maxevals, C.byref(I), C.byref(E), C.byref(ne), C.byref(ns))
    return I.value, E.value, ne.value, ns.value, st


def hcubature(f, a, b, rtol=0.0, atol=0.0, maxevals=2**62, initdiv=1):
    """Bare HCubature.hcubature (no change of variables)."""
    cb = _wrap(f)
    a, b, d = _bounds(a, b)
    I = C.c_double(); E = C.c_double(); ne = C.c_int64(); nb = C.c_int64()
    st = lib().orc_hcubature(cb, None, d, _p(a), _p(b), rtol, atol, maxevals, initdiv,
                             C.byref(I), C.byref(E), C.byref(ne), C.byref(nb))
    return I.value, E.value, ne.value, nb.value, st


def cubrule(f, a, b):
    cb = _wrap(f)
    a, b, d = _bounds(a, b)
    I = C.c_double(); E = C.c_double(); k = C.c_int()
    lib().orc_cubrule(cb, None, d, _p(a), _p(b), C.byref(I), C.byref(E), C.byref(k))
    return I.value, E.value, k.value


def gk15(f, a, b):
    cb = _wrap(f)
    I = C.c_double(); E = C.c_double()
    lib().orc_gk15(cb, None, a, b, C.byref(I), C.byref(E))
    return I.value, E.value


# ---- batch drivers over the registered family (CPU baseline legs) ----------------------------
def _batch_common(fam, dim, lb, ub, params):
    fam = FAM[fam] if isinstance(fam, str) else fam
    params = np.asfortranarray(params, dtype=np.float64)
    if params.ndim == 1:
        params = params.reshape(-1, 1, order="F")
    nparam, nprob = params.shape
    lb = np.asfortranarray(lb, dtype=np.float64); ub = np.asfortranarray(ub, dtype=np.float64)
    bpp = 1 if lb.ndim == 2 or (dim == 1 and lb.ndim == 1 and lb.size == nprob and nprob > 1) else 0
    return fam, params, nparam, nprob, lb, ub, bpp


def quadgk_batch(fam, lb, ub, params, tr="if_inf", reltol=1e-8, abstol=1e-8, maxiters=2**62, nthreads=1):
    fam, params, nparam, nprob, lb, ub, bpp = _batch_common(fam, 1, np.atleast_1d(lb), np.atleast_1d(ub), params)
    I = np.empty(nprob); E = np.empty(nprob); ne = np.empty(nprob, np.int64); st = np.empty(nprob, np.int32)
    rc = lib().orc_quadgk_batch(fam, _p(lb), _p(ub), bpp, TR[tr], params.ctypes.data_as(_dp), nparam, nprob,
                                reltol, abstol, maxiters, _p(I), _p(E), ne.ctypes.data_as(_i64p),
                                st.ctypes.data_as(_i32p), nthreads)
    assert rc == 0, rc
    return I, E, ne, st


def quadgk_rule_batch(fam, lb, ub, params, x, w, wg, tr="if_inf", reltol=1e-8, abstol=1e-8, maxiters=2**62, nthreads=1):
    """QuadGK with the Gauss-Kronrod rule (x, w, wg) of order len(x) - 1 (QuadGK.kronrod layout)"""
    fam, params, nparam, nprob, lb, ub, bpp = _batch_common(fam, 1, np.atleast_1d(lb), np.atleast_1d(ub), params)
    x = _arr(x); w = _arr(w); wg = _arr(wg)
    I = np.empty(nprob); E = np.empty(nprob); ne = np.empty(nprob, np.int64); st = np.empty(nprob, np.int32)
    L = lib()
    L.orc_quadgk_rule_batch.argtypes = [C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int64, C.c_double, C.c_double, C.c_int64,
                                        C.c_int, _dp, _dp, _dp, _dp, _dp, _i64p, _i32p, C.c_int]
    rc = L.orc_quadgk_rule_batch(fam, _p(lb), _p(ub), bpp, TR[tr], params.ctypes.data_as(_dp), nparam, nprob, reltol, abstol, maxiters,
                                 x.size - 1, _p(x), _p(w), _p(wg), _p(I), _p(E), ne.ctypes.data_as(_i64p), st.ctypes.data_as(_i32p), nthreads)
    assert rc == 0, rc
    return I, E, ne, st


def quadgk_vec_batch(fam, lb, ub, params, tr="if_inf", reltol=1e-8, abstol=1e-8, maxiters=2**62, norm="2", nthreads=1):
    """vector-valued QuadGK: params nparam x ncomp x nprob (Fortran order); -> I (ncomp x nprob), E, nevals, status"""
    fam = FAM[fam] if isinstance(fam, str) else fam
    params = np.asfortranarray(params, dtype=np.float64)
    if params.ndim == 2:
        params = params.reshape(params.shape + (1,), order="F")
    nparam, ncomp, nprob = params.shape
    lb = _arr(np.atleast_1d(lb)); ub = _arr(np.atleast_1d(ub))
    bpp = 1 if lb.size == nprob and nprob > 1 else 0
    I = np.empty((ncomp, nprob), order="F"); E = np.empty(nprob); ne = np.empty(nprob, np.int64); st = np.empty(nprob, np.int32)
    L = lib()
    L.orc_quadgk_vec_batch.argtypes = [C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_double,
                                       C.c_int64, C.c_int, _dp, _dp, _i64p, _i32p, C.c_int]
    rc = L.orc_quadgk_vec_batch(fam, _p(lb), _p(ub), bpp, TR[tr], params.ctypes.data_as(_dp), nparam, ncomp, nprob, reltol, abstol,
                                maxiters, 0 if norm == "2" else 1, I.ctypes.data_as(_dp), _p(E), ne.ctypes.data_as(_i64p),
                                st.ctypes.data_as(_i32p), nthreads)
    assert rc == 0, rc
    return I, E, ne, st


def hcubature_batch(fam, lb, ub, params, tr="if_inf", reltol=1e-8, abstol=1e-8, maxiters=2**62, initdiv=1, nthreads=1):
    lb = np.asarray(lb, dtype=np.float64); dim = lb.shape[0]
    fam, params, nparam, nprob, lb, ub, bpp = _batch_common(fam, dim, lb, ub, params)
    I = np.empty(nprob); E = np.empty(nprob); ne = np.empty(nprob, np.int64); st = np.empty(nprob, np.int32)
    rc = lib().orc_hcubature_batch(fam, dim, _p(lb), _p(ub), bpp, TR[tr], params.ctypes.data_as(_dp), nparam, nprob,
                                   reltol, abstol, maxiters, initdiv, _p(I), _p(E), ne.ctypes.data_as(_i64p),
                                   st.ctypes.data_as(_i32p), nthreads)
    assert rc == 0, rc
    return I, E, ne, st


def hcubature_rounds_batch(fam, lb, ub, params, tr="if_inf", reltol=1e-8, abstol=1e-8, maxiters=2**62, initdiv=1, nthreads=1, fraction=0.5):
    """Round-based refinement of a batch (checker of Engine.hcubature_batch(mode="rounds")) -> I, E, nevals, status, rounds"""
    lb = np.asarray(lb, dtype=np.float64); dim = lb.shape[0]
    fam, params, nparam, nprob, lb, ub, bpp = _batch_common(fam, dim, lb, ub, params)
    I = np.empty(nprob); E = np.empty(nprob); ne = np.empty(nprob, np.int64); st = np.empty(nprob, np.int32); nr = np.empty(nprob, np.int64)
    L = lib()
    L.orc_hcubature_rounds_batch.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int64, C.c_double, C.c_double,
                                             C.c_int64, C.c_int, _dp, _dp, _i64p, _i32p, _i64p, C.c_int]
    C.c_double.in_dll(L, "orc_round_fraction").value = fraction
    rc = L.orc_hcubature_rounds_batch(fam, dim, _p(lb), _p(ub), bpp, TR[tr], params.ctypes.data_as(_dp), nparam, nprob,
                                      reltol, abstol, min(maxiters, 2**62), initdiv, _p(I), _p(E), ne.ctypes.data_as(_i64p),
                                      st.ctypes.data_as(_i32p), nr.ctypes.data_as(_i64p), nthreads)
    assert rc == 0, rc
    return I, E, ne, st, nr


def hcubature_vec_batch(fam, lb, ub, params, tr="if_inf", reltol=1e-8, abstol=1e-8, maxiters=2**62, initdiv=1, norm="2", nthreads=1):
    """vector-valued HCubature: params nparam x ncomp x nprob (Fortran order); lb, ub of length dim (or dim x nprob)
    -> I (ncomp x nprob), E, nevals, status"""
    fam = FAM[fam] if isinstance(fam, str) else fam
    params = np.asfortranarray(params, dtype=np.float64)
    if params.ndim == 2:
        params = params.reshape(params.shape + (1,), order="F")
    nparam, ncomp, nprob = params.shape
    lb = np.asfortranarray(np.atleast_1d(lb), dtype=np.float64); ub = np.asfortranarray(np.atleast_1d(ub), dtype=np.float64)
    dim = lb.shape[0]
    bpp = 1 if lb.ndim == 2 else 0
    I = np.empty((ncomp, nprob), order="F"); E = np.empty(nprob); ne = np.empty(nprob, np.int64); st = np.empty(nprob, np.int32)
    L = lib()
    L.orc_hcubature_vec_batch.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int, C.c_int64, C.c_double,
                                          C.c_double, C.c_int64, C.c_int, C.c_int, _dp, _dp, _i64p, _i32p, C.c_int]
    rc = L.orc_hcubature_vec_batch(fam, dim, _p(lb), _p(ub), bpp, TR[tr], params.ctypes.data_as(_dp), nparam, ncomp, nprob, reltol,
                                   abstol, maxiters, initdiv, 0 if norm == "2" else 1, I.ctypes.data_as(_dp), _p(E),
                                   ne.ctypes.data_as(_i64p), st.ctypes.data_as(_i32p), nthreads)
    assert rc == 0, rc
    return I, E, ne, st


def family_grad(fam, x, params):
    """[f, df/dparams...] of a registered family at the point x (orc_family_grad); None for a family without sensitivities"""
    fam = FAM[fam] if isinstance(fam, str) else fam
    x = _arr(np.atleast_1d(x)); params = _arr(np.ravel(params))
    out = np.empty(1 + params.size)
    L = lib()
    L.orc_family_grad.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp]
    return out if L.orc_family_grad(fam, x.size, _p(x), _p(params), _p(out)) == 0 else None


def sens_batch(fam, lb, ub, params, wrt, tr="if_inf", reltol=1e-8, abstol=1e-8, maxiters=2**62, initdiv=1, norm="value", rule="hcubature", nthreads=1):
    """I and dI/dp[wrt] as one vector-valued integral [f, df/dp...] (checker of Engine.sens_batch): params nparam x nprob;
    -> out ((1 + len(wrt)) x nprob), E, nevals, status"""
    fam = FAM[fam] if isinstance(fam, str) else fam
    params = np.asfortranarray(params, dtype=np.float64)
    if params.ndim == 1:
        params = params.reshape(-1, 1, order="F")
    nparam, nprob = params.shape
    idx = np.ascontiguousarray(wrt, dtype=np.int32)
    lb = np.asfortranarray(np.atleast_1d(lb), dtype=np.float64); ub = np.asfortranarray(np.atleast_1d(ub), dtype=np.float64)
    nk = {"2": 0, "inf": 1, "value": 2}[norm]
    I = np.empty((1 + idx.size, nprob), order="F"); E = np.empty(nprob); ne = np.empty(nprob, np.int64); st = np.empty(nprob, np.int32)
    L = lib()
    ip = idx.ctypes.data_as(C.POINTER(C.c_int))
    if rule == "quadgk":
        lb = lb.ravel(); ub = ub.ravel()
        L.orc_quadgk_sens_batch.argtypes = [C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int64, C.POINTER(C.c_int), C.c_int, C.c_double,
                                            C.c_double, C.c_int64, C.c_int, _dp, _dp, _i64p, _i32p, C.c_int]
        rc = L.orc_quadgk_sens_batch(fam, _p(lb), _p(ub), int(lb.size != 1), TR[tr], params.ctypes.data_as(_dp), nparam, nprob, ip, idx.size,
                                     reltol, abstol, maxiters, nk, I.ctypes.data_as(_dp), _p(E), ne.ctypes.data_as(_i64p), st.ctypes.data_as(_i32p), nthreads)
    else:
        dim = lb.shape[0]
        L.orc_hcubature_sens_batch.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int64, C.POINTER(C.c_int), C.c_int,
                                               C.c_double, C.c_double, C.c_int64, C.c_int, C.c_int, _dp, _dp, _i64p, _i32p, C.c_int]
        rc = L.orc_hcubature_sens_batch(fam, dim, _p(lb), _p(ub), int(lb.ndim == 2), TR[tr], params.ctypes.data_as(_dp), nparam, nprob, ip, idx.size,
                                        reltol, abstol, maxiters, initdiv, nk, I.ctypes.data_as(_dp), _p(E), ne.ctypes.data_as(_i64p),
                                        st.ctypes.data_as(_i32p), nthreads)
    assert rc == 0, rc
    return I, E, ne, st


def hcubature_rounds(fam, lb, ub, params, tr="if_inf", reltol=1e-8, abstol=1e-8, max_boxes=1 << 24, m0=0):
    """Round-synchronous refinement of ONE integral (checker for Engine.hcubature_single).
    -> I, E, dict(rule_applications, boxes, rounds, status)."""
    lb = _arr(lb); ub = _arr(ub); params = _arr(np.ravel(params))
    fam = FAM[fam] if isinstance(fam, str) else fam
    I = C.c_double(); E = C.c_double(); na = C.c_int64(); nb = C.c_int64(); nr = C.c_int64()
    st = lib().orc_hcubature_rounds_family(fam, lb.size, _p(lb), _p(ub), TR[tr], _p(params), reltol, abstol, max_boxes, m0,
                                           C.byref(I), C.byref(E), C.byref(na), C.byref(nb), C.byref(nr))
    assert st >= 0, st
    return I.value, E.value, dict(rule_applications=na.value, boxes=nb.value, rounds=nr.value, status=st)


def hcubature_two_level(fam, lb, ub, params, tr="if_inf", reltol=1e-8, abstol=1e-8, max_boxes=1 << 30, switch_boxes=1 << 20, switch_factor=4096.0,
                        sub_max_boxes=1 << 40, alpha=None, sub_arena=0):
    """Two-level refinement of ONE integral (checker for Engine.hcubature_single when the level-1 store is handed over):
    -> I, E, dict(rule_applications, boxes, rounds, status, level2_subproblems, retired_boxes).
    sub_arena > 0: boxes a sub-problem's store may hold (the engine's single_sub_boxes option), with converged-box retirement."""
    lb = _arr(lb); ub = _arr(ub); params = _arr(np.ravel(params))
    fam = FAM[fam] if isinstance(fam, str) else fam
    d = lb.size
    alpha = d / (d + 6.0) if alpha is None else alpha
    I = C.c_double(); E = C.c_double(); na = C.c_int64(); nb = C.c_int64(); nr = C.c_int64(); ns = C.c_int64()
    L = lib()
    L.orc_hcubature_two_level_family.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int, _dp, C.c_double, C.c_double, C.c_int64, C.c_int64, C.c_double,
                                                 C.c_int64, C.c_double, _dp, _dp, _i64p, _i64p, _i64p, _i64p]
    C.c_int64.in_dll(L, "orc_sub_arena").value = int(sub_arena)
    try:
        st = L.orc_hcubature_two_level_family(fam, d, _p(lb), _p(ub), TR[tr], _p(params), reltol, abstol, max_boxes, switch_boxes, switch_factor,
                                              sub_max_boxes, alpha, C.byref(I), C.byref(E), C.byref(na), C.byref(nb), C.byref(nr), C.byref(ns))
    finally:
        C.c_int64.in_dll(L, "orc_sub_arena").value = 0
    assert st >= 0, st
    return I.value, E.value, dict(rule_applications=na.value, boxes=nb.value, rounds=nr.value, status=st, level2_subproblems=ns.value,
                                  retired_boxes=C.c_int64.in_dll(L, "orc_last_retired").value)


def solve_hcubature_rounds(f, lb, ub, tr="if_inf", reltol=1e-8, abstol=1e-8, max_boxes=1 << 20, m0=0):
    """Round-synchronous refinement of ONE integral of a Python callable f(x: ndarray[dim]) (1-D: x is a float): the checker of
    the BatchIntegralFunction bridge (csrc/batch_bridge.cu).  -> I, E, dict(rule_applications, boxes, rounds, status)."""
    cb = _wrap(f)
    lb, ub, d = _bounds(lb, ub)
    I = C.c_double(); E = C.c_double(); na = C.c_int64(); nb = C.c_int64(); nr = C.c_int64()
    L = lib()
    L.orc_solve_hcubature_rounds.argtypes = [_INTEGRAND, C.c_void_p, C.c_int, _dp, _dp, C.c_int, C.c_double, C.c_double, C.c_int64, C.c_int,
                                             C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                             C.POINTER(C.c_int64)]
    st = L.orc_solve_hcubature_rounds(cb, None, d, _p(lb), _p(ub), TR[tr], reltol, abstol, max_boxes, m0,
                                      C.byref(I), C.byref(E), C.byref(na), C.byref(nb), C.byref(nr))
    assert st >= 0, st
    return I.value, E.value, dict(rule_applications=na.value, boxes=nb.value, rounds=nr.value, status=st)


def fixed_tensor_batch(fam, lb, ub, params, x1d, w1d, tr="if_inf", nthreads=1):
    lb = np.asarray(lb, dtype=np.float64); dim = lb.shape[0]
    fam, params, nparam, nprob, lb, ub, bpp = _batch_common(fam, dim, lb, ub, params)
    x1d = _arr(x1d); w1d = _arr(w1d)
    I = np.empty(nprob)
    rc = lib().orc_fixed_tensor_batch(fam, dim, _p(lb), _p(ub), bpp, TR[tr], params.ctypes.data_as(_dp), nparam, nprob,
                                      _p(x1d), _p(w1d), x1d.size, _p(I), nthreads)
    assert rc == 0, rc
    return I


def fixed_nodes_batch(fam, lb, ub, params, nodes, weights, tr="if_inf", nthreads=1):
    lb = np.asarray(lb, dtype=np.float64); dim = lb.shape[0]
    fam, params, nparam, nprob, lb, ub, bpp = _batch_common(fam, dim, lb, ub, params)
    nodes = _arr(nodes).reshape(-1, dim); weights = _arr(weights).ravel()
    I = np.empty(nprob)
    rc = lib().orc_fixed_nodes_batch(fam, dim, _p(lb), _p(ub), bpp, TR[tr], params.ctypes.data_as(_dp), nparam, nprob,
                                     _p(nodes), _p(weights), weights.size, _p(I), nthreads)
    assert rc == 0, rc
    return I


def gauss_legendre_batch(fam, lb, ub, params, nodes, weights, subintervals=1, tr="if_inf", nthreads=1):
    fam, params, nparam, nprob, lb, ub, bpp = _batch_common(fam, 1, np.atleast_1d(lb), np.atleast_1d(ub), params)
    nodes = _arr(nodes); weights = _arr(weights)
    I = np.empty(nprob)
    rc = lib().orc_gauss_legendre_batch(fam, _p(lb), _p(ub), bpp, TR[tr], params.ctypes.data_as(_dp), nparam, nprob,
                                        _p(nodes), _p(weights), nodes.size, subintervals, _p(I), nthreads)
    assert rc == 0, rc
    return I
