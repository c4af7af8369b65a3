"""ctypes binding of libintegrals_b200.so (the C ABI declared in include/integrals_b200.h).

There is no CPU fallback: if the shared library is missing, or no B200 is visible, creating an
Engine raises.  Arrays may be numpy arrays (host; staged by the library) or torch CUDA tensors /
anything with ``data_ptr()`` (device; used in place).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IB200_LIB", os.path.join(_HERE, "libintegrals_b200.so"))   # IB200_LIB: alternative build (kernel experiments)

FAMILY_IDS = dict(sinxp=0, genz_oscillatory=1, genz_product_peak=2, genz_corner_peak=3, genz_gaussian=4,
                  genz_c0=5, genz_discontinuous=6, cosprod=7, gausspoly=8, explin=9, gltest=10)
TRANSFORM_IDS = dict(transformation_if_inf=0, transformation_tan_inf=1, transformation_cot_inf=2)
RULE_IDS = dict(trapezoid=0, simpson=1)

ERR_NAMES = {-1: "IB200_ERR_INVALID", -2: "IB200_ERR_CUDA", -3: "IB200_ERR_OOM", -4: "IB200_ERR_UNSUPPORTED",
             -5: "IB200_ERR_NCCL"}

# every symbol include/integrals_b200.h declares (tests check the .so exports all of them)
SYMBOLS = ["ib200_version", "ib200_create", "ib200_destroy", "ib200_last_error", "ib200_set_stream",
           "ib200_synchronize", "ib200_kernel_launches", "ib200_family_nparam", "ib200_last_kernel_ms",
           "ib200_set_option", "ib200_sampled", "ib200_sampled_slab", "ib200_sampled_f32", "ib200_fixed_rule_tensor_batch", "ib200_fixed_rule_nodes_batch",
           "ib200_gauss_legendre_batch", "ib200_quadgk_batch", "ib200_quadgk_rule_batch", "ib200_quadgk_vec_batch", "ib200_hcubature_batch", "ib200_hcubature_vec_batch", "ib200_quadgk_sens_batch", "ib200_hcubature_sens_batch", "ib200_measure_fp64_peak",
           "ib200_hcubature_single", "ib200_batch_open", "ib200_batch_next", "ib200_batch_points", "ib200_batch_values",
           "ib200_batch_result", "ib200_integrand_compile", "ib200_integrand_release", "ib200_integrate_compiled", "ib200_comm_unique_id", "ib200_comm_init", "ib200_comm_destroy"]


class IB200Error(RuntimeError):
    """A non-zero return code from libintegrals_b200.so."""

    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code
        self.msg = msg


_lib = None


def load_library():
    """Load libintegrals_b200.so; raises ImportError if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C integrals.jl_b200/csrc` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, dp, i32, i64 = C.c_void_p, C.c_double, C.c_int32, C.c_int64
    L.ib200_version.restype = i32
    L.ib200_create.argtypes = [i32, C.POINTER(vp)]; L.ib200_create.restype = i32
    L.ib200_destroy.argtypes = [vp]; L.ib200_destroy.restype = None
    L.ib200_last_error.argtypes = [vp]; L.ib200_last_error.restype = C.c_char_p
    L.ib200_set_stream.argtypes = [vp, vp]; L.ib200_set_stream.restype = i32
    L.ib200_synchronize.argtypes = [vp]; L.ib200_synchronize.restype = i32
    L.ib200_kernel_launches.argtypes = [vp]; L.ib200_kernel_launches.restype = i64
    L.ib200_family_nparam.argtypes = [i32, i32]; L.ib200_family_nparam.restype = i32
    L.ib200_last_kernel_ms.argtypes = [vp]; L.ib200_last_kernel_ms.restype = dp
    L.ib200_set_option.argtypes = [vp, C.c_char_p, i64]; L.ib200_set_option.restype = i32
    L.ib200_sampled.argtypes = [vp, i32, vp, i64, i64, i64, dp, vp, vp]; L.ib200_sampled.restype = i32
    L.ib200_sampled_slab.argtypes = [vp, i32, vp, i64, i64, i64, i64, i64, dp, vp, vp]; L.ib200_sampled_slab.restype = i32
    L.ib200_fixed_rule_tensor_batch.argtypes = [vp, i32, i32, i32, vp, vp, i32, vp, i32, i64, vp, vp, i32, vp]
    L.ib200_fixed_rule_tensor_batch.restype = i32
    L.ib200_fixed_rule_nodes_batch.argtypes = [vp, i32, i32, i32, vp, vp, i32, vp, i32, i64, vp, vp, i64, vp]
    L.ib200_fixed_rule_nodes_batch.restype = i32
    L.ib200_gauss_legendre_batch.argtypes = [vp, i32, i32, vp, vp, i32, vp, i32, i64, vp, vp, i32, i32, vp]
    L.ib200_gauss_legendre_batch.restype = i32
    L.ib200_quadgk_batch.argtypes = [vp, i32, i32, vp, vp, i32, vp, i32, i64, dp, dp, i64, vp, vp, vp, vp]
    L.ib200_quadgk_batch.restype = i32
    L.ib200_quadgk_rule_batch.argtypes = [vp, i32, i32, vp, vp, i32, vp, i32, i64, dp, dp, i64, i32, vp, vp, vp, vp, vp, vp, vp]
    L.ib200_quadgk_rule_batch.restype = i32
    L.ib200_quadgk_vec_batch.argtypes = [vp, i32, i32, vp, vp, i32, vp, i32, i32, i64, dp, dp, i64, i32, vp, vp, vp, vp]
    L.ib200_quadgk_vec_batch.restype = i32
    L.ib200_sampled_f32.argtypes = [vp, i32, vp, i64, i64, i64, C.c_float, vp, vp]; L.ib200_sampled_f32.restype = i32
    L.ib200_hcubature_batch.argtypes = [vp, i32, i32, i32, vp, vp, i32, vp, i32, i64, dp, dp, i64, i32, i32, vp, vp, vp, vp]
    L.ib200_hcubature_batch.restype = i32
    L.ib200_hcubature_vec_batch.argtypes = [vp, i32, i32, i32, vp, vp, i32, vp, i32, i32, i64, dp, dp, i64, i32, i32, vp, vp, vp, vp]
    L.ib200_hcubature_vec_batch.restype = i32
    L.ib200_quadgk_sens_batch.argtypes = [vp, i32, i32, vp, vp, i32, vp, i32, i64, vp, i32, dp, dp, i64, i32, vp, vp, vp, vp]
    L.ib200_quadgk_sens_batch.restype = i32
    L.ib200_hcubature_sens_batch.argtypes = [vp, i32, i32, i32, vp, vp, i32, vp, i32, i64, vp, i32, dp, dp, i64, i32, i32, vp, vp, vp, vp]
    L.ib200_hcubature_sens_batch.restype = i32
    L.ib200_hcubature_single.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, dp, dp, i64, i64, C.POINTER(dp), C.POINTER(dp), vp]
    L.ib200_hcubature_single.restype = i32
    L.ib200_batch_open.argtypes = [vp, i32, i32, vp, vp, dp, dp, i64, i64]; L.ib200_batch_open.restype = i32
    L.ib200_batch_next.argtypes = [vp, C.POINTER(i64)]; L.ib200_batch_next.restype = i32
    L.ib200_batch_points.argtypes = [vp, vp, i64]; L.ib200_batch_points.restype = i32
    L.ib200_batch_values.argtypes = [vp, vp, i64]; L.ib200_batch_values.restype = i32
    L.ib200_batch_result.argtypes = [vp, C.POINTER(dp), C.POINTER(dp), vp]; L.ib200_batch_result.restype = i32
    L.ib200_integrand_compile.argtypes = [vp, C.c_char_p, i32, C.POINTER(i32)]; L.ib200_integrand_compile.restype = i32
    L.ib200_integrand_release.argtypes = [vp, i32]; L.ib200_integrand_release.restype = i32
    L.ib200_integrate_compiled.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, dp, dp, i64, i64, C.POINTER(dp), C.POINTER(dp), vp]
    L.ib200_integrate_compiled.restype = i32
    L.ib200_comm_unique_id.argtypes = [vp]; L.ib200_comm_unique_id.restype = i32
    L.ib200_comm_init.argtypes = [vp, i32, i32, vp]; L.ib200_comm_init.restype = i32
    L.ib200_comm_destroy.argtypes = [vp]; L.ib200_comm_destroy.restype = i32
    L.ib200_measure_fp64_peak.argtypes = [vp, C.POINTER(dp), C.POINTER(dp)]; L.ib200_measure_fp64_peak.restype = i32
    _lib = L
    return L


def _is_device(a):
    return hasattr(a, "data_ptr") and getattr(a, "is_cuda", False)


def _ptr(a):
    """(address, keepalive) for a numpy array, a torch tensor or None."""
    if a is None:
        return None, None
    if hasattr(a, "data_ptr"):
        return a.data_ptr(), a
    return a.ctypes.data, a


def _f64(a, order="F", what="array"):
    """Host arrays are made float64 column-major; device tensors are used in place and must already be float64 and
    contiguous (the kernels read raw doubles)."""
    if a is None:
        return a
    if hasattr(a, "data_ptr"):
        _dev_check(a, "float64", what)
        return a
    return np.asarray(a, dtype=np.float64, order=order)


def _dev_check(a, dtype, what="array"):
    """A tensor passed by pointer: right element type, dense memory."""
    got = str(getattr(a, "dtype", "")).replace("torch.", "")
    if got != dtype:
        raise TypeError(f"{what}: tensor arguments are passed by pointer and must be {dtype}, got {got}")
    if hasattr(a, "is_contiguous") and not a.is_contiguous():
        raise ValueError(f"{what}: tensor arguments must be contiguous")


HCUB_MODES = {"reference_order": 0, "rounds": 1, 0: 0, 1: 1}


MAXEVALS_UNBOUNDED = 2 ** 62   # the reference passes typemax(Int) (src/Integrals.jl:255,350)


class Engine:
    """One context = one GPU (one process per GPU).  Owns the library handle, a stream and the
    reusable device scratch (the cacheval analogue of QuadGK's segbuf / HCubature's buffer)."""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.ib200_create(int(device), C.byref(h))
        if rc != 0:
            raise IB200Error(rc, self.L.ib200_last_error(None).decode())
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.L.ib200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise IB200Error(rc, self.L.ib200_last_error(self.h).decode())

    # ---- misc --------------------------------------------------------------------------------
    def synchronize(self):
        self._check(self.L.ib200_synchronize(self.h))

    def set_stream(self, stream_ptr):
        """Run the engine's kernels on the caller's CUDA stream (0 = back to the context's own stream).  Device-pointer
        arguments are read on the engine's stream.  That stream is a blocking one, so producers on the legacy default stream
        (torch's default) are ordered before the solve implicitly; producers on any other stream join it with this call
        (bench.py does) or synchronise first."""
        self._check(self.L.ib200_set_stream(self.h, stream_ptr))

    def set_option(self, key, value):
        self._check(self.L.ib200_set_option(self.h, key.encode(), int(value)))

    @property
    def kernel_launches(self):
        return int(self.L.ib200_kernel_launches(self.h))

    @property
    def last_kernel_ms(self):
        return float(self.L.ib200_last_kernel_ms(self.h))

    def fp64_peak(self):
        t = C.c_double(); ms = C.c_double()
        self._check(self.L.ib200_measure_fp64_peak(self.h, C.byref(t), C.byref(ms)))
        return t.value, ms.value

    # ---- helpers -----------------------------------------------------------------------------
    def _out(self, out, n, dtype=np.float64):
        if out is None:
            out = np.empty(n, dtype=dtype)
        elif hasattr(out, "data_ptr"):
            _dev_check(out, np.dtype(dtype).name, "output")
            if out.numel() < n:
                raise ValueError(f"output tensor has {out.numel()} elements, {n} needed")
        else:
            if out.dtype != np.dtype(dtype) or out.size < n or not (out.flags.c_contiguous or out.flags.f_contiguous):
                raise ValueError(f"output array must be a dense {np.dtype(dtype).name} array of at least {n} elements")
        return out

    @staticmethod
    def _batch_shapes(params, lb, ub, dim):
        if hasattr(params, "data_ptr"):
            # torch tensor in Julia column-major convention: shape (nprob, nparam) row-major == nparam x nprob col-major
            nprob, nparam = (params.shape[0], params.shape[1]) if params.dim() == 2 else (1, params.shape[0])
        else:
            if params.ndim == 1:
                params = params.reshape(-1, 1, order="F")
            nparam, nprob = params.shape
        # bounds: `dim` numbers shared by all problems, or dim x nprob (1-D problems: a vector of nprob)
        nd = lb.dim() if hasattr(lb, "data_ptr") else lb.ndim
        nel = lb.numel() if hasattr(lb, "data_ptr") else lb.size
        nel_u = ub.numel() if hasattr(ub, "data_ptr") else ub.size
        if nel_u != nel:
            raise ValueError(f"lb and ub differ in size ({nel} vs {nel_u})")
        if nd == 2:
            # per-problem bounds: dim x nprob column-major (numpy (dim, nprob) F-order; torch (nprob, dim) row-major)
            shp = tuple(lb.shape)
            want = (nprob, dim) if hasattr(lb, "data_ptr") else (dim, nprob)
            if shp != want or tuple(ub.shape) != want:
                raise ValueError(f"per-problem bounds must be {dim} x {nprob}; got {shp}")
            bpp = 1
        elif nel == dim:
            bpp = 0
        elif dim == 1 and nel == nprob:
            bpp = 1
        else:
            raise ValueError(f"bounds must have {dim} entries or be {dim} x {nprob}; got {nel}")
        return params, nparam, nprob, bpp

    # ---- sampled -----------------------------------------------------------------------------
    def sampled_f32(self, rule, y, inner, n, outer, h=0.0, x=None, out=None):
        """Float32 samples on a Float32 grid (ib200_sampled_f32); host arrays column-major float32 or torch float32 tensors."""
        if not hasattr(y, "data_ptr"):
            y = np.asarray(y, dtype=np.float32, order="F")
        else:
            _dev_check(y, "float32", "y")
        if x is not None and not hasattr(x, "data_ptr"):
            x = np.ascontiguousarray(x, dtype=np.float32)
        elif x is not None:
            _dev_check(x, "float32", "x")
        out = self._out(out, max(inner * outer, 1), np.float32)
        yp, _k1 = _ptr(y); xp, _k2 = _ptr(x); op, _k3 = _ptr(out)
        self._check(self.L.ib200_sampled_f32(self.h, RULE_IDS[rule] if isinstance(rule, str) else rule, yp, inner, n, outer,
                                             float(h), xp, op))
        return out

    def sampled_slab(self, rule, y_slab, inner, n_slab, outer, n_total, offset, h=0.0, x=None, out=None):
        """the points offset .. offset + n_slab - 1 of a grid of n_total points (a slab of the integration axis); with a communicator the
        slabs' contributions are summed over the ranks (one all-reduce) and every rank gets the whole integral"""
        y_slab = _f64(y_slab, what="y"); x = _f64(x, what="x")
        out = self._out(out, max(inner * outer, 1))
        yp, _k1 = _ptr(y_slab); xp, _k2 = _ptr(x); op, _k3 = _ptr(out)
        self._check(self.L.ib200_sampled_slab(self.h, RULE_IDS[rule] if isinstance(rule, str) else rule, yp, inner, n_slab, outer,
                                              int(n_total), int(offset), float(h), xp, op))
        return out

    def sampled(self, rule, y, inner, n, outer, h=0.0, x=None, out=None):
        y = _f64(y); x = _f64(x)
        out = self._out(out, max(inner * outer, 1))
        yp, _k1 = _ptr(y); xp, _k2 = _ptr(x); op, _k3 = _ptr(out)
        self._check(self.L.ib200_sampled(self.h, RULE_IDS[rule] if isinstance(rule, str) else rule, yp, inner, n, outer,
                                         float(h), xp, op))
        return out

    # ---- fixed rules ---------------------------------------------------------------------------
    def fixed_rule_tensor_batch(self, family, dim, lb, ub, params, x1d, w1d, transform=0, out=None):
        lb = _f64(lb); ub = _f64(ub); params = _f64(params); x1d = _f64(x1d); w1d = _f64(w1d)
        params, nparam, nprob, bpp = self._batch_shapes(params, lb, ub, dim)
        out = self._out(out, nprob)
        a = [_ptr(v) for v in (lb, ub, params, x1d, w1d, out)]
        n1d = x1d.shape[0]
        self._check(self.L.ib200_fixed_rule_tensor_batch(self.h, family, dim, transform, a[0][0], a[1][0], bpp, a[2][0],
                                                         nparam, nprob, a[3][0], a[4][0], n1d, a[5][0]))
        return out

    def fixed_rule_nodes_batch(self, family, dim, lb, ub, params, nodes, weights, transform=0, out=None):
        lb = _f64(lb); ub = _f64(ub); params = _f64(params)
        nodes = _f64(nodes, order="C"); weights = _f64(weights)
        params, nparam, nprob, bpp = self._batch_shapes(params, lb, ub, dim)
        out = self._out(out, nprob)
        N = weights.shape[0]
        a = [_ptr(v) for v in (lb, ub, params, nodes, weights, out)]
        self._check(self.L.ib200_fixed_rule_nodes_batch(self.h, family, dim, transform, a[0][0], a[1][0], bpp, a[2][0],
                                                        nparam, nprob, a[3][0], a[4][0], N, a[5][0]))
        return out

    def gauss_legendre_batch(self, family, lb, ub, params, nodes, weights, subintervals=1, transform=0, out=None):
        lb = _f64(lb); ub = _f64(ub); params = _f64(params)
        nodes = np.ascontiguousarray(nodes, dtype=np.float64); weights = np.ascontiguousarray(weights, dtype=np.float64)
        params, nparam, nprob, bpp = self._batch_shapes(params, lb, ub, 1)
        out = self._out(out, nprob)
        a = [_ptr(v) for v in (lb, ub, params, nodes, weights, out)]
        self._check(self.L.ib200_gauss_legendre_batch(self.h, family, transform, a[0][0], a[1][0], bpp, a[2][0], nparam,
                                                      nprob, a[3][0], a[4][0], nodes.shape[0], subintervals, a[5][0]))
        return out

    # ---- adaptive ------------------------------------------------------------------------------
    def quadgk_batch(self, family, lb, ub, params, reltol=1e-8, abstol=1e-8, maxevals=MAXEVALS_UNBOUNDED, transform=0,
                     out_I=None, out_E=None, out_nevals=None, out_status=None, rule=None):
        """rule: None = the built-in (7,15) pair, or (x, w, wg) of another order in QuadGK.kronrod's layout (kronrod.kronrod(n))"""
        lb = _f64(lb); ub = _f64(ub); params = _f64(params)
        params, nparam, nprob, bpp = self._batch_shapes(params, lb, ub, 1)
        out_I = self._out(out_I, nprob); out_E = self._out(out_E, nprob)
        out_nevals = self._out(out_nevals, nprob, np.int64); out_status = self._out(out_status, nprob, np.int32)
        a = [_ptr(v) for v in (lb, ub, params, out_I, out_E, out_nevals, out_status)]
        if rule is not None:
            x, w, wg = (np.ascontiguousarray(v, dtype=np.float64) for v in rule)
            self._check(self.L.ib200_quadgk_rule_batch(self.h, family, transform, a[0][0], a[1][0], bpp, a[2][0], nparam, nprob,
                                                       float(reltol), float(abstol), int(min(maxevals, MAXEVALS_UNBOUNDED)),
                                                       x.size - 1, x.ctypes.data, w.ctypes.data, wg.ctypes.data,
                                                       a[3][0], a[4][0], a[5][0], a[6][0]))
            return out_I, out_E, out_nevals, out_status
        self._check(self.L.ib200_quadgk_batch(self.h, family, transform, a[0][0], a[1][0], bpp, a[2][0], nparam, nprob,
                                              float(reltol), float(abstol), int(min(maxevals, MAXEVALS_UNBOUNDED)),
                                              a[3][0], a[4][0], a[5][0], a[6][0]))
        return out_I, out_E, out_nevals, out_status

    def quadgk_vec_batch(self, family, lb, ub, params, reltol=1e-8, abstol=1e-8, maxevals=MAXEVALS_UNBOUNDED, transform=0, norm="2"):
        """Vector-valued QuadGK: params nparam x ncomp x nprob (column-major; nparam x ncomp = one problem).
        -> I (ncomp x nprob, column-major), E (nprob), nevals, status."""
        params = np.asfortranarray(params, dtype=np.float64)
        if params.ndim == 2:
            params = params.reshape(params.shape + (1,), order="F")
        nparam, ncomp, nprob = params.shape
        lb = np.ascontiguousarray(np.atleast_1d(lb), dtype=np.float64); ub = np.ascontiguousarray(np.atleast_1d(ub), dtype=np.float64)
        if lb.size == 1:
            bpp = 0
        elif lb.size == nprob:
            bpp = 1
        else:
            raise ValueError(f"bounds must be scalars or have {nprob} entries")
        out_I = np.empty((ncomp, nprob), order="F"); out_E = np.empty(nprob)
        out_nevals = np.empty(nprob, dtype=np.int64); out_status = np.empty(nprob, dtype=np.int32)
        self._check(self.L.ib200_quadgk_vec_batch(self.h, family, transform, lb.ctypes.data, ub.ctypes.data, bpp, params.ctypes.data, nparam,
                                                  ncomp, nprob, float(reltol), float(abstol), int(min(maxevals, MAXEVALS_UNBOUNDED)),
                                                  {"2": 0, "inf": 1}[norm], out_I.ctypes.data, out_E.ctypes.data, out_nevals.ctypes.data,
                                                  out_status.ctypes.data))
        return out_I, out_E, out_nevals, out_status

    def hcubature_batch(self, family, dim, lb, ub, params, reltol=1e-8, abstol=1e-8, maxevals=MAXEVALS_UNBOUNDED,
                        initdiv=1, transform=0, out_I=None, out_E=None, out_nevals=None, out_status=None, mode=0):
        """mode: 0 / "reference_order" (one box per step, the reference's sequence) or 1 / "rounds" (csrc/hcubature_rb.cuh, 2 <= dim <= 5)"""
        lb = _f64(lb); ub = _f64(ub); params = _f64(params)
        params, nparam, nprob, bpp = self._batch_shapes(params, lb, ub, dim)
        out_I = self._out(out_I, nprob); out_E = self._out(out_E, nprob)
        out_nevals = self._out(out_nevals, nprob, np.int64); out_status = self._out(out_status, nprob, np.int32)
        a = [_ptr(v) for v in (lb, ub, params, out_I, out_E, out_nevals, out_status)]
        self._check(self.L.ib200_hcubature_batch(self.h, family, dim, transform, a[0][0], a[1][0], bpp, a[2][0], nparam,
                                                 nprob, float(reltol), float(abstol),
                                                 int(min(maxevals, MAXEVALS_UNBOUNDED)), int(initdiv), HCUB_MODES[mode],
                                                 a[3][0], a[4][0], a[5][0], a[6][0]))
        return out_I, out_E, out_nevals, out_status

    def hcubature_vec_batch(self, family, dim, lb, ub, params, reltol=1e-8, abstol=1e-8, maxevals=MAXEVALS_UNBOUNDED, initdiv=1,
                            transform=0, norm="2"):
        """Vector-valued HCubature: params nparam x ncomp x nprob (column-major; nparam x ncomp = one problem); lb, ub of
        length dim (shared) or dim x nprob.  -> I (ncomp x nprob, column-major), E (nprob), nevals, status."""
        params = np.asfortranarray(params, dtype=np.float64)
        if params.ndim == 2:
            params = params.reshape(params.shape + (1,), order="F")
        nparam, ncomp, nprob = params.shape
        lb = np.asfortranarray(np.atleast_1d(lb), dtype=np.float64); ub = np.asfortranarray(np.atleast_1d(ub), dtype=np.float64)
        if lb.shape != ub.shape or lb.shape[0] != dim or (lb.ndim == 2 and lb.shape[1] != nprob) or lb.ndim > 2:
            raise ValueError(f"bounds must have {dim} entries or be {dim} x {nprob}")
        bpp = 1 if lb.ndim == 2 else 0
        out_I = np.empty((ncomp, nprob), order="F"); out_E = np.empty(nprob)
        out_nevals = np.empty(nprob, dtype=np.int64); out_status = np.empty(nprob, dtype=np.int32)
        self._check(self.L.ib200_hcubature_vec_batch(self.h, family, dim, transform, lb.ctypes.data, ub.ctypes.data, bpp, params.ctypes.data,
                                                     nparam, ncomp, nprob, float(reltol), float(abstol),
                                                     int(min(maxevals, MAXEVALS_UNBOUNDED)), int(initdiv), {"2": 0, "inf": 1}[norm],
                                                     out_I.ctypes.data, out_E.ctypes.data, out_nevals.ctypes.data, out_status.ctypes.data))
        return out_I, out_E, out_nevals, out_status

    def sens_batch(self, family, dim, lb, ub, params, wrt, reltol=1e-8, abstol=1e-8, maxevals=MAXEVALS_UNBOUNDED, initdiv=1, transform=0,
                   norm="value", rule="hcubature"):
        """Parameter sensitivities: I and dI/dp[wrt] of every problem as ONE vector-valued integral [f, df/dp...] (ib200_quadgk_sens_batch for
        rule = "quadgk", 1-D; ib200_hcubature_sens_batch otherwise).  params: nparam x nprob; wrt: 1..7 parameter indices (0-based);
        norm: "value" (the value alone drives the refinement, as dual numbers flowing through the reference's quadrature do), "2" or "inf"
        (a norm over value and derivatives).  -> out ((1 + len(wrt)) x nprob, column-major: row 0 = I, rows 1.. = dI/dp), E, nevals, status."""
        params = np.asfortranarray(params, dtype=np.float64)
        if params.ndim == 1:
            params = params.reshape(-1, 1, order="F")
        nparam, nprob = params.shape
        idx = np.ascontiguousarray(wrt, dtype=np.int32)
        if idx.ndim != 1 or not 1 <= idx.size <= 7:
            raise ValueError("wrt: 1 to 7 parameter indices per call")
        lb = np.asfortranarray(np.atleast_1d(lb), dtype=np.float64); ub = np.asfortranarray(np.atleast_1d(ub), dtype=np.float64)
        nk = {"2": 0, "inf": 1, "value": 2}[norm]
        out_I = np.empty((1 + idx.size, nprob), order="F"); out_E = np.empty(nprob)
        out_nevals = np.empty(nprob, dtype=np.int64); out_status = np.empty(nprob, dtype=np.int32)
        if rule == "quadgk":
            if dim != 1:
                raise ValueError("QuadGKB200 only accepts one-dimensional quadrature problems.")
            lb = lb.ravel(); ub = ub.ravel()
            if lb.size not in (1, nprob) or ub.size != lb.size:
                raise ValueError(f"bounds must be scalars or have {nprob} entries")
            self._check(self.L.ib200_quadgk_sens_batch(self.h, family, transform, lb.ctypes.data, ub.ctypes.data, int(lb.size != 1), params.ctypes.data,
                                                       nparam, nprob, idx.ctypes.data, idx.size, float(reltol), float(abstol),
                                                       int(min(maxevals, MAXEVALS_UNBOUNDED)), nk, out_I.ctypes.data, out_E.ctypes.data,
                                                       out_nevals.ctypes.data, out_status.ctypes.data))
        else:
            if lb.shape != ub.shape or lb.shape[0] != dim or (lb.ndim == 2 and lb.shape[1] != nprob) or lb.ndim > 2:
                raise ValueError(f"bounds must have {dim} entries or be {dim} x {nprob}")
            self._check(self.L.ib200_hcubature_sens_batch(self.h, family, dim, transform, lb.ctypes.data, ub.ctypes.data, int(lb.ndim == 2),
                                                          params.ctypes.data, nparam, nprob, idx.ctypes.data, idx.size, float(reltol), float(abstol),
                                                          int(min(maxevals, MAXEVALS_UNBOUNDED)), int(initdiv), nk, out_I.ctypes.data,
                                                          out_E.ctypes.data, out_nevals.ctypes.data, out_status.ctypes.data))
        return out_I, out_E, out_nevals, out_status

    def hcubature_single(self, family, dim, lb, ub, params, reltol=1e-8, abstol=1e-8, max_boxes=1 << 24, store_boxes=0,
                         transform=0):
        """One large domain, region-parallel rounds (all ranks of the communicator call this together).
        Returns (I, E, stats dict)."""
        lb = np.ascontiguousarray(lb, dtype=np.float64); ub = np.ascontiguousarray(ub, dtype=np.float64)
        params = np.ascontiguousarray(np.ravel(params), dtype=np.float64)
        I = C.c_double(); E = C.c_double(); st = np.zeros(8, dtype=np.int64)
        self._check(self.L.ib200_hcubature_single(self.h, family, dim, transform, lb.ctypes.data, ub.ctypes.data, params.ctypes.data,
                                                  params.size, float(reltol), float(abstol), int(max_boxes), int(store_boxes),
                                                  C.byref(I), C.byref(E), st.ctypes.data))
        return I.value, E.value, dict(nevals=int(st[0]), rule_applications=int(st[1]), boxes=int(st[2]), rounds=int(st[3]),
                                      status=int(st[4]), world=int(st[5]), level2_subproblems=int(st[6]), level2_unconverged=int(st[7]),
                                      note=self.L.ib200_last_error(self.h).decode() if (st[7] or st[6]) else "")

    # ---- device-resident user integrands (CUDA source compiled with NVRTC) -------------------------
    def compile_integrand(self, body, dim):
        """body of `__device__ double f(const double *x, const double *p)` in CUDA C++ -> handle (ib200_integrand_compile; needs no GPU)"""
        h = C.c_int32(-1)
        self._check(self.L.ib200_integrand_compile(self.h, body.encode(), int(dim), C.byref(h)))
        return h.value

    def release_integrand(self, handle):
        self._check(self.L.ib200_integrand_release(self.h, int(handle)))

    def integrate_compiled(self, handle, dim, lb, ub, params=(), reltol=1e-8, abstol=1e-8, max_boxes=1 << 20, store_boxes=0, transform=0):
        """One integral of a compiled integrand: the bridge's rounds with every evaluation on the device (ib200_integrate_compiled).
        -> I, E, stats dict"""
        lb = np.ascontiguousarray(np.atleast_1d(lb), dtype=np.float64); ub = np.ascontiguousarray(np.atleast_1d(ub), dtype=np.float64)
        if lb.shape != (dim,) or ub.shape != (dim,):
            raise ValueError(f"bounds must have {dim} entries")
        par = np.ascontiguousarray(np.ravel(params), dtype=np.float64)
        I = C.c_double(); E = C.c_double(); st = np.zeros(6, dtype=np.int64)
        self._check(self.L.ib200_integrate_compiled(self.h, int(handle), int(dim), int(transform), lb.ctypes.data, ub.ctypes.data,
                                                    par.ctypes.data if par.size else None, int(par.size), float(reltol), float(abstol),
                                                    int(min(max_boxes, MAXEVALS_UNBOUNDED)), int(store_boxes), C.byref(I), C.byref(E), st.ctypes.data))
        return I.value, E.value, dict(nevals=int(st[0]), rule_applications=int(st[1]), boxes=int(st[2]), rounds=int(st[3]), status=int(st[4]))

    # ---- BatchIntegralFunction bridge: caller-evaluated integrand ----------------------------------
    def batch_open(self, dim, lb, ub, reltol=1e-8, abstol=1e-8, max_boxes=1 << 20, store_boxes=0, transform=0):
        lb = np.ascontiguousarray(np.atleast_1d(lb), dtype=np.float64); ub = np.ascontiguousarray(np.atleast_1d(ub), dtype=np.float64)
        if lb.shape != (dim,) or ub.shape != (dim,):
            raise ValueError(f"bounds must have {dim} entries")
        self._check(self.L.ib200_batch_open(self.h, int(dim), int(transform), lb.ctypes.data, ub.ctypes.data, float(reltol), float(abstol),
                                            int(min(max_boxes, MAXEVALS_UNBOUNDED)), int(store_boxes)))

    def batch_next(self):
        """number of points of the next round; 0 = the solve has finished"""
        n = C.c_int64()
        self._check(self.L.ib200_batch_next(self.h, C.byref(n)))
        return n.value

    def batch_points(self, x, n):
        """fill x (dim x n column-major: numpy array of shape (n, dim) C-order / (dim, n) F-order, or a torch CUDA tensor (n, dim))"""
        xp, _k = _ptr(x)
        self._check(self.L.ib200_batch_points(self.h, xp, int(n)))

    def batch_values(self, y, n):
        yp, _k = _ptr(y)
        self._check(self.L.ib200_batch_values(self.h, yp, int(n)))

    def batch_result(self):
        I = C.c_double(); E = C.c_double(); st = np.zeros(6, dtype=np.int64)
        self._check(self.L.ib200_batch_result(self.h, C.byref(I), C.byref(E), st.ctypes.data))
        return I.value, E.value, dict(nevals=int(st[0]), rule_applications=int(st[1]), boxes=int(st[2]), rounds=int(st[3]), status=int(st[4]))

    def integrate_batch(self, f, dim, lb, ub, reltol=1e-8, abstol=1e-8, max_boxes=1 << 20, store_boxes=0, transform=0, device=False,
                        max_batch=None):
        """Adaptive integral of a caller-evaluated batch integrand.  f(pts) -> values: pts is a float64 array of shape (dim, n)
        ((n,) in 1-D), one point per column like the reference's BatchIntegralFunction; with device=True pts is a torch CUDA
        tensor of shape (n, dim) ((n,) in 1-D; the same memory) and f returns a CUDA tensor of n values, no host copies.
        max_batch (BatchIntegralFunction's max_batch): f is called on at most that many points at a time (a round is split on the
        host side).  -> I, E, stats."""
        if max_batch is not None and int(max_batch) < 1:
            raise ValueError("max_batch must be positive")
        self.batch_open(dim, lb, ub, reltol=reltol, abstol=abstol, max_boxes=max_boxes, store_boxes=store_boxes, transform=transform)
        batches = 0
        while True:
            n = self.batch_next()
            if n == 0:
                break
            if device:
                import torch
                x = torch.empty((n, dim) if dim > 1 else (n,), dtype=torch.float64, device=torch.device("cuda", self.device))
                self.batch_points(x, n)
                self.synchronize()                          # the points are written on the engine's stream
                y = f(x) if max_batch is None or n <= max_batch else torch.cat([f(x[i:i + max_batch]) for i in range(0, n, int(max_batch))])
                if not (hasattr(y, "is_cuda") and y.is_cuda and y.dtype == torch.float64 and y.numel() == n):
                    raise ValueError(f"the batch integrand must return {n} float64 values on the device")
                y = y.contiguous()
                torch.cuda.current_stream(y.device).synchronize()
            else:
                x = np.empty((dim, n) if dim > 1 else (n,), dtype=np.float64, order="F")
                self.batch_points(x, n)
                if max_batch is None or n <= max_batch:
                    y = np.ascontiguousarray(f(x), dtype=np.float64)
                else:
                    y = np.concatenate([np.ravel(np.asarray(f(x[..., i:i + max_batch]), dtype=np.float64)) for i in range(0, n, int(max_batch))])
                if y.shape != (n,):
                    raise ValueError(f"the batch integrand must return {n} values, got shape {y.shape}")
            self.batch_values(y, n)
            batches += 1
        I, E, stats = self.batch_result()
        stats["batches"] = batches
        return I, E, stats

    # ---- multi-GPU communicator (single-domain solve) ------------------------------------------
    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        rc = self.L.ib200_comm_unique_id(buf)
        if rc != 0:
            raise IB200Error(rc, self.L.ib200_last_error(None).decode())
        return buf.raw

    def comm_init(self, nranks, rank, id128):
        self._keep_id = C.create_string_buffer(bytes(id128), 128)
        self._check(self.L.ib200_comm_init(self.h, int(nranks), int(rank), self._keep_id))

    def comm_init_from_torch(self, group=None):
        """Build the library's NCCL communicator over the ranks of a torch.distributed group: rank 0 creates the
        id, torch broadcasts the 128 bytes (control plane)."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        dev = torch.device("cuda", self.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t = torch.frombuffer(bytearray(self.comm_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(t, src=0, group=group)
        self.comm_init(world, rank, bytes(t.cpu().numpy().tobytes()))

    def comm_destroy(self):
        self._check(self.L.ib200_comm_destroy(self.h))


_default_engines = {}


def default_engine(device=None):
    """Process-wide Engine for `device` (default: LOCAL_RANK, else 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _default_engines:
        _default_engines[device] = Engine(device)
    return _default_engines[device]
