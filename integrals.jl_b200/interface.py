"""Host-side mirror of the Integrals.jl problem / algorithm / solve interface for the B200 engine.

The reference toolchain (Julia) is absent from this image, so the caller above the C ABI is
written in Python with the reference's names, argument meaning and error behaviour; the Julia
extension a maintainer would add is in INTEGRATION.md.  What mirrors what:

  IntegralProblem(f, domain, p; kwargs...)        SciMLBase type re-exported at src/Integrals.jl:7-10
  SampledIntegralProblem(y, x; dim)               same; src/common.jl:319-343
  init / solve / solve_cache (= solve!)           src/common.jl:79-133, 216-222, 248-253, 376-410
  kwarg whitelist (reltol, abstol, maxiters)      src/Integrals.jl:142-179 (CommonKwargError)
  ChangeOfVariables + transformation_*_inf        src/algorithms_meta.jl:67-71, src/common.jl:97-103
  QuadGKB200 / HCubatureB200                      QuadGKJL / HCubatureJL, src/algorithms.jl:51-56, 108-115
  QuadratureRuleB200 / GaussLegendreB200          src/algorithms.jl:229-249, 289-298
  TrapezoidalB200 / SimpsonsB200                  src/algorithms_sampled.jl:44, 71
  BatchIntegralFunction / IntegralFunction        SciMLBase wrappers as used at src/Integrals.jl:274-335
  IntegralSolution(u, resid, retcode, chi)        SciMLBase.build_solution call sites Integrals.jl:338, 396

Differences from the reference, all forced by "CUDA cannot run host closures":
  * `prob.f` is a registered integrand marker (integrands.py; runs on the device), or a BatchIntegralFunction / IntegralFunction
    (any closure: evaluated by the caller on the batches of points the engine hands out, src/Integrals.jl:274-315);
  * batches: `p` may be a matrix nparam x nprob (one problem per column) and the bounds may be
    dim x nprob; then `sol.u` and `sol.resid` are vectors of length nprob;
  * `sol.stats` carries the per-problem evaluation counts and status codes the engine returns.
There is no CPU solve path: every solve goes through libintegrals_b200.so on a B200.
"""
import numpy as np

from . import _lib
from .integrands import RegisteredIntegrand, VectorOf

ALLOWED_KEYWORDS = ("maxiters", "abstol", "reltol")          # src/Integrals.jl:142


class CommonKwargError(TypeError):
    """Unrecognized keyword arguments passed to `solve` (src/Integrals.jl:163-179)."""

    def __init__(self, kwargs):
        self.kwargs = dict(kwargs)
        bad = [k for k in self.kwargs if k not in ALLOWED_KEYWORDS]
        super().__init__("Unrecognized keyword arguments found.\nThe only allowed keyword arguments to `solve` are:\n"
                         f"{ALLOWED_KEYWORDS}\nUnrecognized keyword arguments: {bad}")


def checkkwargs(kwargs):
    if any(k not in ALLOWED_KEYWORDS for k in kwargs):
        raise CommonKwargError(kwargs)


class ReturnCode:
    Success = "Success"
    MaxIters = "MaxIters"
    Failure = "Failure"


class DomainError(ValueError):
    """The integrand produced a non-finite value: QuadGK throws DomainError where a segment's error estimate is NaN / Inf."""


def retcode_of(inner, status, nevals, maxiters, slack=0):
    """Per-problem engine status -> retcode of the solve.  The reference always reports Success (src/Integrals.jl:338,396) because
    QuadGK / HCubature only return at the tolerance or at the caller's maxiters.  The engine can also stop because its box / segment
    store is full (status 1 with fewer evaluations than maxiters): that is not a successful solve under the reference's contract and is
    reported as MaxIters.  A non-finite estimate raises for the 1-D solver (QuadGK's DomainError) and is returned as it is for
    HCubature, which stops and returns as well.  slack: budgets counted in whole rule applications."""
    st = np.atleast_1d(np.asarray(status.cpu() if hasattr(status, "cpu") else status))
    ne = np.atleast_1d(np.asarray(nevals.cpu() if hasattr(nevals, "cpu") else nevals))
    if isinstance(inner, QuadGKB200) and np.any(st == 2):
        raise DomainError(f"integrand produced a non-finite value (problem {int(np.flatnonzero(st == 2)[0])})")
    capacity = np.any((st == 1) & (ne + slack < min(int(maxiters), _lib.MAXEVALS_UNBOUNDED)))
    return ReturnCode.MaxIters if capacity else ReturnCode.Success


# ---------------------------------------------------------------------------------------------
# transformations (markers; the formulas run on the device, csrc/family.cuh t2ujac)
# ---------------------------------------------------------------------------------------------
class _Transformation:
    def __init__(self, name):
        self.name = name
        self.id = _lib.TRANSFORM_IDS[name]

    def __repr__(self):
        return self.name


transformation_if_inf = _Transformation("transformation_if_inf")     # src/infinity_handling.jl:165-170
transformation_tan_inf = _Transformation("transformation_tan_inf")   # :277-282
transformation_cot_inf = _Transformation("transformation_cot_inf")   # :405-410


# ---------------------------------------------------------------------------------------------
# problems
# ---------------------------------------------------------------------------------------------
class BatchIntegralFunction:
    """BatchIntegralFunction(f) (SciMLBase; used at src/Integrals.jl:274-315): f(pts, p) evaluates the integrand at a whole batch of
    points -- pts has shape (dim, n), one point per column ((n,) in 1-D) -- and returns n values.  This is the one way to
    integrate an arbitrary (host or device) closure with the B200 engine: the engine owns the points and the adaptive
    refinement on the device (ib200_batch_*), the caller evaluates.  device=True: pts is a torch CUDA tensor of shape (n, dim)
    ((n,) in 1-D; the same memory as the column-major dim x n matrix) and f returns a CUDA tensor -- no host copies.
    max_batch bounds the points per call of f (a round's batch is split on the host side)."""

    def __init__(self, f, device=False, max_batch=None, inplace=False):
        """max_batch: largest number of points per call of f (the reference's `max_batch`, default unbounded).
        inplace=True: the reference's in-place style f(y, pts, p) (src/Integrals.jl:277-297) -- y is a preallocated vector of n values
        (numpy, or a CUDA tensor with device=True) that f fills."""
        if not callable(f):
            raise TypeError("BatchIntegralFunction wraps a callable f(pts, p)")
        if max_batch is not None and int(max_batch) < 1:
            raise ValueError("max_batch must be positive")
        self.f, self.device, self.max_batch = f, bool(device), None if max_batch is None else int(max_batch)
        self.inplace = bool(inplace)

    def __call__(self, pts, p):
        if not self.inplace:
            return self.f(pts, p)
        n = pts.shape[0] if self.device else pts.shape[-1]
        if self.device:
            import torch
            y = torch.empty(n, dtype=torch.float64, device=pts.device)
        else:
            y = np.empty(n)
        self.f(y, pts, p)
        return y


class CudaIntegralFunction:
    """An integrand stated as CUDA C++ source: the body of `__device__ double f(const double *x, const double *p)`, e.g.
    CudaIntegralFunction("return exp(-p[0] * x[0] * x[0]) * cos(p[1] * x[1]);").  The counterpart of the reference's closure f(x, p)
    (src/Integrals.jl:327-335, 380-393) for integrands outside the registered families: compiled at run time (NVRTC) into one elementwise
    kernel that the engine runs between the two kernels of the BatchIntegralFunction bridge, so every evaluation of a solve stays on the
    device (ib200_integrand_compile / ib200_integrate_compiled).  One problem per solve; integrated by QuadGKB200 (1-D) or HCubatureB200."""

    def __init__(self, body):
        if not isinstance(body, str) or "return" not in body:
            raise TypeError("CudaIntegralFunction takes the body of  double f(const double *x, const double *p)  as a string (with a return statement)")
        self.body = body
        self._handles = {}          # dim -> handle (compiled on first use)

    def handle(self, engine, dim):
        if dim not in self._handles:
            self._handles[dim] = engine.compile_integrand(self.body, dim)
        return self._handles[dim]


class IntegralFunction(BatchIntegralFunction):
    """IntegralFunction(f) (SciMLBase; the reference's default wrapper of a closure f(x, p), src/Integrals.jl:327-335, 366-378): a
    POINTWISE closure.  It cannot run on the device, so it is evaluated where the reference evaluates it -- in the caller's process --
    one point at a time over the batches of points the engine hands out (the BatchIntegralFunction bridge); points, weighted sums,
    error estimates and the refinement stay on the GPU.  x is a float in 1-D, else a vector of length dim.  An explicit opt-in: plain
    callables are still rejected by IntegralProblem, because a per-point host loop is orders of magnitude slower than a registered
    integrand or a vectorised BatchIntegralFunction."""

    def __init__(self, f, max_batch=None):
        if not callable(f):
            raise TypeError("IntegralFunction wraps a callable f(x, p)")
        self.point_f = f
        BatchIntegralFunction.__init__(self, self._batch, device=False, max_batch=max_batch)

    def _batch(self, pts, p):
        f = self.point_f
        if pts.ndim == 1:
            return np.array([f(float(x), p) for x in pts], dtype=np.float64)
        return np.array([f(pts[:, j], p) for j in range(pts.shape[1])], dtype=np.float64)


class IntegralProblem:
    """IntegralProblem(f, domain, p=None, **kwargs).  domain = (lb, ub): scalars (1-D), vectors of
    length dim, or dim x nprob arrays (per-problem bounds).  p: nparam numbers or nparam x nprob."""

    def __init__(self, f, domain, p=None, **kwargs):
        if not isinstance(f, (RegisteredIntegrand, BatchIntegralFunction, CudaIntegralFunction)):
            raise TypeError("the B200 engine integrates registered device-side integrands only "
                            "(integrals_b200.integrands), a CudaIntegralFunction (CUDA source compiled at run time) or a "
                            "BatchIntegralFunction (caller-evaluated batches of points); "
                            "arbitrary pointwise host closures cannot run on the device")
        if not (isinstance(domain, (tuple, list)) and len(domain) == 2):
            raise TypeError("domain must be a tuple (lb, ub)")
        self.f = f
        self.domain = (domain[0], domain[1])
        self.p = p
        self.kwargs = dict(kwargs)

    def remake(self, **changes):
        new = IntegralProblem(self.f, self.domain, self.p, **self.kwargs)
        for k, v in changes.items():
            setattr(new, k, v)
        return new


class Range:
    """Uniform grid, the analogue of Julia's `range(start, stop, length=n)`: selects the uniform
    weight types (src/trapezoidal.jl:43, src/simpsons.jl:85)."""

    def __init__(self, start, stop, length, dtype=np.float64):
        """dtype=np.float32 is Julia's `range(0f0, 1f0, length=n)`: a Float32 grid (Float32 weights)."""
        self.dtype = np.dtype(dtype)
        self.start, self.stop, self.length = float(start), float(stop), int(length)
        self.step = (self.stop - self.start) / (self.length - 1) if self.length > 1 else 0.0
        if self.dtype == np.float32:
            self.step = float(np.float32(np.float32(self.stop) - np.float32(self.start)) / np.float32(self.length - 1)) if self.length > 1 else 0.0

    def __len__(self):
        return self.length

    def toarray(self):
        return np.linspace(self.start, self.stop, self.length).astype(self.dtype)


class SampledIntegralProblem:
    """SampledIntegralProblem(y, x, dim=-1): integrate samples y along axis `dim` (0-based; default the
    last axis, like the reference's `dim = ndims(y)`).  y: numpy array or torch CUDA tensor."""

    def __init__(self, y, x, dim=-1):
        nd = y.dim() if hasattr(y, "data_ptr") else np.ndim(y)
        if not -nd <= dim < nd:
            raise ValueError("dim out of range")
        self.y, self.x, self.dim = y, x, dim % nd
        if len(x) != y.shape[self.dim]:
            raise ValueError("The integrand y must have the same length as the sampling points x along the integrated dimension.")


class IntegralSolution:
    def __init__(self, u, resid, prob, alg, retcode=ReturnCode.Success, chi=None, stats=None, du_dp=None):
        self.u, self.resid, self.prob, self.alg, self.retcode, self.chi, self.stats = u, resid, prob, alg, retcode, chi, stats
        self.du_dp = du_dp          # parameter sensitivities (sensealg = ParameterSensitivities): len(wrt) [x nprob]

    def __repr__(self):
        return f"IntegralSolution(u={self.u!r}, resid={self.resid!r}, retcode={self.retcode})"


# ---------------------------------------------------------------------------------------------
# algorithms
# ---------------------------------------------------------------------------------------------
class AbstractB200Algorithm:
    pass


class ParameterSensitivities:
    """`sensealg` for solve / init: return dI/dp next to I (`sol.du_dp`).  The reference differentiates solve() through the integrand --
    ForwardDiff duals flowing through QuadGKJL / HCubatureJL (ext/IntegralsForwardDiffExt.jl:14-50: the quadrature's norm sees the value
    only -> control = "value"), or the vector-valued integral of df/dp that the Zygote rule solves with the same algorithm
    (ext/IntegralsZygoteExt.jl:144-178 -> control = "2" / "inf": a norm over value and derivatives).  Here the derivative of a registered
    family is written out on the device (csrc/family.cuh Grad) and integrated as extra components on the nodes and subdivision of I.
    wrt: parameter indices (0-based rows of `p`); None = every parameter.  More than 7 are solved in groups of 7."""

    def __init__(self, wrt=None, control="value"):
        if control not in ("value", "2", "inf"):
            raise ValueError('ParameterSensitivities: control is "value", "2" or "inf"')
        self.wrt = None if wrt is None else [int(i) for i in wrt]
        self.control = control


class QuadGKB200(AbstractB200Algorithm):
    """Adaptive Gauss-Kronrod, drop-in for QuadGKJL (src/algorithms.jl:51-56).  order: 7 (default, built-in tables) or any
    2..30 (tables from kronrod.kronrod(order), the analogue of QuadGK.kronrod)."""

    def __init__(self, order=7, norm="2"):
        """norm: "2" (LinearAlgebra.norm, the QuadGKJL default) or "inf"; only used by vector-valued integrands (VectorOf)."""
        if not 2 <= int(order) <= 30:
            raise ValueError("QuadGKB200 supports Gauss-Kronrod orders 2..30")
        if norm not in ("2", "inf"):
            raise ValueError('QuadGKB200 supports norm = "2" or "inf"')
        self.order, self.norm = int(order), norm
        self.rule = None
        if self.order != 7:
            from .kronrod import kronrod
            self.rule = kronrod(self.order)


class HCubatureB200(AbstractB200Algorithm):
    """Adaptive Genz-Malik cubature, drop-in for HCubatureJL (src/algorithms.jl:108-115).
    single_domain=True sends ONE expensive integral (scalar problem, dim >= 2) to the region-parallel driver
    (ib200_hcubature_single; all ranks of the engine's communicator take part); `maxiters` then bounds the
    integrand evaluations summed over all ranks, as it does for HCubatureJL."""

    def __init__(self, initdiv=1, single_domain=False, store_boxes=0, norm="2", refinement="auto"):
        """norm: "2" (LinearAlgebra.norm, the HCubatureJL default, src/algorithms.jl:115) or "inf"; only used by vector-valued
        integrands (VectorOf).
        refinement (batched scalar solves): "reference_order" = one box per step, largest error first -- HCubature's own sequence of
        boxes; "rounds" = every box above a per-problem threshold is bisected in the same round (csrc/hcubature_rb.cuh, 2 <= dim <= 5):
        same rule, split axis, stopping test and closing sum, 5-20 % more evaluations, several times the throughput; "auto" = rounds
        where they are built (dim >= 2), else reference order."""
        if norm not in ("2", "inf"):
            raise ValueError('HCubatureB200 supports norm = "2" or "inf"')
        if refinement not in ("auto", "rounds", "reference_order"):
            raise ValueError('HCubatureB200 supports refinement = "auto", "rounds" or "reference_order"')
        self.norm = norm
        self.refinement = refinement
        self.initdiv = int(initdiv)
        self.single_domain = bool(single_domain)
        self.store_boxes = int(store_boxes)


class QuadratureRuleB200(AbstractB200Algorithm):
    """Fixed rule, drop-in for QuadratureRule(q; n) (src/algorithms.jl:289-298): q(n) returns
    (nodes, weights) on [-1,1]^dim; nodes of shape (N,) or (N, dim).  tensor=True: q(n) is a 1-D rule
    and the dim-fold tensor product is meant (test/quadrule_tests.jl:41-44) without materialising it."""

    def __init__(self, q, n=250, tensor=False):
        if n <= 0:
            raise ValueError("number of nodes must be positive")          # src/algorithms.jl:293-295
        self.q, self.n, self.tensor = q, int(n), bool(tensor)


class GaussLegendreB200(AbstractB200Algorithm):
    """1-D (composite) Gauss-Legendre, drop-in for GaussLegendre (src/algorithms.jl:229-249)."""

    def __init__(self, n=250, subintervals=1, nodes=None, weights=None):
        if subintervals <= 0:
            raise ValueError("Need at least 1 subinterval")
        if nodes is None or weights is None:
            nodes, weights = np.polynomial.legendre.leggauss(int(n))
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        self.weights = np.ascontiguousarray(weights, dtype=np.float64)
        self.subintervals = int(subintervals)


class AbstractSampledB200Algorithm:
    rule = None


class TrapezoidalB200(AbstractSampledB200Algorithm):
    rule = "trapezoid"


class SimpsonsB200(AbstractSampledB200Algorithm):
    rule = "simpson"


class ChangeOfVariables:
    """ChangeOfVariables(fu2gv, alg) (src/algorithms_meta.jl:67-71); fu2gv must be one of the three
    registered transformations."""

    def __init__(self, fu2gv, alg):
        if not isinstance(fu2gv, _Transformation):
            raise TypeError("the B200 engine supports transformation_if_inf, transformation_tan_inf and "
                            "transformation_cot_inf only")
        self.fu2gv, self.alg = fu2gv, alg


# ---------------------------------------------------------------------------------------------
# caches and solve
# ---------------------------------------------------------------------------------------------
def _problem_shapes(prob):
    """-> dim, lb, ub (float64, column-major), params (nparam x nprob), nprob, scalar_problem."""
    lb, ub = prob.domain
    lb = np.asarray(lb, dtype=np.float64) if not hasattr(lb, "data_ptr") else lb
    ub = np.asarray(ub, dtype=np.float64) if not hasattr(ub, "data_ptr") else ub
    p = prob.p
    if p is None:
        raise ValueError("registered integrands need parameters p")
    if hasattr(p, "data_ptr"):
        nprob = p.shape[0] if p.dim() == 2 else 1
        scalar = p.dim() < 2
    else:
        p = np.asarray(p, dtype=np.float64)
        scalar = p.ndim < 2
        p = np.asfortranarray(p.reshape(-1, 1) if scalar else p)
        nprob = p.shape[1]
    nd = lb.dim() if hasattr(lb, "data_ptr") else lb.ndim
    dim = 1 if nd == 0 else (lb.shape[1] if hasattr(lb, "data_ptr") and nd == 2 else lb.shape[0])
    if nd == 0:
        lb = np.atleast_1d(lb); ub = np.atleast_1d(ub)
    return dim, lb, ub, p, nprob, scalar


class IntegralCache:
    """init(prob, alg; kwargs...) result (src/common.jl:1-12).  `cacheval` is the engine context: its
    device scratch is reused by every solve (QuadGK segbuf / HCubature buffer analogue)."""

    def __init__(self, prob, alg, kwargs, engine, sensealg=None):
        self.f, self.domain, self.p = prob.f, prob.domain, prob.p
        self.prob_kwargs = prob.kwargs
        self.alg, self.kwargs, self.cacheval = alg, kwargs, engine
        self.sensealg = sensealg

    @property
    def lb(self):
        return self.domain[0]

    @property
    def ub(self):
        return self.domain[1]

    def solve(self):
        return solve_cache(self)


def init(prob, alg, engine=None, sensealg=None, do_inf_transformation=None, verbose=None, **kws):
    """init(prob, alg; sensealg, do_inf_transformation, verbose, kws...) (src/common.jl:79-133).  `sensealg`: a ParameterSensitivities
    makes the solve return dI/dp as well (the reference's AD back-ends, which cannot see through a device-side integrand, are otherwise
    accepted and ignored); the deprecated `do_inf_transformation` and `verbose` are accepted like the reference does and have no effect."""
    if isinstance(prob, SampledIntegralProblem):
        return _init_sampled(prob, alg, engine, **kws)
    kwargs = {**prob.kwargs, **kws}                      # problem kwargs under solve kwargs, common.jl:87
    checkkwargs(kwargs)
    if alg is None:
        raise ValueError("No integration algorithm `alg` was supplied as the second positional argument.")  # common.jl:163-177
    if not isinstance(alg, ChangeOfVariables):
        alg = ChangeOfVariables(transformation_if_inf, alg)   # always for tuple domains, common.jl:97-103
    if not isinstance(alg.alg, AbstractB200Algorithm):
        raise TypeError(f"{type(alg.alg).__name__} is not a B200 integration algorithm")
    return IntegralCache(prob, alg, kwargs, engine or _lib.default_engine(), sensealg if isinstance(sensealg, ParameterSensitivities) else None)


def solve(prob, alg=None, engine=None, **kws):
    """solve(prob, alg; reltol, abstol, maxiters) = solve!(init(prob, alg; ...)) (src/common.jl:216-222)."""
    return solve_cache(init(prob, alg, engine, **kws))


def solve_cache(cache):
    """solve!(cache) (src/common.jl:248-253, 408-410)."""
    if isinstance(cache, SampledIntegralCache):
        return _solve_sampled(cache)
    eng = cache.cacheval
    prob = IntegralProblem(cache.f, cache.domain, cache.p, **cache.prob_kwargs)     # build_problem, common.jl:255-257
    inner, tr = cache.alg.alg, cache.alg.fu2gv.id
    if isinstance(cache.f, BatchIntegralFunction):
        return _solve_batch_function(cache, prob, inner, tr)
    if isinstance(cache.f, CudaIntegralFunction):
        return _solve_cuda_function(cache, prob, inner, tr)
    if isinstance(cache.f, VectorOf):
        return _solve_vector(cache, prob, inner, tr)
    if getattr(cache, "sensealg", None) is not None:
        return _solve_sensitivities(cache, prob, inner, tr)
    dim, lb, ub, p, nprob, scalar = _problem_shapes(prob)
    fam = cache.f.family
    if not cache.f.supports_dim(dim):
        raise ValueError(f"{cache.f!r} is not defined in {dim} dimensions")
    kw = cache.kwargs
    reltol = kw.get("reltol", 1e-8)                      # defaults of src/Integrals.jl:254-255, 349-350
    abstol = kw.get("abstol", 1e-8)
    maxiters = kw.get("maxiters", _lib.MAXEVALS_UNBOUNDED)
    stats = None
    retcode = ReturnCode.Success
    if isinstance(inner, QuadGKB200):
        if dim != 1:
            raise ValueError("QuadGKB200 only accepts one-dimensional quadrature problems.")    # Integrals.jl:263-266
        extra = {} if inner.rule is None else {"rule": inner.rule}
        I, E, ne, st = eng.quadgk_batch(fam, lb, ub, p, reltol=reltol, abstol=abstol, maxevals=maxiters, transform=tr, **extra)
        stats = dict(nevals=ne, status=st)
        retcode = retcode_of(inner, st, ne, maxiters)
    elif isinstance(inner, HCubatureB200) and inner.single_domain:
        if not scalar or dim < 2:
            raise ValueError("HCubatureB200(single_domain=True) solves one problem of dimension >= 2 per call")
        if inner.initdiv != 1:
            raise ValueError("HCubatureB200(single_domain=True) does not take initdiv")
        npts = 1 + 4 * dim + 2 * dim * (dim - 1) + (1 << dim)
        max_boxes = max(1, min(int(maxiters), _lib.MAXEVALS_UNBOUNDED) // npts)
        Iv, Ev, stats = eng.hcubature_single(fam, dim, lb, ub, p[:, 0], reltol=reltol, abstol=abstol, max_boxes=max_boxes,
                                             store_boxes=inner.store_boxes, transform=tr)
        return IntegralSolution(Iv, Ev, prob, inner, retcode_of(inner, stats.get("status", 0), stats.get("nevals", 0), maxiters, slack=2 * npts), None, stats)
    elif isinstance(inner, HCubatureB200):
        rounds = inner.refinement == "rounds" or (inner.refinement == "auto" and dim >= 2)
        I, E, ne, st = eng.hcubature_batch(fam, dim, lb, ub, p, reltol=reltol, abstol=abstol, maxevals=maxiters,
                                           initdiv=inner.initdiv, transform=tr, mode="rounds" if rounds else "reference_order")
        stats = dict(nevals=ne, status=st, refinement="rounds" if rounds else "reference_order")
        retcode = retcode_of(inner, st, ne, maxiters)
    elif isinstance(inner, QuadratureRuleB200):
        nodes, weights = inner.q(inner.n)                # init_cacheval, src/quadrules.jl:21-23
        if inner.tensor:
            I = eng.fixed_rule_tensor_batch(fam, dim, lb, ub, p, nodes, weights, transform=tr)
        else:
            nodes = np.asarray(nodes, dtype=np.float64).reshape(len(weights), -1)
            if nodes.shape[1] != dim:
                raise ValueError("quadrature nodes and domain have different dimensions")
            I = eng.fixed_rule_nodes_batch(fam, dim, lb, ub, p, nodes, weights, transform=tr)
        E = None                                         # resid = nothing, quadrules.jl:50-51
    elif isinstance(inner, GaussLegendreB200):
        if dim != 1:
            raise ValueError("GaussLegendreB200 only accepts one-dimensional quadrature problems.")   # ext/…FastGaussQuadratureExt.jl:41-43
        I = eng.gauss_legendre_batch(fam, lb, ub, p, inner.nodes, inner.weights, subintervals=inner.subintervals, transform=tr)
        E = None
    else:
        raise TypeError(f"unsupported algorithm {type(inner).__name__}")
    if scalar and not hasattr(I, "data_ptr"):
        I = float(I[0]); E = None if E is None else float(E[0])
    return IntegralSolution(I, E, prob, inner, retcode, None, stats)   # Success (Integrals.jl:338,396) unless the engine's store filled up


def _solve_sensitivities(cache, prob, inner, tr):
    """I and dI/dp as one vector-valued integral (ParameterSensitivities; engine.sens_batch)"""
    sa = cache.sensealg
    dim, lb, ub, p, nprob, scalar = _problem_shapes(prob)
    if hasattr(p, "data_ptr") or hasattr(lb, "data_ptr"):
        raise TypeError("parameter sensitivities take host (numpy) parameters and bounds")
    if not cache.f.supports_dim(dim):
        raise ValueError(f"{cache.f!r} is not defined in {dim} dimensions")
    if isinstance(inner, QuadGKB200):
        if dim != 1:
            raise ValueError("QuadGKB200 only accepts one-dimensional quadrature problems.")    # Integrals.jl:263-266
        if inner.order != 7:
            raise ValueError("parameter sensitivities use the (7,15) pair")
        rule, extra = "quadgk", {}
    elif isinstance(inner, HCubatureB200) and not inner.single_domain:
        rule, extra = "hcubature", {"initdiv": inner.initdiv}
    else:
        raise TypeError("parameter sensitivities are computed by QuadGKB200 and HCubatureB200")
    kw = cache.kwargs
    maxiters = kw.get("maxiters", _lib.MAXEVALS_UNBOUNDED)
    nparam = p.shape[0]
    wrt = list(range(nparam)) if sa.wrt is None else sa.wrt
    if not wrt or any(not 0 <= i < nparam for i in wrt):
        raise ValueError(f"ParameterSensitivities: wrt indexes the {nparam} parameters of {cache.f!r}")
    I = E = ne = st = None
    dI = np.empty((len(wrt), nprob), order="F")
    for g0 in range(0, len(wrt), 7):                     # the kernels carry up to 8 components: the value + 7 derivatives per pass
        grp = wrt[g0:g0 + 7]
        out, Eg, neg, stg = cache.cacheval.sens_batch(cache.f.family, dim, lb, ub, p, grp, reltol=kw.get("reltol", 1e-8), abstol=kw.get("abstol", 1e-8),
                                                      maxevals=maxiters, transform=tr, norm=sa.control, rule=rule, **extra)
        dI[g0:g0 + len(grp)] = out[1:]
        if I is None:
            I, E, ne, st = out[0].copy(), Eg, neg, stg
        else:                                            # the worst of the groups
            E = np.maximum(E, Eg); ne = np.maximum(ne, neg); st = np.maximum(st, stg)
    retcode = retcode_of(inner, st, ne, maxiters)
    stats = dict(nevals=ne, status=st, wrt=wrt, control=sa.control)
    if scalar:
        return IntegralSolution(float(I[0]), float(E[0]), prob, inner, retcode, None, stats, du_dp=dI[:, 0].copy())
    return IntegralSolution(I, E, prob, inner, retcode, None, stats, du_dp=dI)


def _solve_batch_function(cache, prob, inner, tr):
    """BatchIntegralFunction (src/Integrals.jl:274-315): the engine hands out each refinement round's points, `f(pts, p)`
    evaluates them.  QuadGKB200 (1-D, GK(7,15) per segment) and HCubatureB200 (Genz-Malik per box) share the round driver."""
    lb, ub = prob.domain
    lb = np.atleast_1d(np.asarray(lb, dtype=np.float64)); ub = np.atleast_1d(np.asarray(ub, dtype=np.float64))
    if lb.ndim != 1 or lb.shape != ub.shape:
        raise ValueError("a BatchIntegralFunction problem has one domain: bounds of length dim")
    dim = lb.shape[0]
    if isinstance(inner, QuadGKB200):
        if dim != 1:
            raise ValueError("QuadGKB200 only accepts one-dimensional quadrature problems.")    # Integrals.jl:263-266
        if inner.order != 7:
            raise ValueError("BatchIntegralFunction solves use the (7,15) pair")
    elif isinstance(inner, HCubatureB200):
        if inner.initdiv != 1:
            raise ValueError("BatchIntegralFunction solves do not take initdiv")
    else:
        raise TypeError("a BatchIntegralFunction is integrated by QuadGKB200 (1-D) or HCubatureB200")
    kw = cache.kwargs
    npts = 15 if dim == 1 else 1 + 4 * dim + 2 * dim * (dim - 1) + (1 << dim)
    maxiters = kw.get("maxiters", _lib.MAXEVALS_UNBOUNDED)
    max_boxes = max(1, min(int(maxiters), _lib.MAXEVALS_UNBOUNDED) // npts)
    p = prob.p
    I, E, stats = cache.cacheval.integrate_batch(lambda pts: cache.f(pts, p), dim, lb, ub, reltol=kw.get("reltol", 1e-8),
                                                 abstol=kw.get("abstol", 1e-8), max_boxes=max_boxes,
                                                 store_boxes=getattr(inner, "store_boxes", 0), transform=tr, device=cache.f.device,
                                                 **({} if cache.f.max_batch is None else {"max_batch": cache.f.max_batch}))
    return IntegralSolution(I, E, prob, inner, retcode_of(inner, stats.get("status", 0), stats.get("nevals", 0), maxiters, slack=2 * npts), None, stats)


def _solve_cuda_function(cache, prob, inner, tr):
    """CudaIntegralFunction: the bridge's rounds with the compiled integrand evaluated on the device (engine.integrate_compiled)"""
    lb, ub = prob.domain
    lb = np.atleast_1d(np.asarray(lb, dtype=np.float64)); ub = np.atleast_1d(np.asarray(ub, dtype=np.float64))
    if lb.ndim != 1 or lb.shape != ub.shape:
        raise ValueError("a CudaIntegralFunction problem has one domain: bounds of length dim")
    dim = lb.shape[0]
    if isinstance(inner, QuadGKB200):
        if dim != 1:
            raise ValueError("QuadGKB200 only accepts one-dimensional quadrature problems.")    # Integrals.jl:263-266
        if inner.order != 7:
            raise ValueError("CudaIntegralFunction solves use the (7,15) pair")
    elif isinstance(inner, HCubatureB200):
        if inner.initdiv != 1:
            raise ValueError("CudaIntegralFunction solves do not take initdiv")
    else:
        raise TypeError("a CudaIntegralFunction is integrated by QuadGKB200 (1-D) or HCubatureB200")
    kw = cache.kwargs
    npts = 15 if dim == 1 else 1 + 4 * dim + 2 * dim * (dim - 1) + (1 << dim)
    maxiters = kw.get("maxiters", _lib.MAXEVALS_UNBOUNDED)
    max_boxes = max(1, min(int(maxiters), _lib.MAXEVALS_UNBOUNDED) // npts)
    p = () if prob.p is None else np.ravel(np.asarray(prob.p, dtype=np.float64))
    eng = cache.cacheval
    I, E, stats = eng.integrate_compiled(cache.f.handle(eng, dim), dim, lb, ub, p, reltol=kw.get("reltol", 1e-8), abstol=kw.get("abstol", 1e-8),
                                         max_boxes=min(max_boxes, 1 << 40), store_boxes=getattr(inner, "store_boxes", 0), transform=tr)
    return IntegralSolution(I, E, prob, inner, retcode_of(inner, stats.get("status", 0), stats.get("nevals", 0), maxiters, slack=2 * npts), None, stats)


def _solve_vector(cache, prob, inner, tr):
    """array-valued integrand: QuadGK (src/Integrals.jl:316-326) or HCubature (:380-393) with a norm over the components"""
    p = np.asarray(prob.p, dtype=np.float64)
    if p.ndim not in (2, 3):
        raise ValueError("VectorOf needs p of shape nparam x ncomp (or nparam x ncomp x nprob)")
    kw = cache.kwargs
    tol = dict(reltol=kw.get("reltol", 1e-8), abstol=kw.get("abstol", 1e-8), maxevals=kw.get("maxiters", _lib.MAXEVALS_UNBOUNDED))
    lb, ub = prob.domain
    if isinstance(inner, HCubatureB200):
        if inner.single_domain:
            raise ValueError("HCubatureB200(single_domain=True) integrates scalar integrands")
        lb = np.atleast_1d(np.asarray(lb, dtype=np.float64)); ub = np.atleast_1d(np.asarray(ub, dtype=np.float64))
        dim = lb.shape[0]
        if not cache.f.supports_dim(dim):
            raise ValueError(f"{cache.f!r} is not defined in {dim} dimensions")
        I, E, ne, st = cache.cacheval.hcubature_vec_batch(cache.f.family, dim, lb, ub, p, initdiv=inner.initdiv, transform=tr,
                                                          norm=inner.norm, **tol)
    elif isinstance(inner, QuadGKB200):
        if inner.order != 7:
            raise ValueError("vector-valued QuadGKB200 solves use the (7,15) pair")
        if np.ndim(lb) > 1 or (np.ndim(lb) == 1 and np.size(lb) != 1 and p.ndim != 3):
            raise ValueError("QuadGKB200 only accepts one-dimensional quadrature problems.")
        I, E, ne, st = cache.cacheval.quadgk_vec_batch(cache.f.family, lb, ub, p, transform=tr, norm=inner.norm, **tol)
    else:
        raise TypeError("vector-valued registered integrands (VectorOf) are integrated by QuadGKB200 or HCubatureB200")
    rc = retcode_of(inner, st, ne, tol["maxevals"])
    if p.ndim == 2:
        return IntegralSolution(I[:, 0].copy(), float(E[0]), prob, inner, rc, None, dict(nevals=ne, status=st))
    return IntegralSolution(I, E, prob, inner, rc, None, dict(nevals=ne, status=st))


# ---------------------------------------------------------------------------------------------
# sampled problems
# ---------------------------------------------------------------------------------------------
class SampledIntegralCache:
    """src/common.jl:264-281: assigning `x` marks the weights stale (isfresh)."""

    def __init__(self, prob, alg, engine):
        object.__setattr__(self, "isfresh", False)
        self.y, self.x, self.dim, self.alg, self.cacheval = prob.y, prob.x, prob.dim, alg, engine

    def __setattr__(self, name, value):
        if name == "x":
            object.__setattr__(self, "isfresh", True)
        object.__setattr__(self, name, value)

    def solve(self):
        return solve_cache(self)


def _init_sampled(prob, alg, engine=None, verbose=None, **kws):
    if kws:
        raise ValueError("There are no keyword arguments allowed to `solve`")     # common.jl:325-326
    if not isinstance(alg, AbstractSampledB200Algorithm):
        raise TypeError(f"{type(alg).__name__} is not a sampled B200 integration algorithm")
    return SampledIntegralCache(prob, alg, engine or _lib.default_engine())


def _solve_sampled(cache):
    object.__setattr__(cache, "isfresh", False)          # weights are recomputed on the device every solve
    u = _sampled_u(cache.cacheval, cache.alg.rule, cache.y, cache.x, cache.dim)
    prob = SampledIntegralProblem(cache.y, cache.x, cache.dim)
    return IntegralSolution(u, None, prob, cache.alg, ReturnCode.Success)          # sampled.jl:166-167


def _sampled_u(eng, rule, y, x, dim):
    """evalrule(data, weights, dim) on the device for host arrays (float64 / float32 / complex) or device tensors."""
    if hasattr(y, "data_ptr"):
        # torch tensor (row-major): axes after `dim` are the contiguous ones == column-major `inner`
        if not y.is_contiguous():
            raise ValueError("device samples must be a contiguous tensor")
        shape = tuple(y.shape)
        inner = int(np.prod(shape[dim + 1:], dtype=np.int64)); outer = int(np.prod(shape[:dim], dtype=np.int64))
        out_shape = shape[:dim] + shape[dim + 1:]
    else:
        y = np.asarray(y)
        xdt = x.dtype if isinstance(x, Range) else np.asarray(x).dtype
        if np.iscomplexobj(y):
            # Complex samples (test/interface_tests.jl:501-509): the weights are real, so Complex{Float64}[inner, n, outer]
            # is integrated as Float64[2*inner, n, outer] (real and imaginary parts interleaved, the complex layout)
            yr = np.empty((2,) + y.shape, dtype=np.float64, order="F")
            yr[0] = y.real; yr[1] = y.imag
            ur = np.asarray(_sampled_u(eng, rule, yr, x, dim + 1))
            u = ur[0] + 1j * ur[1]
            return complex(u) if np.ndim(u) == 0 else np.asfortranarray(u)
        f32 = y.dtype == np.float32 and xdt == np.float32      # Float32 type preservation (test/interface_tests.jl:524-539)
        y = np.asfortranarray(y, dtype=np.float32 if f32 else np.float64)
        shape = y.shape
        inner = int(np.prod(shape[:dim], dtype=np.int64)); outer = int(np.prod(shape[dim + 1:], dtype=np.int64))
        out_shape = shape[:dim] + shape[dim + 1:]
        if f32:
            n = shape[dim]
            h, xs = (x.step, None) if isinstance(x, Range) else (0.0, np.ascontiguousarray(x, dtype=np.float32))
            out = eng.sampled_f32(rule, y, inner, n, outer, h=h, x=xs)[:inner * outer]
            return out.reshape(out_shape, order="F") if out_shape else np.float32(out[0])
    n = shape[dim]
    if isinstance(x, Range):
        h, xs = x.step, None
    else:
        h, xs = 0.0, np.ascontiguousarray(x, dtype=np.float64) if not hasattr(x, "data_ptr") else x
    if hasattr(y, "data_ptr"):
        import torch
        if y.dtype == torch.float32:                      # Float32 samples on the device: the Float32 kernel (Float32 grid)
            xs32 = None if xs is None else (xs.to(torch.float32) if hasattr(xs, "data_ptr") else np.ascontiguousarray(xs, dtype=np.float32))
            out = torch.empty(max(inner * outer, 1), dtype=torch.float32, device=y.device)
            eng.sampled_f32(rule, y, inner, n, outer, h=h, x=xs32, out=out)
            return out[:inner * outer].reshape(out_shape) if out_shape else out[0]
        if y.dtype != torch.float64:
            raise TypeError(f"device samples must be float64 or float32, got {y.dtype}")
        out = torch.empty(max(inner * outer, 1), dtype=torch.float64, device=y.device)
        eng.sampled(rule, y, inner, n, outer, h=h, x=xs, out=out)
        return out[:inner * outer].reshape(out_shape) if out_shape else out[0]
    out = eng.sampled(rule, y, inner, n, outer, h=h, x=xs)[:inner * outer]
    return out.reshape(out_shape, order="F") if out_shape else float(out[0])
