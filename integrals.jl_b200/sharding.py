"""Multi-GPU execution of batched problems: one process per GPU, problems split across ranks.

Batches of independent integrals (QuadGK / HCubature / fixed-rule batches) need no data-path
collective: rank r solves the contiguous block shard_range(nprob, r, world) on its own GPU and the
blocks are concatenated with one all_gather of the results (control plane; torch.distributed over
NCCL on the GPUs, gloo in the CPU tests).  The reference has no counterpart: a sweep
over p is a serial loop of solve calls (docs/src/tutorials/caching_interface.md:32-47).
"""
import numpy as np


def shard_range(nprob, rank, world):
    """[lo, hi) of the problems rank `rank` owns: contiguous blocks whose sizes differ by at most 1."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(int(nprob), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_indices(nprob, rank, world, interleave=False):
    """problem indices of rank `rank`: a contiguous block (default), or every world-th problem starting at `rank`.
    Interleaving balances batches whose cost varies smoothly with the index (BASELINE config 1: the cost of
    int sin(x p) grows with p, so contiguous blocks leave the last rank with the expensive end of the sweep)."""
    if interleave:
        if world < 1 or not 0 <= rank < world:
            raise ValueError("bad rank/world")
        return np.arange(rank, int(nprob), world)
    lo, hi = shard_range(nprob, rank, world)
    return np.arange(lo, hi)


def _is_vector(prob):
    from .integrands import VectorOf
    return isinstance(prob.f, VectorOf)


def _slice_problem(prob, idx):
    """the sub-batch `idx`: the last axis of p is the problem axis (nparam x nprob; VectorOf: nparam x ncomp x nprob)"""
    lb, ub = prob.domain
    p = np.asarray(prob.p, dtype=np.float64)
    if p.ndim < (3 if _is_vector(prob) else 2):
        raise ValueError("solve_sharded needs a batch: p must be nparam x nprob (VectorOf: nparam x ncomp x nprob)")
    lb = np.asarray(lb, dtype=np.float64); ub = np.asarray(ub, dtype=np.float64)
    if lb.ndim == 2:
        lb, ub = lb[:, idx], ub[:, idx]
    return prob.remake(p=np.asfortranarray(p[..., idx]), domain=(lb, ub))


def solve_sharded(prob, alg, group=None, local_solve=None, interleave=False, **kws):
    """solve(prob, alg; kws...) with the problems of prob.p (its last axis) split across the ranks of `group` (contiguous blocks,
    or round-robin with interleave=True).  Every rank returns the full solution (u, resid, stats in problem order; for VectorOf
    problems u is ncomp x nprob).  `local_solve(prob_block, alg, **kws)` defaults to interface.solve on this rank's GPU."""
    import torch
    import torch.distributed as dist
    from . import interface
    if local_solve is None:
        local_solve = interface.solve
    if not dist.is_initialized():
        return local_solve(prob, alg, **kws)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    nprob = np.asarray(prob.p).shape[-1]
    idx = shard_indices(nprob, rank, world, interleave)
    sol = local_solve(_slice_problem(prob, idx), alg, **kws)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    # pack [u (ncomp rows), resid, nevals, status] per problem into one padded buffer -> one all_gather
    n = len(idx)
    u = np.asarray(sol.u, dtype=np.float64).reshape(-1, n) if n else np.zeros((np.asarray(prob.p).shape[1] if _is_vector(prob) else 1, 0))
    nu = u.shape[0]
    maxn = max(len(shard_indices(nprob, r, world, interleave)) for r in range(world))
    pack = np.zeros((nu + 3, maxn))
    pack[:nu, :n] = u
    pack[nu, :n] = np.nan if sol.resid is None else np.atleast_1d(sol.resid)
    if sol.stats is not None:
        pack[nu + 1, :n] = sol.stats["nevals"]; pack[nu + 2, :n] = sol.stats["status"]
    mine = torch.from_numpy(pack).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    full = np.zeros((nu + 3, nprob))
    for r, t in enumerate(parts):
        ir = shard_indices(nprob, r, world, interleave)
        full[:, ir] = t.cpu().numpy()[:, :len(ir)]
    resid = None if sol.resid is None else full[nu]
    stats = None if sol.stats is None else dict(nevals=full[nu + 1].astype(np.int64), status=full[nu + 2].astype(np.int32))
    uo = np.asfortranarray(full[:nu]) if _is_vector(prob) else full[0]
    return interface.IntegralSolution(uo, resid, prob, sol.alg, sol.retcode, None, stats)


def sampled_shard_axis(shape, dim):
    """the axis a sampled problem is split along: the longest axis other than the integration axis `dim` (ties -> the last,
    i.e. the slowest-varying one in the column-major layout, so every rank's slab is one contiguous block)"""
    axes = [k for k in range(len(shape)) if k != dim]
    if not axes:
        raise ValueError("a 1-D sampled problem has nothing to shard")
    return max(axes, key=lambda k: (shape[k], k))


def _engine_slab(engine):
    """this rank's slab through ib200_sampled_slab (no library communicator: the partials are summed by torch.distributed)"""
    def local_slab(prob, alg, lo, hi):
        from . import interface, _lib
        eng = engine or _lib.default_engine()
        y = np.asarray(prob.y, dtype=np.float64)
        sl = [slice(None)] * y.ndim
        sl[prob.dim] = slice(lo, hi)
        ys = np.asfortranarray(y[tuple(sl)])
        shape = y.shape
        inner = int(np.prod(shape[:prob.dim], dtype=np.int64)); outer = int(np.prod(shape[prob.dim + 1:], dtype=np.int64))
        x = prob.x
        h, xs = (x.step, None) if isinstance(x, interface.Range) else (0.0, np.ascontiguousarray(x, dtype=np.float64))
        out = eng.sampled_slab(alg.rule, ys, inner, hi - lo, outer, shape[prob.dim], lo, h=h, x=xs)[:inner * outer]
        return out.reshape(shape[:prob.dim] + shape[prob.dim + 1:], order="F")
    return local_slab


def solve_sampled_split_axis(prob, alg, group=None, local_slab=None, engine=None):
    """solve(SampledIntegralProblem(y, x; dim), alg) with the INTEGRATION axis cut into one contiguous slab per rank (SURVEY 8(e) row 2,
    the preferred split for BASELINE config 2: every rank streams one contiguous block; it is also the only way to spread a single long
    vector, src/sampled.jl:51-64).  Rank r integrates the points shard_range(n, r, world) with the weights of the WHOLE grid
    (ib200_sampled_slab: end corrections only at the global ends) and the partial results -- 32 KiB at config 2 -- are summed by ONE
    all-reduce.  `local_slab(prob, alg, lo, hi)` -> this rank's contribution (default: the engine on this rank's GPU).  Every rank
    returns the whole solution.  The sum over the slabs is formed in another order than the unsharded solve's: results agree to
    rounding (<= 1e-12 sum|w y|), not bit for bit."""
    import torch
    import torch.distributed as dist
    from . import interface
    if local_slab is None:
        local_slab = _engine_slab(engine)
    y = np.asarray(prob.y)
    n = y.shape[prob.dim]
    if not dist.is_initialized():
        u = np.asarray(local_slab(prob, alg, 0, n))
        return interface.IntegralSolution(u if u.ndim else float(u), None, prob, alg, interface.ReturnCode.Success)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(n, rank, world)
    part = np.array(local_slab(prob, alg, lo, hi), dtype=np.float64, order="C")        # (ascontiguousarray would turn a scalar into shape (1,))
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.from_numpy(part.reshape(-1).copy()).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)          # the one collective of the solve
    u = t.cpu().numpy().reshape(part.shape)
    return interface.IntegralSolution(np.asfortranarray(u) if u.ndim else float(u), None, prob, alg, interface.ReturnCode.Success)


def solve_sampled_sharded(prob, alg, group=None, local_solve=None):
    """solve(SampledIntegralProblem(y, x; dim), alg) with the independent rows of `y` split across the ranks of `group`
    (BASELINE config 2; SURVEY 8(e): the outputs are independent, so there is NO data-path collective -- every rank streams its
    slab of `y` once; the per-rank result vectors are concatenated by one all_gather, control plane).  `prob.y` is the whole
    host array on every rank (each rank uploads only its slab) -- or use shard_range directly when the data is generated on
    the devices.  Every rank returns the full solution."""
    import torch
    import torch.distributed as dist
    from . import interface
    if local_solve is None:
        local_solve = interface.solve
    if not dist.is_initialized():
        return local_solve(prob, alg)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    y = np.asarray(prob.y)
    ax = sampled_shard_axis(y.shape, prob.dim)
    n = y.shape[ax]
    lo, hi = shard_range(n, rank, world)
    sl = [slice(None)] * y.ndim
    sl[ax] = slice(lo, hi)
    sol = local_solve(interface.SampledIntegralProblem(y[tuple(sl)], prob.x, prob.dim), alg)
    out_ax = ax - (1 if ax > prob.dim else 0)                       # the integration axis is gone from the result
    out_shape = tuple(v for k, v in enumerate(y.shape) if k != prob.dim)
    mine = np.moveaxis(np.asarray(sol.u, dtype=np.result_type(sol.u, np.float32)).reshape(
        tuple(hi - lo if k == out_ax else v for k, v in enumerate(out_shape))), out_ax, 0)
    maxn = max(b - a for a, b in (shard_range(n, r, world) for r in range(world)))
    pad = np.zeros((maxn,) + mine.shape[1:], dtype=mine.dtype)
    pad[:hi - lo] = mine
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    cplx = np.iscomplexobj(pad)
    t = torch.from_numpy(np.ascontiguousarray(pad.view(np.float64) if cplx else pad)).to(dev)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    blocks = []
    for r, part in enumerate(parts):
        a, b = shard_range(n, r, world)
        arr = part.cpu().numpy()
        if cplx:
            arr = arr.view(np.complex128)
        blocks.append(arr[:b - a])
    full = np.moveaxis(np.concatenate(blocks, axis=0), 0, out_ax)
    return interface.IntegralSolution(np.asfortranarray(full), None, prob, sol.alg, sol.retcode)
