"""Multi-rank host logic on CPU: world_size-2 gloo groups.  The per-rank GPU solve is replaced by
the CPU oracle (tests may call it), so what is under test is the partitioning, the result gather and the initial
domain split of the single-domain driver -- not the kernels."""
import os
import socket

import numpy as np
import pytest

import integrals_b200 as ib
from integrals_b200 import sharding


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 65536, 2 ** 18 + 3):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[r][1] == spans[r + 1][0] for r in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_shard_indices_interleaved():
    for n in (0, 1, 7, 65536):
        for w in (1, 2, 3, 8):
            parts = [sharding.shard_indices(n, r, w, interleave=True) for r in range(w)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n)) and max(map(len, parts)) - min(map(len, parts)) <= 1
            assert np.array_equal(np.concatenate([sharding.shard_indices(n, r, w) for r in range(w)]), np.arange(n))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _oracle_solve(prob, alg, **kw):
    """stand-in for interface.solve on a rank without a GPU: same problem block through the CPU oracle"""
    import oracle as O
    from integrals_b200 import interface
    lb, ub = prob.domain
    p = np.asarray(prob.p)
    if isinstance(prob.f, ib.VectorOf):
        I, E, ne, st = O.hcubature_vec_batch("gauss", np.asarray(lb), np.asarray(ub), p, reltol=kw.get("reltol", 1e-8), abstol=kw.get("abstol", 1e-8))
    elif isinstance(alg, ib.QuadGKB200):
        I, E, ne, st = O.quadgk_batch("sinxp", lb, ub, p, reltol=kw.get("reltol", 1e-8), abstol=kw.get("abstol", 1e-8))
    else:
        I, E, ne, st = O.hcubature_batch("gauss", np.asarray(lb), np.asarray(ub), p, reltol=kw.get("reltol", 1e-8), abstol=kw.get("abstol", 1e-8))
    return interface.IntegralSolution(I, E, prob, alg, interface.ReturnCode.Success, None, dict(nevals=ne, status=st))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = (0.1 + 20.0 * np.arange(37) / 36).reshape(1, -1)
        prob = ib.IntegralProblem(ib.SinXP(), (-2.0, 5.0), p)
        sol = sharding.solve_sharded(prob, ib.QuadGKB200(), local_solve=_oracle_solve, reltol=1e-10)
        par = np.asfortranarray(np.vstack([np.linspace(1, 3, 9), np.linspace(2, 1, 9), np.full(9, .3), np.full(9, .6)]))
        prob2 = ib.IntegralProblem(ib.GenzGaussian(), (np.zeros(2), np.ones(2)), par)
        sol2 = sharding.solve_sharded(prob2, ib.HCubatureB200(), local_solve=_oracle_solve, reltol=1e-7)
        sol3 = sharding.solve_sharded(prob, ib.QuadGKB200(), local_solve=_oracle_solve, interleave=True, reltol=1e-10)
        assert np.array_equal(sol3.u, sol.u) and np.array_equal(sol3.stats["nevals"], sol.stats["nevals"])     # same answers, other partition
        # vector-valued batch (VectorOf: p is nparam x ncomp x nprob; the problem axis is the last one)
        pv = np.asfortranarray(np.vstack([np.linspace(1, 3, 3 * 7), np.linspace(2, 1, 3 * 7), np.full(21, .3), np.full(21, .6)]).reshape(4, 3, 7, order="F"))
        probv = ib.IntegralProblem(ib.VectorOf(ib.GenzGaussian()), (np.zeros(2), np.ones(2)), pv)
        solv = sharding.solve_sharded(probv, ib.HCubatureB200(), local_solve=_oracle_solve, reltol=1e-7)
        import oracle as O
        Iv, Ev, nev, stv = O.hcubature_vec_batch("gauss", np.zeros(2), np.ones(2), pv, reltol=1e-7)
        assert solv.u.shape == (3, 7) and np.array_equal(solv.u, Iv) and np.array_equal(solv.resid, Ev) and np.array_equal(solv.stats["nevals"], nev)
        q.put((rank, sol.u, sol.resid, sol.stats["nevals"], sol2.u, sol2.stats["status"]))
    finally:
        dist.destroy_process_group()


def test_solve_sharded_world2_gloo():
    import torch.multiprocessing as mp
    import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in procs]
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    p = (0.1 + 20.0 * np.arange(37) / 36).reshape(1, -1)
    I, E, ne, st = O.quadgk_batch("sinxp", -2.0, 5.0, p, reltol=1e-10)
    par = np.asfortranarray(np.vstack([np.linspace(1, 3, 9), np.linspace(2, 1, 9), np.full(9, .3), np.full(9, .6)]))
    I2, *_ = O.hcubature_batch("gauss", np.zeros(2), np.ones(2), par, reltol=1e-7)
    for rank, u, resid, nev, u2, st2 in res:            # every rank holds the full, ordered result
        assert np.array_equal(u, I) and np.array_equal(resid, E) and np.array_equal(nev, ne)
        assert np.array_equal(u2, I2) and np.all(st2 == 0)


def _oracle_sampled_solve(prob, alg):
    import oracle as O
    from integrals_b200 import interface
    x = prob.x
    kw = dict(h=x.step) if isinstance(x, ib.Range) else dict(x=np.asarray(x))
    u = O.sampled_evalrule(alg.rule, np.asfortranarray(prob.y), dim=prob.dim, **kw)
    return interface.IntegralSolution(u, None, prob, alg, interface.ReturnCode.Success)


def _sampled_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        out = []
        for shape, dim in (((7, 101), 1), ((101, 7), 0), ((5, 64, 3), 1), ((4, 6, 33), 2)):
            y = rng.standard_normal(shape) + 1.5
            for alg, x in ((ib.TrapezoidalB200(), ib.Range(0.0, 1.0, shape[dim])), (ib.SimpsonsB200(), np.sort(rng.random(shape[dim])))):
                sol = sharding.solve_sampled_sharded(ib.SampledIntegralProblem(y, x, dim=dim), alg, local_solve=_oracle_sampled_solve)
                out.append(np.asarray(sol.u))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_solve_sampled_sharded_world2_gloo():
    """BASELINE config 2 across ranks: the rows of y are independent -> each rank integrates its slab, no data-path
    collective; every rank ends with the full result, equal to the unsharded solve bit for bit"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sampled_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in procs]
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    rng = np.random.default_rng(3)
    ref = []
    for shape, dim in (((7, 101), 1), ((101, 7), 0), ((5, 64, 3), 1), ((4, 6, 33), 2)):
        y = rng.standard_normal(shape) + 1.5
        for alg, x in ((ib.TrapezoidalB200(), ib.Range(0.0, 1.0, shape[dim])), (ib.SimpsonsB200(), np.sort(rng.random(shape[dim])))):
            ref.append(np.asarray(_oracle_sampled_solve(ib.SampledIntegralProblem(y, x, dim=dim), alg).u))
    for rank, out in res:
        assert len(out) == len(ref)
        for a, b in zip(out, ref):
            assert a.shape == b.shape and np.array_equal(a, b)
    assert sharding.sampled_shard_axis((4096, 1 << 20), 1) == 0 and sharding.sampled_shard_axis((5, 64, 3), 1) == 0
    assert sharding.sampled_shard_axis((4, 6, 33), 0) == 2


def test_round_oracle_presplit_variants_agree():
    """orc_hcubature_rounds with a pre-bisected domain (m0 > 0) reaches the same integral within the tolerance as the
    single-box start the driver uses on any number of ranks (replicated store: the refinement is rank-count independent)."""
    import oracle as O
    par = np.array([2.0, 3.0, 1.5, 0.3, 0.6, 0.5])
    base = O.hcubature_rounds("gauss", np.zeros(3), np.ones(3), par, reltol=1e-7, abstol=0.0, m0=0)
    for m0 in (4, 5, 6):
        I, E, s = O.hcubature_rounds("gauss", np.zeros(3), np.ones(3), par, reltol=1e-7, abstol=0.0, m0=m0)
        assert s["status"] == 0 and abs(I - base[0]) <= 1e-7 * abs(I) and E <= 1e-7 * abs(I)


def _oracle_slab(prob, alg, lo, hi):
    """this rank's slab of the integration axis with the weights of the WHOLE grid (CPU restatement of ib200_sampled_slab)"""
    import oracle as O
    x = prob.x
    n = np.asarray(prob.y).shape[prob.dim]
    w = O.sampled_weights(alg.rule, n, h=x.step) if isinstance(x, ib.Range) else O.sampled_weights(alg.rule, n, x=np.asarray(x))
    y = np.moveaxis(np.asarray(prob.y, dtype=np.float64), prob.dim, -1)
    return np.tensordot(y[..., lo:hi], w[lo:hi], axes=([-1], [0]))


def _split_axis_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(4)
        out = []
        for shape, dim in (((1001,), 0), ((7, 101), 1), ((5, 64, 3), 1), ((3, 2), 1)):
            y = rng.standard_normal(shape) + 1.5
            for alg, x in ((ib.TrapezoidalB200(), ib.Range(0.0, 1.0, shape[dim])), (ib.SimpsonsB200(), ib.Range(0.0, 2.0, shape[dim])),
                           (ib.TrapezoidalB200(), np.sort(rng.random(shape[dim])))):
                sol = sharding.solve_sampled_split_axis(ib.SampledIntegralProblem(y, x, dim=dim), alg, local_slab=_oracle_slab)
                out.append(np.asarray(sol.u))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_solve_sampled_split_axis_world2_gloo():
    """the integration axis cut into one slab per rank + ONE all-reduce of the partial results (preferred variant):
    a single long vector shards too, end corrections only at the global ends, every rank gets the whole answer"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_split_axis_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in procs]
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    rng = np.random.default_rng(4)
    ref = []
    for shape, dim in (((1001,), 0), ((7, 101), 1), ((5, 64, 3), 1), ((3, 2), 1)):
        y = rng.standard_normal(shape) + 1.5
        for alg, x in ((ib.TrapezoidalB200(), ib.Range(0.0, 1.0, shape[dim])), (ib.SimpsonsB200(), ib.Range(0.0, 2.0, shape[dim])),
                       (ib.TrapezoidalB200(), np.sort(rng.random(shape[dim])))):
            ref.append(np.asarray(_oracle_sampled_solve(ib.SampledIntegralProblem(y, x, dim=dim), alg).u))
    assert np.array_equal(res[0][1], res[1][1]) if False else all(np.array_equal(a, b) for a, b in zip(res[0][1], res[1][1]))   # the same on both ranks
    for a, b in zip(res[0][1], ref):
        assert a.shape == b.shape and np.allclose(a, b, rtol=1e-13, atol=1e-13)
