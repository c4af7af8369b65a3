"""GPU parity of the two refinement drivers of ib200_hcubature_batch (through the C ABI):

  mode 0  reference order (one box per step, largest error first)     vs  orc_hcubature           (HCubature.jl restated)
  mode 1  rounds (every box above a per-problem threshold per round)  vs  orc_hcubature_rounds_*  (the same scheme restated)
          and vs orc_hcubature: (I, E) within the requested tolerance -- the bar of SURVEY 7.3.2

including BASELINE config 4 at its stated difficulty and cap (reltol 1e-6, abstol 1e-8, 4e6 evaluations), capped problems
included.  Replaces HCubature.hcubature as called at /root/reference/src/Integrals.jl:347-397.
"""
import numpy as np
import pytest

import integrals_b200 as ib
import integrals_b200.workloads as WL
from util_gpu import F, ORC_FAM, GENZ, genz_params, genz_exact

pytestmark = pytest.mark.gpu
# families for which Genz-Malik's error estimate is dependable at these difficulties.  Not the C0 / discontinuous families (non-smooth) and
# not the corner peak at the default abstol: its integrals are ~1e-5, abstol = 1e-8 ends the solve at a relative accuracy of ~1e-3, and two
# refinement orders that both stop at E <= abstol can sit further apart than abstol (the CPU restatements of the two orders do, too)
SMOOTH = ("genz_oscillatory", "genz_product_peak", "genz_gaussian")


def _mean(x):
    return float(np.mean(x)) if np.size(x) else 1.0


@pytest.fixture(scope="module")
def eng():
    return ib.Engine(0)


@pytest.fixture(scope="module")
def O():
    import oracle
    return oracle


def _agree(I, E, st, Ir, Er, str_, reltol, abstol):
    """the parity bar on (I, E): converged problems within the requested tolerance of each other, problems that ran out of
    budget within each other's error estimate; error estimates within a factor 4"""
    tol = np.maximum(abstol, reltol * np.abs(Ir))
    both = (st == 0) & (str_ == 0)
    ok_conv = np.abs(I - Ir)[both] <= tol[both]
    capped = (st == 1) | (str_ == 1)
    ok_cap = np.abs(I - Ir)[capped] <= np.maximum(tol, np.minimum(E, Er))[capped]
    ok_E = (E <= 4 * Er + 1e-300) & (Er <= 4 * E + 1e-300)
    return ok_conv, ok_cap, ok_E


@pytest.mark.parametrize("family", GENZ)
@pytest.mark.parametrize("dim", [2, 3, 4, 5, 6, 7, 8])
def test_rounds_match_their_oracle(eng, O, family, dim):
    """mode 1 against the CPU restatement of the same round scheme: the same refinement (identical evaluation counts) except where
    a decision sits on a rounding-level tie (the kernel forms a point's value from per-axis table entries, with FMA contraction)"""
    nprob = 24 if dim <= 5 else 12          # dim <= 5: one thread per box; dim >= 6: thread per box (product families) / warp per box
    par = genz_params(family, dim, nprob, scale=0.25)
    lb, ub = np.zeros(dim), np.ones(dim)
    reltol, abstol, cap = 1e-6, 1e-8, 400000
    I, E, ne, st = eng.hcubature_batch(F[family], dim, lb, ub, par, reltol=reltol, abstol=abstol, maxevals=cap, mode="rounds")
    Ir, Er, ner, str_, _ = O.hcubature_rounds_batch(ORC_FAM[family], lb, ub, par, reltol=reltol, abstol=abstol, maxiters=cap, nthreads=O.max_threads())
    same = ne == ner
    assert np.mean(same) >= 0.7
    assert np.all(st[same] == str_[same])
    scale = np.maximum(np.abs(Ir), abstol)
    assert np.all(np.abs(I - Ir)[same] <= 1e-10 * scale[same]) and np.all(np.abs(E - Er)[same] <= 1e-6 * np.maximum(Er[same], 1e-300) + 1e-10 * scale[same])
    ok_conv, ok_cap, ok_E = _agree(I, E, st, Ir, Er, str_, reltol, abstol)
    assert np.all(ok_conv) and np.all(ok_cap) and np.all(ok_E)


@pytest.mark.parametrize("family", GENZ)
def test_rounds_agree_with_reference_order(eng, O, family):
    """mode 1 against the reference-order restatement (orc_hcubature): the north-star bar for adaptive solves"""
    dim, nprob = 4, 32
    par = genz_params(family, dim, nprob, scale=0.5)
    lb, ub = np.zeros(dim), np.ones(dim)
    reltol, abstol = 1e-6, 1e-8
    I, E, ne, st = eng.hcubature_batch(F[family], dim, lb, ub, par, reltol=reltol, abstol=abstol, maxevals=4_000_000, mode="rounds")
    Ir, Er, ner, str_ = O.hcubature_batch(ORC_FAM[family], lb, ub, par, reltol=reltol, abstol=abstol, maxiters=4_000_000, nthreads=O.max_threads())
    tol = np.maximum(abstol, reltol * np.abs(Ir))
    conv = (st == 0) & (str_ == 0)
    assert np.mean(st == str_) >= 0.9
    assert np.all(E[st == 0] <= tol[st == 0] * (1 + 1e-9))
    if family in SMOOTH:
        assert np.all(np.abs(I - Ir)[conv] <= tol[conv])
        exact = np.array([genz_exact(family, par[:, k], dim) for k in range(nprob)])
        assert np.all(np.abs(I - exact)[conv] <= 10 * tol[conv])
    else:
        # C0 / discontinuous / corner peak at abstol: Genz-Malik's estimate is not a bound here, so two refinement orders that both
        # stop at E <= tol can differ by more than tol (the CPU restatements of the two orders do too); the round order must be no
        # further from the closed form than the reference order is, up to a factor
        exact = np.array([genz_exact(family, par[:, k], dim) for k in range(nprob)])
        e1, e0 = np.abs(I - exact), np.abs(Ir - exact)
        assert np.all(e1[conv] <= np.maximum(10 * tol[conv], 4 * e0[conv]))
    assert np.all(ne[conv] <= 1.6 * ner[conv] + 5700)      # rounds cost 5-20 % more evaluations on these families, never a multiple


def test_rounds_initdiv_infinite_and_degenerate(eng, O):
    lb, ub = np.zeros(3), np.ones(3)
    par = genz_params("genz_gaussian", 3, 6, scale=0.5)
    I, E, ne, st = eng.hcubature_batch(F["genz_gaussian"], 3, lb, ub, par, reltol=1e-7, abstol=0.0, initdiv=3, mode="rounds")
    Ir, Er, ner, str_, _ = O.hcubature_rounds_batch("gauss", lb, ub, par, reltol=1e-7, abstol=0.0, initdiv=3)
    assert np.array_equal(st, str_) and np.all(np.abs(I - Ir) <= 1e-7 * np.abs(Ir)) and np.mean(ne == ner) >= 0.5
    # infinite and semi-infinite axes (transformation_if_inf and the tan / cot maps), per-problem bounds
    par = np.asfortranarray(np.array([[1.0, 2.0, 0.2, -0.4], [1.5, 0.7, 0.0, 0.3], [0.8, 1.1, 0.5, 0.1]]).T)
    lbp = np.asfortranarray(np.array([[-np.inf, -np.inf], [0.0, -np.inf], [-np.inf, -1.0]]).T)
    ubp = np.asfortranarray(np.array([[np.inf, np.inf], [np.inf, 2.0], [0.5, np.inf]]).T)
    for trn, tr in (("if_inf", 0), ("tan", 1), ("cot", 2)):
        I, E, ne, st = eng.hcubature_batch(F["genz_gaussian"], 2, lbp, ubp, par, reltol=1e-7, abstol=1e-10, transform=tr, mode="rounds")
        Ir, Er, ner, str_, _ = O.hcubature_rounds_batch("gauss", lbp, ubp, par, tr=trn, reltol=1e-7, abstol=1e-10)
        Ih, *_ = O.hcubature_batch("gauss", lbp, ubp, par, tr=trn, reltol=1e-7, abstol=1e-10)
        assert np.array_equal(st, str_) and np.all(st == 0)
        assert np.all(np.abs(I - Ir) <= 1e-7 * np.abs(Ir) + 1e-10) and np.all(np.abs(I - Ih) <= 1e-7 * np.abs(Ih) + 1e-10)
    # degenerate domain (lb == ub on one axis): HCubature returns after the first box
    I, E, ne, st = eng.hcubature_batch(F["genz_gaussian"], 2, np.array([0.0, 0.5]), np.array([1.0, 0.5]), par[:, :1], mode="rounds")
    assert I[0] == 0.0 and ne[0] == 17 and st[0] == 0
    # budget: a solve that cannot converge stops at the cap with status MAXEVALS, like the oracle
    par = genz_params("genz_product_peak", 4, 4)
    I, E, ne, st = eng.hcubature_batch(F["genz_product_peak"], 4, np.zeros(4), np.ones(4), par, reltol=1e-12, abstol=0.0, maxevals=20000, mode="rounds")
    Ir, Er, ner, str_, _ = O.hcubature_rounds_batch("prodpeak", np.zeros(4), np.ones(4), par, reltol=1e-12, abstol=0.0, maxiters=20000)
    assert np.all(st == 1) and np.array_equal(str_, st) and np.all(ne >= 20000) and np.all(np.abs(I - Ir) <= np.minimum(E, Er))
    with pytest.raises(ib.IB200Error):          # 1-D (the Gauss-Kronrod rule) runs in reference order only
        eng.hcubature_batch(F["genz_gaussian"], 1, np.zeros(1), np.ones(1), genz_params("genz_gaussian", 1, 2), mode="rounds")


@pytest.mark.parametrize("family", ["genz_gaussian", "genz_product_peak", "genz_c0"])
@pytest.mark.parametrize("dim", [6, 7, 8])
def test_rounds_high_dim_infinite_axes(eng, O, family, dim):
    """6 .. 8 dimensions with infinite and semi-infinite axes (the thread-per-box rule of gm_rule_thread3.cuh with the Jacobian folded
    into the per-axis terms) against the CPU restatement of the same rounds and against the reference order"""
    nprob = 6
    par = genz_params(family, dim, nprob, scale=0.25)
    par[:dim] = np.maximum(par[:dim], 0.6)                     # decay fast enough on the infinite axes
    lb, ub = np.zeros(dim), np.ones(dim)
    lb[0], ub[0] = -np.inf, np.inf
    ub[1] = np.inf
    lb[dim - 1] = -np.inf
    reltol, abstol, cap = 1e-4, 0.0, 600000
    for trn, tr in (("if_inf", 0), ("tan", 1)):
        I, E, ne, st = eng.hcubature_batch(F[family], dim, lb, ub, par, reltol=reltol, abstol=abstol, maxevals=cap, transform=tr, mode="rounds")
        Ir, Er, ner, str_, _ = O.hcubature_rounds_batch(ORC_FAM[family], lb, ub, par, tr=trn, reltol=reltol, abstol=abstol, maxiters=cap,
                                                       nthreads=O.max_threads())
        assert np.mean(ne == ner) >= 0.5 and np.mean(st == str_) >= 0.8
        ok_conv, ok_cap, ok_E = _agree(I, E, st, Ir, Er, str_, reltol, abstol)
        assert np.all(ok_conv) and np.all(ok_cap) and np.all(ok_E)
        same = ne == ner
        assert np.all(np.abs(I - Ir)[same] <= 1e-10 * np.abs(Ir[same]))


def test_rounds_result_independent_of_batch(eng):
    """a problem's result is bitwise the same whatever batch (and warp) it is solved in"""
    d = 4
    par = genz_params("genz_gaussian", d, 4096, scale=0.5)
    lb, ub = np.zeros(d), np.ones(d)
    I, E, ne, st = eng.hcubature_batch(F["genz_gaussian"], d, lb, ub, par, reltol=1e-6, abstol=1e-8, maxevals=4_000_000, mode="rounds")
    pick = np.random.default_rng(5).choice(4096, size=300, replace=False)
    I2, E2, ne2, st2 = eng.hcubature_batch(F["genz_gaussian"], d, lb, ub, np.asfortranarray(par[:, pick]), reltol=1e-6, abstol=1e-8,
                                          maxevals=4_000_000, mode="rounds")
    assert np.array_equal(I2, I[pick]) and np.array_equal(E2, E[pick]) and np.array_equal(ne2, ne[pick]) and np.array_equal(st2, st[pick])


@pytest.mark.parametrize("family", GENZ)
def test_c4_stated_config_both_modes(eng, O, family, capsys):
    """BASELINE config 4 as stated (standard Genz difficulty, reltol 1e-6, abstol 1e-8, cap 4e6 evaluations), 256 problems per family,
    BOTH drivers against their CPU restatements, capped problems included.  Prints the share of problems refined identically."""
    c = WL.C4; d = c["dim"]; nprob = 256
    par = WL.genz_params(family, d, nprob)
    lb, ub = np.zeros(d), np.ones(d)
    kw = dict(reltol=c["reltol"], abstol=c["abstol"])
    nt = O.max_threads()
    eng.set_option("hcub_mode", 2)           # the pair-mode kernel (what a full-size batch runs) even for this small batch
    try:
        I0, E0, ne0, st0 = eng.hcubature_batch(F[family], d, lb, ub, par, maxevals=c["maxiters"], mode="reference_order", **kw)
    finally:
        eng.set_option("hcub_mode", 0)
    Ir0, Er0, ner0, str0 = O.hcubature_batch(ORC_FAM[family], lb, ub, par, maxiters=c["maxiters"], nthreads=nt, **kw)
    I1, E1, ne1, st1 = eng.hcubature_batch(F[family], d, lb, ub, par, maxevals=c["maxiters"], mode="rounds", **kw)
    Ir1, Er1, ner1, str1, _ = O.hcubature_rounds_batch(ORC_FAM[family], lb, ub, par, maxiters=c["maxiters"], nthreads=nt, **kw)
    rep = {}
    for name, (I, E, ne, st, Ir, Er, ner, str_) in {"reference_order": (I0, E0, ne0, st0, Ir0, Er0, ner0, str0),
                                                     "rounds": (I1, E1, ne1, st1, Ir1, Er1, ner1, str1)}.items():
        ok_conv, ok_cap, ok_E = _agree(I, E, st, Ir, Er, str_, c["reltol"], c["abstol"])
        tol = np.maximum(c["abstol"], c["reltol"] * np.abs(Ir))
        flip = st != str_
        rep[name] = dict(identical_nevals=_mean(ne == ner), status_equal=_mean(st == str_), capped=_mean(st == 1),
                         conv_within_tol=_mean(ok_conv), capped_within_E=_mean(ok_cap), E_within_x4=_mean(ok_E),
                         worst_E_over_tol_where_status_flips=float(np.max(np.maximum(E, Er)[flip] / tol[flip])) if flip.any() else None)
        # the reference-order driver reproduces its restatement's sequence of boxes; the round driver decides on thresholds over many
        # boxes at once, so a rounding-level difference in one error estimate (table-form rule, FMA contraction) can move a box to the
        # next round -- evaluation counts then differ, and a status can flip for a problem that converges within a round of the cap
        assert _mean(st == str_) >= (0.98 if name == "reference_order" else 0.9), (name, rep[name])
        assert np.all(np.maximum(E, Er)[flip] <= 8 * tol[flip]), (name, rep[name])      # ... and only there
        assert np.all(ok_E[~flip]), (name, rep[name])
        if family in SMOOTH:
            assert np.all(ok_conv) and np.all(ok_cap), (name, rep[name])
        else:
            assert _mean(ok_conv) >= 0.9 and _mean(ok_cap) >= 0.9, (name, rep[name])
        # where the refinement was the same the results agree to rounding
        same = (ne == ner) & (st == str_)
        assert np.all(np.abs(I - Ir)[same] <= 1e-9 * np.maximum(np.abs(Ir[same]), c["abstol"]))
    with capsys.disabled():
        print(f"\n[c4 parity] {family}: {rep}")


@pytest.mark.parametrize("family", list(GENZ) + ["cosprod", "explin"])
@pytest.mark.parametrize("dim", [2, 3, 4, 5])
def test_compact_records_reproduce_full_records(eng, family, dim):
    """the 32-byte dyadic box records of the round-based kernel (hcubature_rb.cuh kCptBytes: batch mode, finite domains, initdiv = 1) decode
    to exactly the corners the full records hold, so (I, E, nevals, status) are the same doubles with them, without them, and when the
    hand-back to the full-record kernel is forced early (a box halved more than 3 times along one axis) -- per-problem bounds included"""
    nprob = 96
    if family in GENZ:
        par = genz_params(family, dim, nprob, scale=0.5)
    else:
        rng = np.random.default_rng(31 + dim)
        par = np.asfortranarray(0.5 + 4.0 * rng.random((dim, nprob)))
    rng = np.random.default_rng(7)
    # lb >= 0 keeps the corner peak's denominator 1 + a.x positive: with the singular line inside the domain the integral does not exist
    # (the reference-order driver ends those problems with status 2; see DESIGN.md "Known issue" for the round-based one)
    lb = np.asfortranarray(0.05 * rng.random((dim, nprob))); ub = np.asfortranarray(1.0 + 0.25 * rng.random((dim, nprob)))
    res = {}
    try:
        for tag, compact, level in (("compact", 1, 31), ("full", 0, 31), ("handed back", 1, 3)):
            eng.set_option("hcub_compact", compact); eng.set_option("hcub_compact_level", level)
            res[tag] = eng.hcubature_batch(F[family], dim, lb, ub, par, reltol=1e-6, abstol=1e-9, maxevals=150000, mode="rounds")
    finally:
        eng.set_option("hcub_compact", 1); eng.set_option("hcub_compact_level", 31)
    assert res["full"][2].max() > 50 * (2 ** dim + 2 * dim * dim + 2 * dim + 1)        # the problems do refine
    for tag in ("compact", "handed back"):
        for x, y in zip(res[tag], res["full"]):
            assert np.array_equal(x, y), (tag, family, dim)


def test_compact_records_options_are_checked(eng):
    with pytest.raises(ib.IB200Error, match="hcub_compact_level"):
        eng.set_option("hcub_compact_level", 32)
    with pytest.raises(ib.IB200Error, match="hcub_compact_level"):
        eng.set_option("hcub_compact_level", 0)
