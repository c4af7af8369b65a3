"""Test infrastructure: a line-by-line numpy/Python restatement of the DEVICE side of the BatchIntegralFunction bridge --
bb::points_kernel / bb::combine_kernel (csrc/batch_bridge.cu) and the decide / count / scan / scatter bookkeeping of
csrc/hcub_rounds.cuh -- behind the same batch_open / next / points / values / result protocol as Engine.  It lets the CPU suite
check (i) that this round scheme, fed by a caller-evaluated integrand, refines exactly like the C oracle orc_hcubature_rounds,
and (ii) the host loop Engine.integrate_batch and the solve-interface routing, without a GPU.  It is not a product path and not
a fallback: nothing under integrals.jl_b200/ imports it."""
import math
import numpy as np

C_X = [-0.991455371120812639206854697526329, -0.949107912342758524526189684047851, -0.864864423359769072789712788640926,
       -0.741531185599394439863864773280788, -0.586087235467691130294144838258730, -0.405845151377397166906606412076961,
       -0.207784955007898467600689403773245, 0.0]
C_W = [0.022935322010529224963732008058970, 0.063092092629978553290700663189204, 0.104790010322250183839876322541518,
       0.140653259715525918745189590510238, 0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
       0.204432940075298892414161999234649, 0.209482141084727828012999174891714]
C_WG = [0.129484966168869693270611432679082, 0.279705391489276667901467771423780, 0.381830050505118944950369775488975,
        0.417959183673469387755102040816327]


def npoints(D):
    return 15 if D == 1 else 1 + 4 * D + 2 * D * (D - 1) + (1 << D)


def point_table(D):
    NP = npoints(D)
    pt = []
    for q in range(NP):
        e = 0; cls = 0
        if q == 0:
            cls = 0
        elif q < 1 + 2 * D:
            r = q - 1; cls = 1; e = (1 + (r & 1)) << (3 * (r >> 1))
        elif q < 1 + 4 * D:
            r = q - 1 - 2 * D; cls = 2; e = (3 + (r & 1)) << (3 * (r >> 1))
        elif q < 1 + 4 * D + 2 * D * (D - 1):
            r = q - 1 - 4 * D; s = r & 3; r >>= 2
            i, j = 0, 1
            for _ in range(r):
                j += 1
                if j == D:
                    i += 1; j = i + 1
            cls = 3; e = ((3 + ((s >> 1) & 1)) << (3 * i)) | ((3 + (s & 1)) << (3 * j))
        else:
            r = q - (1 + 4 * D + 2 * D * (D - 1)); cls = 4
            for k in range(D):
                e |= (5 + ((r >> k) & 1)) << (3 * k)
        pt.append(e | (cls << 24))
    return pt


def t2ujac(tr, t, lb, ub):
    il, iu = math.isinf(lb), math.isinf(ub)
    if not il and not iu:
        den = (ub - lb) * 0.5
        return lb + (1.0 + t) * den, den
    if tr == 0:
        if il and iu:
            den = 1.0 / (1.0 - t * t)
            return t * den, (1.0 + t * t) * (den * den)
        sg = (-1.0 if math.copysign(1.0, lb if il else ub) < 0 else 1.0)
        den = 1.0 / (1.0 - t * sg)
        return (ub if il else lb) + t * den, den * den
    if tr == 2 and iu and not il:
        arg = (math.pi / 2) * (1.0 - t)
        c = 1.0 / math.tan(arg)
        return lb + c, (math.pi / 2) * (1.0 + c * c)
    tv = math.tan((math.pi / 2) * t)
    u = tv if (il and iu) else ((ub if il else lb) + tv)
    return u, (math.pi / 2) * (1.0 + tv * tv)


def u2t(lb, ub):
    il, iu = math.isinf(lb), math.isinf(ub)
    fs = lambda y: -1.0 if math.copysign(1.0, y) < 0 else 1.0
    if il and iu: return fs(lb), fs(ub)
    if il: return fs(lb), 0.0
    if iu: return 0.0, fs(ub)
    return -1.0, 1.0


class BridgeModel:
    def __init__(self, frac=0.5):
        self.frac = frac
        self.phase = 0

    def batch_open(self, dim, lb, ub, reltol=1e-8, abstol=1e-8, max_boxes=1 << 20, store_boxes=0, transform=0):
        self.D = dim; self.lb = list(np.atleast_1d(lb).astype(float)); self.ub = list(np.atleast_1d(ub).astype(float))
        self.rtol, self.atol, self.max_boxes, self.tr = reltol, abstol, max_boxes, transform
        self.cap = store_boxes if store_boxes > 0 else (max_boxes + 1024 if max_boxes < (1 << 22) else 1 << 22)
        D = dim
        a = []; b = []
        for k in range(D):
            tl, th = u2t(self.lb[k], self.ub[k]); a.append(tl); b.append(th)
        self.rec = [a + b + [0.0, 0.0]]
        self.err = [0.0]
        self.todo = [0]
        self.I = 0.0; self.E = 0.0; self.napplied = 0; self.bound = 0.0; self.theta = 0.0
        self.done = False; self.status = 0; self.full = False; self.round = 0
        self.NP = npoints(D); self.pt = point_table(D) if D >= 2 else None
        self.phase = 1

    def batch_next(self):
        assert self.phase in (1, 3)
        return 0 if self.phase == 3 else len(self.todo) * self.NP

    def batch_points(self, x, n):
        assert self.phase == 1 and n == len(self.todo) * self.NP
        D, NP = self.D, self.NP
        l2, l3, l5 = math.sqrt(9.0 / 70.0), math.sqrt(9.0 / 10.0), math.sqrt(9.0 / 19.0)
        self.jac = np.empty(n)
        xv = x.reshape(-1, order="F") if D > 1 else x
        for j in range(n):
            t = j // NP; q = j - t * NP
            r = self.rec[self.todo[t]]
            J = 1.0
            if D == 1:
                a, b = r[0], r[1]
                c = (a + b) * 0.5; Dl = (b - a) * 0.5
                dx = Dl * C_X[q if q < 7 else q - 7] if q < 14 else 0.0
                tt = c + dx if q < 7 else (c - dx if q < 14 else c)
                u, J = t2ujac(self.tr, tt, self.lb[0], self.ub[0])
                xv[j] = u
            else:
                e = self.pt[q]
                for k in range(D):
                    o = (e >> (3 * k)) & 7
                    a, b = r[k], r[D + k]
                    c = 0.5 * (a + b); h = 0.5 * abs(b - a)
                    lam = 0.0 if o == 0 else (l2 if o <= 2 else (l3 if o <= 4 else l5))
                    dl = h * lam
                    tt = c if o == 0 else (c + dl if (o & 1) else c - dl)
                    u, jk = t2ujac(self.tr, tt, self.lb[k], self.ub[k])
                    xv[j * D + k] = u
                    J = jk if k == 0 else J * jk
            self.jac[j] = J
        if D > 1:
            x[...] = xv.reshape(x.shape, order="F")
        self.phase = 2

    def batch_values(self, y, n):
        assert self.phase == 2 and n == len(self.todo) * self.NP
        D, NP = self.D, self.NP
        NAX = 1 + 4 * D
        if D >= 2:
            twon = float(1 << D)
            w = [twon * ((12824 - 9120 * D + 400 * D * D) / 19683.0), twon * (980 / 6561.0), twon * ((1820 - 400 * D) / 19683.0),
                 twon * (200 / 19683.0), 6859 / 19683.0, twon * ((729 - 950 * D + 50 * D * D) / 729.0), twon * (245 / 486.0),
                 twon * ((265 - 100 * D) / 1458.0), twon * (25 / 729.0), 0.0]
        dI = 0.0; dE = 0.0; mx = 0.0
        for t, slot in enumerate(self.todo):
            r = self.rec[slot]
            f = y[t * NP:(t + 1) * NP] * self.jac[t * NP:(t + 1) * NP]
            if D == 1:
                sI = sum(f[q] * (C_W[q if q < 7 else q - 7] if q < 14 else C_W[7]) for q in range(NP))
                def wg(q):
                    i7 = q if q < 7 else q - 7
                    return (C_WG[i7 >> 1] if (i7 & 1) else 0.0) if q < 14 else C_WG[3]
                sP = sum(f[q] * wg(q) for q in range(NP))
                Dl = (r[1] - r[0]) * 0.5
                Iv = sI * Dl; Er = abs(Iv - sP * Dl); kd = 0
            else:
                sI = sum(w[self.pt[q] >> 24] * f[q] for q in range(NP))
                sP = sum(w[5 + (self.pt[q] >> 24)] * f[q] for q in range(NP))
                hh = [0.5 * abs(r[D + k] - r[k]) for k in range(D)]
                V = hh[0]
                for k in range(1, D): V *= hh[k]
                Iv = V * sI; Er = abs(Iv - V * sP)
                deltaf = Er / (10.0 ** D * V); twelvef1 = 12 * f[0]
                kd = 0; maxdd = 0.0; hkd = hh[0]
                for k in range(D):
                    f2k = f[1 + 2 * k] + f[2 + 2 * k]
                    f3k = f[1 + 2 * D + 2 * k] + f[2 + 2 * D + 2 * k]
                    dd = abs(f3k + twelvef1 - 7 * f2k)
                    delta = dd - maxdd
                    if delta > deltaf: kd = k; maxdd = dd; hkd = hh[k]
                    elif abs(delta) <= deltaf and hh[k] > hkd: kd = k; hkd = hh[k]
            r[2 * D] = Iv; r[2 * D + 1] = float(kd); self.err[slot] = Er
            dI += Iv; dE += Er; mx = max(mx, Er)
        self._bookkeep(dI, dE, mx)
        while not self.done and not self.todo:
            self._bookkeep(0.0, 0.0, 0.0)

    def _bookkeep(self, dI, dE, mx):
        D = self.D
        # decide_kernel
        self.I += dI; self.E += dE
        self.napplied += len(self.todo)
        self.round += 1
        self.todo = []
        rtol = self.rtol
        if rtol == 0.0 and self.atol == 0.0: rtol = 1.4901161193847656e-08
        tol = max(rtol * abs(self.I), self.atol)
        if self.E <= tol: self.done = True; self.status = 0
        elif not math.isfinite(self.E): self.done = True; self.status = 2
        elif self.napplied >= self.max_boxes or self.full: self.done = True; self.status = 1
        if self.done:
            self.phase = 3; return
        self.bound = max(mx, self.theta)
        theta = 0.5 * self.bound
        cnt = [0.0] * 2048
        for e in self.err:
            cnt[(np.float64(e).view(np.uint64) >> np.uint64(52)) & np.uint64(0x7ff)] += 1.0
        occ = [b for b in range(2048) if cnt[b]]
        b0 = occ[0]; b1 = min(occ[-1], 2046)
        bh = b0 - 1; low = 0.0; broke = False
        for b in range(b0, b1 + 1):
            c = cnt[b]
            if c != 0.0:
                low += c * (0.0 if b == 0 else math.ldexp(1.0, b - 1023))
                if not (low <= tol): broke = True; break
            bh = b
        if not broke: bh = 2046
        remaining = float(self.max_boxes) - self.napplied
        bb = b1 + 1; above = 0.0; broke = False
        for b in range(b1, b0 - 1, -1):
            above += cnt[b]
            if 2.0 * above > remaining: broke = True; break
            bb = b
        if not broke: bb = 0
        bf = b1 + 1; top = 0.0; broke = False
        for b in range(b1, b0 - 1, -1):
            top += cnt[b]
            if top > self.frac * len(self.err): broke = True; break
            bf = b
        if not broke: bf = 0
        if bf > bb: bb = bf
        bs = bh + 1 if bh + 1 > bb else bb
        theta_h = 0.0 if bs <= 0 else (1.7976931348623157e308 if bs >= 2047 else math.ldexp(1.0, bs - 1023))
        self.theta = min(theta, theta_h)
        # count / scan / scatter
        D2 = 2 * D
        idx = [i for i, e in enumerate(self.err) if e >= self.theta]
        if len(self.err) + len(idx) > self.cap:
            self.full = True
            # next exchange ends the solve; the CUDA path has ntodo = 0 here
            self.phase = 1
            return
        n0 = len(self.err)
        for rank_in, i in enumerate(idx):
            r = self.rec[i]
            self.I -= r[D2]; self.E -= self.err[i]
        for rank_in, i in enumerate(idx):
            r = self.rec[i]
            kd = int(r[D2 + 1]); a = r[:D]; b = r[D:D2]
            wd = (b[kd] - a[kd]) * 0.5
            child = a + [(b[k] - wd) if k == kd else b[k] for k in range(D)] + [0.0, 0.0]
            r[kd] = a[kd] + wd
            self.rec.append(child); self.err.append(0.0)
            assert len(self.rec) - 1 == n0 + rank_in
            self.todo += [i, n0 + rank_in]
        self.phase = 1

    def batch_result(self):
        assert self.phase == 3
        D2 = 2 * self.D
        I = sum(r[D2] for r in self.rec); E = sum(self.err)
        return I, E, dict(nevals=self.napplied * self.NP, rule_applications=self.napplied, boxes=len(self.rec), rounds=self.round, status=self.status)
