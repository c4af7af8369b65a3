"""A/B of round-based kernel variants (IB200_LIB selects the library): per C4 family, kernel ms of the second of two launches
and a hash of (I, E, nevals, status) -- variants that only move data differently must reproduce the baseline bit for bit.
usage: probe_variants.py <nprob> [family ...]"""
import hashlib, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import integrals_b200 as ib
import integrals_b200.workloads as WL
eng = ib.Engine(0); F = ib.FAMILY_IDS
nprob = int(sys.argv[1]); d = 4
fams = sys.argv[2:] or ["genz_product_peak", "genz_oscillatory", "genz_gaussian", "genz_c0"]
tag = os.path.basename(os.environ.get("IB200_LIB", "baseline"))
if os.environ.get("PV_COMPACT"): eng.set_option("hcub_compact", int(os.environ["PV_COMPACT"])); tag += " compact=" + os.environ["PV_COMPACT"]
if os.environ.get("PV_LEVEL"): eng.set_option("hcub_compact_level", int(os.environ["PV_LEVEL"])); tag += " level=" + os.environ["PV_LEVEL"]
for fam in fams:
    par = torch.from_numpy(np.ascontiguousarray(WL.genz_params(fam, d, nprob).T)).cuda()
    oI = torch.empty(nprob, dtype=torch.float64, device="cuda"); oE = torch.empty_like(oI)
    oN = torch.empty(nprob, dtype=torch.int64, device="cuda"); oS = torch.empty(nprob, dtype=torch.int32, device="cuda")
    ms = []
    for rep in range(3):
        eng.hcubature_batch(F[fam], d, np.zeros(d), np.ones(d), par, reltol=1e-6, abstol=1e-8, maxevals=4000000, out_I=oI, mode=1,
                            out_E=oE, out_nevals=oN, out_status=oS)
        eng.synchronize(); ms.append(eng.last_kernel_ms)
    h = hashlib.sha1(oI.cpu().numpy().tobytes() + oE.cpu().numpy().tobytes() + oN.cpu().numpy().tobytes() + oS.cpu().numpy().tobytes()).hexdigest()[:12]
    print(f"{tag:26s} {fam:20s} ms {ms[1]:9.3f} {ms[2]:9.3f}  evals {int(oN.sum().item()):15d}  hash {h}", flush=True)
