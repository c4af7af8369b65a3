"""BASELINE config 5 probe: one 8-D Genz Gaussian, axes 1-2 on (-inf, inf), axes 3-8 on [0,1], reltol 1e-9, abstol 0,
at a sequence of box budgets (python scripts/probe_c5.py [log2 budgets...]); under torchrun it uses all ranks."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                      # noqa: E402
import torch.distributed as dist  # noqa: E402

import integrals_b200 as ib       # noqa: E402
from integrals_b200.workloads import c5_problem  # noqa: E402

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lrank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lrank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
eng = ib.Engine(lrank)
if world > 1:
    eng.comm_init_from_torch()
lb, ub, par, exact = c5_problem()
budgets = [int(a) for a in sys.argv[1:]] or [16, 20, 22]
res = []
for lg in budgets:
    for rep in range(2):
        I, E, s = eng.hcubature_single(ib.FAMILY_IDS["genz_gaussian"], 8, lb, ub, par, reltol=1e-9, abstol=0.0, max_boxes=1 << lg)
        ms = eng.last_kernel_ms
    rec = dict(log2_budget=lg, I=I, E=E, rel_E=E / abs(I), true_rel_err=abs(I - exact) / abs(exact), ms=ms,
               boxes_per_s=s["rule_applications"] / (ms * 1e-3), gevals_per_s=s["nevals"] / (ms * 1e-3) / 1e9, **s)
    res.append(rec)
    if rank == 0:
        print(json.dumps(rec), flush=True)
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/probe_c5_w{world}.json", "w"), indent=1)
if world > 1:
    eng.comm_destroy(); dist.destroy_process_group()
