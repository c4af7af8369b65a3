"""One large domain on N GPUs (run under torchrun, one rank per GPU): ib200_hcubature_single with the library's
NCCL communicator and peer-mapped replicated stores, checked against the round-synchronous oracle.
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/single_domain_mgpu.py
Prints one JSON line on rank 0; exit code != 0 on a parity failure."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                     # noqa: E402
import torch.distributed as dist  # noqa: E402

import integrals_b200 as ib      # noqa: E402
from integrals_b200.workloads import genz_params  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lrank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lrank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    eng = ib.Engine(lrank)
    if world > 1:
        eng.comm_init_from_torch()
    F = ib.FAMILY_IDS
    out = dict(world=world, cases=[])
    ok = True
    for family, orc, dim, reltol, lbv, ubv in [("genz_gaussian", "gauss", 4, 1e-8, None, None), ("genz_oscillatory", "osc", 3, 1e-9, None, None),
                                                ("genz_gaussian", "gauss", 6, 1e-6, [-np.inf, -np.inf, 0, 0, 0, 0], [np.inf, np.inf, 1, 1, 1, 1])]:
        par = genz_params(family, dim, 1, seed=77, scale=0.5)[:, 0]
        lb = np.zeros(dim) if lbv is None else np.array(lbv, dtype=np.float64)
        ub = np.ones(dim) if ubv is None else np.array(ubv, dtype=np.float64)
        print(f"[rank {rank}] solve {family} {dim}-D", file=sys.stderr, flush=True)
        I, E, s = eng.hcubature_single(F[family], dim, lb, ub, par, reltol=reltol, abstol=0.0, max_boxes=1 << 24)
        print(f"[rank {rank}] done {family} {dim}-D rounds {s['rounds']}", file=sys.stderr, flush=True)
        rec = dict(family=family, dim=dim, I=I, E=E, **s, ms=eng.last_kernel_ms)
        if rank == 0:
            import oracle as O
            # replicated store: the refinement does not depend on the number of ranks (only the summation order of a round does)
            Ir, Er, sr = O.hcubature_rounds(orc, lb, ub, par, reltol=reltol, abstol=0.0, max_boxes=1 << 24)
            same = s["rule_applications"] == sr["rule_applications"] and s["rounds"] == sr["rounds"]
            good = s["status"] == 0 and abs(I - Ir) <= (1e-12 if same else reltol) * abs(Ir) and E <= reltol * abs(I)
            rec.update(oracle_I=Ir, oracle_applied=sr["rule_applications"], oracle_rounds=sr["rounds"], identical_refinement=same, ok=bool(good))
            ok = ok and good
        out["cases"].append(rec)
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.broadcast(flag, src=0)
        ok = bool(flag.item())
        eng.comm_destroy()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
