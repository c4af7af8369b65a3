"""BASELINE config 5 (one 8-D Genz Gaussian, two infinite axes, abstol 0) to a requested tolerance through the two-level solve.
usage: probe_c5_two_level.py <reltol> [switch_boxes] [arena MiB]      (under torchrun: all ranks take part)"""
import json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import integrals_b200 as ib
import integrals_b200.workloads as WL
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
eng = ib.Engine(local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng.comm_init_from_torch()
reltol = float(sys.argv[1])
if len(sys.argv) > 2: eng.set_option("single_switch_boxes", int(sys.argv[2]))
if len(sys.argv) > 3: eng.set_option("single_arena_mib", int(sys.argv[3]))
if os.environ.get("C5_SWITCH_FACTOR"): eng.set_option("single_switch_factor", int(os.environ["C5_SWITCH_FACTOR"]))
if os.environ.get("C5_SUB_BOXES"): eng.set_option("single_sub_boxes", int(os.environ["C5_SUB_BOXES"]))
lb, ub, par, exact = WL.c5_problem()
F = ib.FAMILY_IDS
eng.hcubature_single(F["genz_gaussian"], 8, lb, ub, par, reltol=1e-4, abstol=0.0, max_boxes=1 << 40)     # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
I, E, s = eng.hcubature_single(F["genz_gaussian"], 8, lb, ub, par, reltol=reltol, abstol=0.0, max_boxes=1 << 40)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
res = dict(reltol=reltol, rank=rank, seconds=dt, kernel_ms=eng.last_kernel_ms, I=I, E=E, rel_E=E / abs(I), true_rel_err=abs(I - exact) / abs(exact),
           boxes_per_s=s["rule_applications"] / dt, **s)
print(json.dumps(res), flush=True)
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/c5_two_level_{reltol:g}_n{world}.json", "w") as f:
        json.dump(res, f, indent=1)
if world > 1:
    eng.comm_destroy()
    dist.destroy_process_group()
