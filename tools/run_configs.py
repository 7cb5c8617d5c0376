"""Full-size timing of BASELINE.json configs C3, C4, C5 on the B200 (the bench line is C2).
Under torchrun the films (C3/C4) and the evaluation points (C5) are sharded over the ranks.

python tools/run_configs.py [c3] [c4] [c5]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import configs, parallel

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
comm = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{os.environ['LOCAL_RANK']}"))
    comm = parallel.DistComm()
which = [a for a in sys.argv[1:] if a in ("c3", "c4", "c5")] or ["c3", "c4", "c5"]
out = {"n_gpus": world}

def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

def timed(fn):
    sync(); t0 = time.perf_counter(); r = fn(); sync(); return r, time.perf_counter() - t0

def median_timed(fn, reps=7):
    ts = []
    for _ in range(reps):
        r, t = timed(fn); ts.append(t)
    return r, float(np.median(ts[1:]))

if "c3" in which:
    device, polys = configs.c3_susceptometer(4000)
    n = {k: len(m.sites) for k, m in device.meshes.items()}
    model, t_fact = median_timed(lambda: sc.factorize_model(device=device, current_units="uA", circulating_currents={"fc_center": "1 mA"}, comm=comm))
    sols, t_solve = median_timed(lambda: sc.solve(model=model, iterations=5))
    fl = sols[-1].hole_fluxoid("pl_center", points=polys["pl_center"], with_units=False)
    out["c3"] = {"vertices": n, "factorize_s": t_fact, "solve_iter5_s": t_solve, "M_Phi0_per_A": sum(fl) / 1e-3}
if "c4" in which:
    device, polys = configs.c4_ring_array(8, 5000)
    M, t_M = median_timed(lambda: np.array(device.mutual_inductance_matrix(polys, units="pH", iterations=5, comm=comm)), reps=5)
    out["c4"] = {"vertices_per_ring": len(device.meshes["ring0"].sites), "mutual_inductance_matrix_iter5_s": t_M,
                 "M00_pH": float(M[0, 0]), "M01_pH": float(M[0, 1]), "asym": float(np.abs(M - M.T).max() / abs(M[0, 1]))}
if "c5" in which:
    t0 = time.perf_counter(); device, fields = configs.c5_large(60000); t_mesh_host = time.perf_counter() - t0
    n = len(device.meshes["film"].sites)
    model, t_fact = timed(lambda: sc.factorize_model(device=device, current_units="uA"))
    n_int = len(model.film_systems["film"].indices)
    batch, t_batch_first = timed(lambda: sc.solve_batch(model=model, applied_fields=[sc.ConstantField(float(f)) for f in fields]))
    batch, t_batch = timed(lambda: sc.solve_batch(model=model, applied_fields=[sc.ConstantField(float(f)) for f in fields]))
    one, t_one = timed(lambda: sc.solve(model=model, applied_field=sc.ConstantField(1.0)))
    grid = configs.evaluation_grid(1000)
    for rep in range(2):
        Bz, t_field = timed(lambda: parallel.field_at_position_sharded(batch[9][0], grid, comm=comm, units="mT"))
    lin = float(np.linalg.norm(batch[63][0].film_solutions["film"].stream - fields[63] * one[0].film_solutions["film"].stream)
                / np.linalg.norm(fields[63] * one[0].film_solutions["film"].stream))
    # Lambda sweep: 4 refactorizations x 16 fields each; under torchrun one Lambda per rank (round-robin)
    mine = [lam for k, lam in enumerate(configs.C5_LAMBDA_SWEEP) if k % world == rank]
    def sweep():
        res = {}
        for lam in mine:
            m = sc.factorize_model(device=configs.with_lambda(device, lam), current_units="uA")
            sols = sc.solve_batch(model=m, applied_fields=[sc.ConstantField(float(f)) for f in fields[:16]])
            res[lam] = float(np.abs(sols[15][0].film_solutions["film"].stream).max())
        return res
    sweep_res, t_sweep = timed(sweep)
    out["c5"] = {"lambda_sweep_4x16_s": t_sweep, "lambda_sweep_max_stream_uA": sweep_res, "vertices": n, "n_interior": n_int, "host_mesh_s": t_mesh_host, "factorize_s": t_fact,
                 "lu_tflops_incl_assembly": (2 / 3) * n_int**3 / t_fact * 1e-12, "solve_64rhs_s": t_batch, "solve_64rhs_first_call_s": t_batch_first, "solve_1rhs_s": t_one,
                 "field_at_position_1M_s": t_field, "gpairs_per_s": 1e6 * n / t_field * 1e-9, "linearity_rel": lin,
                 "Bz_center_mT": float(Bz[len(Bz) // 2 + 500])}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
