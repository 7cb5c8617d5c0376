"""Where the wall time of C3 / C4 goes (host sections timed with a device sync on both sides).
python tools/prof_configs.py  [under torchrun: films sharded over the ranks]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import configs, parallel, units as _u

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
comm = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{os.environ['LOCAL_RANK']}"))
    comm = parallel.DistComm()

def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()

def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        sync(); t0 = time.perf_counter(); r = fn(); sync(); ts.append(time.perf_counter() - t0)
    return r, float(np.median(ts[1:])) * 1e3

out = {"n_gpus": world}
device, polys = configs.c4_ring_array(8, 5000)
holes = list(device.holes)
model, t_fact = timed(lambda: sc.factorize_model(device=device, current_units="mA", comm=comm))
kw = dict(model=model, applied_fields=[None] * 8, circulating_currents=[{h: 1.0} for h in holes], iterations=5,
          last_only=True, gather=comm is None)
batch, t_solve = timed(lambda: sc.solve_batch(**kw))
films_by_hole = {h.name: film for film, hs in device.holes_by_film().items() for h in hs}
def fluxoids():
    k = 0
    for j in range(8):
        s = batch[j][-1]
        for name in holes:
            if films_by_hole[name] in s.film_solutions:
                s.polygon_fluxoid(polys[name], film=films_by_hole[name], units="Phi_0", with_units=False); k += 1
    return k
nfl, t_flux = timed(fluxoids)
M, t_total = timed(lambda: device.mutual_inductance_matrix(polys, units="pH", iterations=5, comm=comm))
out["c4_ms"] = {"factorize_model": t_fact, "solve_batch_iter5": t_solve, "fluxoids": t_flux, "n_fluxoids": nfl,
                "mutual_inductance_matrix_total": t_total}
# the iteration loop alone, no host results
from superscreen_b200.solver import solve as S
device3, polys3 = configs.c3_susceptometer(4000)
model3, t_fact3 = timed(lambda: sc.factorize_model(device=device3, current_units="uA", circulating_currents={"fc_center": "1 mA"}, comm=comm))
sols3, t_solve3 = timed(lambda: sc.solve(model=model3, iterations=5))
_, t_solve3_nores = timed(lambda: sc.solve(model=model3, iterations=5, return_solutions=False))
out["c3_ms"] = {"factorize_model": t_fact3, "solve_iter5": t_solve3, "solve_iter5_no_solutions": t_solve3_nores}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
