"""Multi-GPU check (run under torchrun): film sharding + J exchange over NCCL and sharded
field_at_position must reproduce the single-process result bit for bit / to rounding.

torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/run_sharded.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import superscreen_b200 as sc
from superscreen_b200.geometry import circle
from superscreen_b200.synthetic import disk_mesh

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{os.environ['LOCAL_RANK']}"))
nfilms = int(os.environ.get("NFILMS", "4"))
layers = [sc.Layer("base", Lambda=0.5, z0=0.0)]
films, holes, meshes, polys = [], [], {}, {}
for k in range(nfilms):
    c = (12.0 * (k % 4), 12.0 * (k // 4))
    fp, hp = circle(4.0, 64, center=c), circle(2.0, 40, center=c)
    films.append(sc.Polygon(f"ring{k}", layer="base", points=fp))
    holes.append(sc.Polygon(f"hole{k}", layer="base", points=hp))
    meshes[f"ring{k}"] = disk_mesh(4.4, 1500, embedded=[fp, hp], seed=k, center=c)
    polys[f"hole{k}"] = circle(3.0, 101, center=c)
device = sc.Device("array", layers=layers, films=films, holes=holes)
device.set_meshes(meshes)
comm = sc.parallel.DistComm()
cc = {"hole0": "1 mA"}
model = sc.factorize_model(device=device, current_units="uA", circulating_currents=cc, comm=comm)
assert set(model.film_systems) == {f"ring{k}" for k in range(nfilms) if k % world == rank}
sols = sc.solve(model=model, applied_field=sc.ConstantField(0.1), iterations=3)
single = sc.solve(device, applied_field=sc.ConstantField(0.1), circulating_currents=cc, iterations=3)
err = 0.0
for a, b in zip(sols, single):
    for name in device.films:
        fa, fb = a.film_solutions[name], b.film_solutions[name]
        err = max(err, np.linalg.norm(fa.stream - fb.stream) / np.linalg.norm(fb.stream))
        err = max(err, np.linalg.norm(fa.total_field - fb.total_field) / np.linalg.norm(fb.total_field))
M = device.mutual_inductance_matrix(polys, units="pH", iterations=2, comm=comm)
M1 = device.mutual_inductance_matrix(polys, units="pH", iterations=2)
xs = np.linspace(-6, 42, 200)
pos = np.array([(x, y) for x in xs for y in np.linspace(-6, 18, 50)])
f = sc.parallel.field_at_position_sharded(sols[-1], pos, zs=1.0, comm=comm, units="mT")
f1 = single[-1].field_at_position(pos, zs=1.0, units="mT", with_units=False)
errM = np.abs(M - M1).max() / np.abs(M1).max()
errF = np.linalg.norm(f - f1) / np.linalg.norm(f1)
print(f"rank {rank}/{world}: sharded-vs-single stream/field {err:.2e}  M {errM:.2e}  field_at_position {errF:.2e}", flush=True)
assert err < 1e-12 and errM < 1e-10 and errF < 1e-12
dist.barrier()
dist.destroy_process_group()
