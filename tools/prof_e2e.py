"""Host profile of one end-to-end sc.solve on the C2 workload (what bench.py's e2e leg times)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200.geometry import box
from superscreen_b200.synthetic import square_mesh
sites, elements = square_mesh(10.0, 20164, seed=0)
poly = box(10.0, points=4)
def step():
    device = sc.Device("c2", layers=[sc.Layer("layer", Lambda=0.1, z0=0.0)], films=[sc.Polygon("film", layer="layer", points=poly)])
    device.set_meshes({"film": (sites, elements)})
    return sc.solve(device, applied_field=sc.ConstantField(1.0), field_units="mT", current_units="uA")[0]
for _ in range(3): step()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("e2e median %.2f ms" % (np.median(ts) * 1e3))
pr = cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
