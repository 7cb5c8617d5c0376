"""Per-repetition wall times of the C4 / C2 host paths with allocator counters (jitter hunt)."""
import gc, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import configs
torch.cuda.set_device(0)

def stats():
    s = torch.cuda.memory_stats()
    h = {}
    try:
        h = torch.cuda.host_memory_stats()
    except Exception:
        pass
    return s.get("num_device_alloc", 0), s.get("num_device_free", 0), h.get("num_host_alloc", -1), h.get("num_host_free", -1)

def run(label, fn, reps=12):
    out = []
    for r in range(reps):
        torch.cuda.synchronize(); a = stats(); t0 = time.perf_counter(); x = fn(); torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3; b = stats()
        out.append(f"{dt:.1f}ms(dev+{b[0]-a[0]}/-{b[1]-a[1]} host+{b[2]-a[2]}/-{b[3]-a[3]})")
        del x
    print(label, " ".join(out), flush=True)

device, polys = configs.c4_ring_array(8, 5000)
holes = list(device.holes)
model = sc.factorize_model(device=device, current_units="mA")
kw = dict(model=model, applied_fields=[None] * 8, circulating_currents=[{h: 1.0} for h in holes], iterations=5, last_only=True)
run("c4 solve_batch", lambda: sc.solve_batch(**kw))
run("c4 factorize", lambda: sc.factorize_model(device=device, current_units="mA"))
run("c4 M", lambda: device.mutual_inductance_matrix(polys, units="pH", iterations=5))
gc.disable()
run("c4 M nogc", lambda: device.mutual_inductance_matrix(polys, units="pH", iterations=5))
gc.enable()
from superscreen_b200.geometry import box
from superscreen_b200.synthetic import square_mesh
sites, elements = square_mesh(10.0, 20164, seed=0)
def e2e():
    d = sc.Device("c2", layers=[sc.Layer("layer", Lambda=0.1, z0=0.0)], films=[sc.Polygon("film", layer="layer", points=box(10.0, points=4))])
    d.set_meshes({"film": (sites, elements)})
    return sc.solve(d, applied_field=sc.ConstantField(1.0))[0]
run("c2 e2e", e2e, reps=8)
