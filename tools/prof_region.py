"""One warm call, then the same call between cudaProfilerStart/Stop (for `ncu --profile-from-start off`).
python tools/prof_region.py c4|c3|c4solve"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import configs

which = sys.argv[1]
torch.cuda.set_device(0)
if which.startswith("c4"):
    device, polys = configs.c4_ring_array(8, 5000)
    fn = lambda: device.mutual_inductance_matrix(polys, units="pH", iterations=5)
elif which == "c3":
    device, polys = configs.c3_susceptometer(4000)
    model = sc.factorize_model(device=device, current_units="uA", circulating_currents={"fc_center": "1 mA"})
    fn = lambda: sc.solve(model=model, iterations=5)
fn(); fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
t0 = time.perf_counter()
fn()
torch.cuda.synchronize()
print("region wall ms", (time.perf_counter() - t0) * 1e3)
torch.cuda.cudart().cudaProfilerStop()
