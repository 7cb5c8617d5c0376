"""cProfile of the host-bound C3 solve (4 films, iterations=5) after warm-up."""
import cProfile, io, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import configs
device3, polys3 = configs.c3_susceptometer(4000)
model3 = sc.factorize_model(device=device3, current_units="uA", circulating_currents={"fc_center": "1 mA"})
fn = lambda: sc.solve(model=model3, iterations=5)
for _ in range(6):
    fn()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    fn()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45); print(s.getvalue()[:12000])
