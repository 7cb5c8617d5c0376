"""cProfile of the C4 mutual-inductance call (host-side hot spots), after warm-up."""
import cProfile, io, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import configs
torch.cuda.set_device(0)
device, polys = configs.c4_ring_array(8, 5000)
fn = lambda: device.mutual_inductance_matrix(polys, units="pH", iterations=5)
for _ in range(6):
    fn()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    fn()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25); print(s.getvalue()[:6000])
