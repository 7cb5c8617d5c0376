import cProfile, pstats, sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import configs
device, polys = configs.c4_ring_array(8, 5000)
M = device.mutual_inductance_matrix(polys, units="pH", iterations=5)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
M = device.mutual_inductance_matrix(polys, units="pH", iterations=5)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
