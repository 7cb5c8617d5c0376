"""Host profile of Solution.field_at_position on a 1M-point grid (C5-like, smaller film for speed)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import configs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
device, fields = configs.c5_large(n)
sol = sc.solve(device, applied_field=sc.ConstantField(1.0))[0]
grid = configs.evaluation_grid(1000)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    B = sol.field_at_position(grid, units="mT", with_units=False)
    torch.cuda.synchronize(); print("field_at_position 1M:", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable()
B = sol.field_at_position(grid, units="mT", with_units=False)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
