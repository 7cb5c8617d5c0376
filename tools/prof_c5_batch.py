"""Where the time of a 64-field batch on the C5 film goes (host profile + device stage timers)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import configs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
device, fields = configs.c5_large(n)
model = sc.factorize_model(device=device, current_units="uA")
torch.cuda.synchronize()
fl = [sc.ConstantField(float(f)) for f in fields]
sc.solve_batch(model=model, applied_fields=fl[:16]); torch.cuda.synchronize()
t0 = time.perf_counter(); sc.solve_batch(model=model, applied_fields=fl); torch.cuda.synchronize()
print("solve_batch 64:", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable()
sc.solve_batch(model=model, applied_fields=fl); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
# device-only pieces
from superscreen_b200.solver.solve_film import lu_solve, apply_operator
system, info = model.film_systems["film"], model.film_info["film"]
h = torch.randn(len(system.indices), 64, dtype=torch.float64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(2):
    e0.record(); x = lu_solve(system, h); e1.record(); torch.cuda.synchronize()
print("lu_solve 64 rhs: %.2f ms" % e0.elapsed_time(e1))
V = torch.zeros(info.mesh._data.n, 64, dtype=torch.float64, device="cuda"); V[system.indices_dev] = x
for rep in range(2):
    e0.record(); o = apply_operator(info, V, with_sparse=False); e1.record(); torch.cuda.synchronize()
print("apply_operator (dense part) 64 rhs: %.2f ms" % e0.elapsed_time(e1))
for rep in range(2):
    e0.record(); o = apply_operator(info, V[:, :8].contiguous(), with_sparse=False); e1.record(); torch.cuda.synchronize()
print("apply_operator (dense part) 8 rhs: %.2f ms" % e0.elapsed_time(e1))
