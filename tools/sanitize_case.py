"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck): a 2k-vertex two-ring
device (mesh build, assembly, symmetric LU with its latency kernels and graph replay, multi-RHS solves,
film-to-film coupling with split sources, field evaluation)
plus a 1.2k-vertex inhomogeneous film through the general and the pivoted factorization."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import configs

torch.cuda.set_device(0)
device, polys = configs.c4_ring_array(2, 2000)
for _ in range(3):  # (the repeated factorizations are captured into CUDA graphs and replayed)
    M = device.mutual_inductance_matrix(polys, units="pH", iterations=2)
model = sc.factorize_model(device=device, current_units="uA", circulating_currents={"hole0": "1 mA"})
sols = sc.solve(model=model, applied_field=sc.ConstantField(0.5), iterations=2)
batch = sc.solve_batch(model=model, applied_fields=[sc.ConstantField(0.1 * k) for k in range(20)], iterations=1)
pos = np.column_stack([np.linspace(-5, 17, 3000), np.linspace(-5, 5, 3000), np.full(3000, 1.0)])
Bz = sols[-1].field_at_position(pos, units="mT", with_units=False)
Bv = sols[-1].screening_field_at_position(pos, vector=True, units="mT", with_units=False)
sq = configs.c2_square(1200, seed=1)
sq.layers["layer"].Lambda = lambda x, y: 0.1 * (1.0 + 0.5 * np.sin(x) * np.cos(y))
for mode in ("0", "1"):
    os.environ["SCB_PIVOT"] = mode
    m2 = sc.factorize_model(device=sq, current_units="uA")
    s2 = sc.solve(model=m2, applied_field=sc.ConstantField(1.0))[0]
torch.cuda.synchronize()
print("sanitize case ok: M00 %.6f pH, |Bz|max %.3e mT, stream max %.3e" % (M[0][0], np.abs(Bz).max(), np.abs(s2.film_solutions["film"].stream).max()))
