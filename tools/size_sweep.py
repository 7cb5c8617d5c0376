"""Residual of the solve over a range of (small and odd) sizes, symmetric and general LU, 1 and 12 right-hand sides."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200.geometry import box
from superscreen_b200.synthetic import square_mesh
from superscreen_b200.solver.solve_film import apply_operator, lu_solve
worst = 0.0
for n in (40, 90, 160, 300, 700, 1100, 2050, 4200):
    sites, elements = square_mesh(10.0, n, seed=n)
    for mode in ("1", "0"):
        os.environ["SCB_SYMMETRIC"] = mode
        device = sc.Device("sq", layers=[sc.Layer("layer", Lambda=0.3, z0=0.0)], films=[sc.Polygon("film", layer="layer", points=box(10.0, points=4))])
        device.set_meshes({"film": (sites, elements)})
        model = sc.factorize_model(device=device, current_units="uA")
        system, info = model.film_systems["film"], model.film_info["film"]
        ix = system.indices_dev
        for nrhs in (1, 5, 12, 40):
            h = torch.randn(len(system.indices), nrhs, dtype=torch.float64, device="cuda")
            x = lu_solve(system, h)
            V = torch.zeros(info.mesh._data.n, nrhs, dtype=torch.float64, device="cuda"); V[ix] = x
            res = (-apply_operator(info, V, src_idx=ix)[ix] - h).abs().max().item() / h.abs().max().item()
            worst = max(worst, res)
            flag = "" if res < 1e-10 else "   <-- LARGE"
            print(f"n={len(sites):5d} n_int={len(system.indices):5d} n_pad={system.n_pad:5d} sym={mode} nrhs={nrhs:2d} residual {res:.2e}{flag}")
print("worst residual", worst)
