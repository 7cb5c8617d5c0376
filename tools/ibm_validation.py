"""Field-coil <-> pickup-loop mutual inductance of the IBM susceptometers against the measured values
(69 +- 7 and 166 +- 4 Phi_0/A; reference docs/notebooks/scanning-squid.ipynb cell 3)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import configs

torch.cuda.set_device(0)
out = {}
for size, exp in (("small", (69, 7)), ("medium", (166, 4))):
    for nv in [int(a) for a in (sys.argv[1:] or ["6000", "12000"])]:
        t0 = time.perf_counter()
        device, rings = configs.ibm_susceptometer(size, nv)
        model = sc.factorize_model(device=device, current_units="uA", circulating_currents={"fc_center": "1 mA"})
        sols = sc.solve(model=model, iterations=5)
        Ms = [sum(s.hole_fluxoid("pl_center", points=rings["pl_center"], with_units=False)) / 1e-3 for s in sols]
        dt = time.perf_counter() - t0
        n_int = {f: len(s.indices) for f, s in model.film_systems.items()}
        out[f"{size}_{nv}"] = {"M_Phi0_per_A_by_iteration": Ms, "measured": exp, "n_interior": n_int, "wall_s": dt}
        print(size, nv, "M by iteration:", ["%.2f" % m for m in Ms], "measured", exp, "n_int", n_int, "%.2fs" % dt, flush=True)
print(json.dumps(out))
