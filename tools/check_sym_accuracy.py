import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import configs
from superscreen_b200.solver.solve_film import apply_operator
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20164
device = configs.c2_square(n)
out = {}
for mode in ("0", "1"):
    os.environ["SCB_SYMMETRIC"] = mode
    model = sc.factorize_model(device=device, current_units="uA")
    system, info = model.film_systems["film"], model.film_info["film"]
    s = sc.solve(model=model, applied_field=sc.ConstantField(1.0))[0].film_solutions["film"]
    g = torch.as_tensor(s.stream).cuda(); ix = system.indices_dev
    conv = sc.field_conversion_factor("mT", "uA", "um").magnitude
    res = -apply_operator(info, g, src_idx=ix)[ix] - conv
    out[mode] = s.stream
    print(f"SCB_SYMMETRIC={mode}: sym_scale={'yes' if system.sym_scale is not None else 'no'} residual_inf/h = {float(res.abs().max())/conv:.3e}")
print("rel-L2 stream sym vs general:", np.linalg.norm(out['1']-out['0'])/np.linalg.norm(out['0']))
