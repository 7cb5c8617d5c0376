"""Timing of the tensor-core kernel-matrix GEMM (matrix-free screening mat-vec, 64 right-hand sides)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import configs
from superscreen_b200.solver.utils import make_film_info
from superscreen_b200.solver.solve_film import apply_operator
for n in (20164, 60000):
    device = configs.c2_square(n)
    info = make_film_info(device=device, vortices=[], circulating_currents={}, terminal_currents={})["film"]
    V = torch.randn(info.mesh._data.n, 64, dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for rep in range(4):
        e0.record(); o = apply_operator(info, V, with_sparse=False); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    nn = info.mesh._data.n
    print(f"SCB_KGEMM_MAXCT={os.environ.get('SCB_KGEMM_MAXCT','8')} n={nn}: 64 rhs {min(ts):.2f} ms = {2*nn*nn*64/min(ts)*1e-9:.1f} TFLOP/s (tensor flops)")
