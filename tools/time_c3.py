"""Median timings of the C3 (4-film susceptometer) factorization and 5-iteration solve."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import configs
device, polys = configs.c3_susceptometer(4000)
tf, ts = [], []
for rep in range(15):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    model = sc.factorize_model(device=device, current_units="uA", circulating_currents={"fc_center": "1 mA"})
    torch.cuda.synchronize(); t1 = time.perf_counter()
    sols = sc.solve(model=model, iterations=5)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    tf.append(t1 - t0); ts.append(t2 - t1)
print(f"SCB_FILM_STREAMS={os.environ.get('SCB_FILM_STREAMS', '8')}: factorize median {np.median(tf[3:])*1e3:.2f} ms (min {min(tf)*1e3:.2f}), "
      f"solve(iter=5) median {np.median(ts[3:])*1e3:.2f} ms (min {min(ts)*1e3:.2f})")
