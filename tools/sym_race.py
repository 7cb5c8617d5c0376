"""Debug aid: run an LU variant several times on the same matrix and locate run-to-run differences.
usage: sym_race.py n sym(0/1) [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import configs, _lib
from superscreen_b200.solver.solve_film import assemble_negA
from superscreen_b200.solver.utils import make_film_info

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20164
sym = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
device = configs.c2_square(n)
info = make_film_info(device=device, vortices=[], circulating_currents={}, terminal_currents={})["film"]
info.dev["T"] = None
L = _lib.lib()
d = info.mesh._data
sym_full = torch.sqrt(d.t["vertex_areas"])
ix = torch.as_tensor(info.interior_indices).cuda()
n_int = len(info.interior_indices); n_pad = -(-n_int // 128) * 128
S, _ = assemble_negA(info, ix, n_int, n_pad, None, want_margin=True, sym_scale_full=sym_full)
fn = L.scb_getrf_sym_nopiv if sym else L.scb_getrf_nopiv
outs = []
for rep in range(reps):
    M = S.clone()
    dinv = torch.zeros(int(L.scb_getrf_dinv_bytes(n_pad)) // 8, dtype=torch.float64, device=M.device)
    lu_info = torch.zeros(1, dtype=torch.int32, device=M.device)
    _lib.check(fn(n_pad, _lib.ptr(M), _lib.ptr(dinv), _lib.ptr(lu_info), _lib.stream_ptr()))
    torch.cuda.synchronize()
    outs.append(torch.tril(M) if os.environ.get("SCB_SYM_DEBUG") else M)
nb = n_pad // 128
print("checksums (int64 sum of the bit patterns):", [int(o.view(torch.int64).sum().item()) for o in outs])
tag = f"n={n} sym={sym} dbg={os.environ.get('SCB_SYM_DEBUG','0')} la={os.environ.get('SCB_LU_LOOKAHEAD','1')} band={os.environ.get('SCB_LU_BAND','16')}"
for rep in range(1, reps):
    D = (outs[rep] - outs[0]).abs().view(nb, 128, nb, 128).amax(dim=(1, 3)).cpu().numpy()
    bad = np.argwhere(D > 0)
    msg = f"[{tag}] run {rep} vs 0: differing blocks {len(bad)} of {nb*nb}; max diff {D.max():.3e}"
    if len(bad):
        low = bad[bad[:, 0] >= bad[:, 1]]
        if len(low): msg += f" first lower by col: {low[np.lexsort((low[:, 0], low[:, 1]))][:6].tolist()}"
    print(msg)
if not os.environ.get("SCB_SYM_DEBUG"):
    M = outs[0]
    Lf = torch.tril(M, -1) + torch.eye(n_pad, dtype=M.dtype, device=M.device)
    Uf = torch.triu(M)
    rows = torch.arange(0, n_pad, 97, device=M.device)
    print(f"[{tag}] max |LU - S| on sampled rows: {float((Lf[rows] @ Uf - S[rows]).abs().max()):.3e}")
