"""sha256 of the symmetric LU factors (workspace M and block inverses) of seeded square films -- printed as
JSON, compared across environment switches (SCB_LU_LAT, SCB_LU_GRAPH ...) by the bit-identity tests.
    python tools/lu_factor_hash.py 2000 5300"""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import _lib
from superscreen_b200.geometry import box
from superscreen_b200.synthetic import square_mesh
from superscreen_b200.solver.solve_film import assemble_negA

L = _lib.lib()
out = {}
for n in [int(a) for a in sys.argv[1:]] or [2000]:
    sites, elements = square_mesh(10.0, n, seed=3)
    device = sc.Device("sq", layers=[sc.Layer("layer", Lambda=0.1, z0=0.0)],
                       films=[sc.Polygon("film", layer="layer", points=box(10.0, points=4))])
    device.set_meshes({"film": (sites, elements)})
    info = sc.solver.utils.make_film_info(device=device, vortices=[], circulating_currents={}, terminal_currents={})["film"]
    info.dev["T"] = None
    ix = torch.as_tensor(info.interior_indices).cuda()
    n_int = len(info.interior_indices); n_pad = -(-n_int // 128) * 128
    M = torch.empty(n_pad, n_pad, dtype=torch.float64, device="cuda")
    M0 = torch.empty_like(M)
    nb = n_pad // 128
    dinv = torch.zeros(int(L.scb_getrf_dinv_bytes(n_pad)) // 8, dtype=torch.float64, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    sym_full = torch.sqrt(info.mesh._data.t["vertex_areas"])
    assemble_negA(info, ix, n_int, n_pad, None, out=M0, sym_scale_full=sym_full)
    hs = []
    for rep in range(3):  # (the third call replays the captured graph when SCB_LU_GRAPH is on)
        M.copy_(M0)
        _lib.check(L.scb_getrf_sym_nopiv(n_pad, _lib.ptr(M), _lib.ptr(dinv), _lib.ptr(flag), _lib.stream_ptr()))
        torch.cuda.synchronize()
        h = hashlib.sha256()
        h.update(M.cpu().numpy().tobytes())
        h.update(dinv[: nb * 2 * 128 * 128].cpu().numpy().tobytes())
        hs.append(h.hexdigest())
    assert int(flag.item()) == 0
    out[str(n)] = hs
print(json.dumps(out))
