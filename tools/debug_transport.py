import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import superscreen_b200 as sc
from oracle import port
from superscreen_b200.solver.solve_film import solve_for_terminal_current_stream, apply_operator, lu_solve
import scipy.linalg as la
g=dict(np.load('tests/golden/transport.npz'))
sites, el = g['in_sites'], g['in_elements']
terminals = {"bar": [sc.Polygon("source", points=g["in_source_polygon"]), sc.Polygon("drain", points=g["in_drain_polygon"])]}
device = sc.Device("bar", layers=[sc.Layer("layer", Lambda=0.4, z0=0.0)], films=[sc.Polygon("bar", layer="layer", points=g["in_film_polygon"])],
                   holes=[sc.Polygon("hole", layer="layer", points=g["in_hole_polygon"])], terminals=terminals)
device.set_meshes({"bar": (sites, el)})
model = sc.factorize_model(device=device, current_units="uA", terminal_currents={"bar": {"source": 10.0, "drain": -10.0}})
info = model.film_info["bar"]; ts = model.terminal_systems["bar"]
mesh = port.build_mesh(sites, el); n=len(sites); w=mesh.vertex_areas; Lam=g['in_Lambda']
Afull = mesh.Q*w[None,:] - Lam[None,:]*mesh.laplacian.toarray()
boundary=g['in_boundary_ordered']; interior=g['in_interior_indices']; hole=g['in_hole_indices']
rel=lambda a,b: np.linalg.norm(a-b)/max(np.linalg.norm(b),1e-300)
# 1. apply_operator with boundary sources
v=np.zeros(n); v[boundary]=np.linspace(-5,5,len(boundary))
out=apply_operator(info, torch.as_tensor(v).cuda(), src_idx=ts.boundary.indices_dev).cpu().numpy()
ref=Afull@v
print('apply_operator boundary src, interior rows:', rel(out[interior], ref[interior]))
# 2. lu_solve with sys_all
h=np.random.default_rng(0).standard_normal(len(interior))
x=lu_solve(ts.film_without_boundary, torch.as_tensor(h).cuda()).cpu().numpy()
xr=la.solve(-Afull[np.ix_(interior,interior)], h)
print('lu_solve with-holes system:', rel(x,xr), 'indices equal', np.array_equal(ts.film_without_boundary.indices, interior))
nh=np.setdiff1d(interior,hole)
h=np.random.default_rng(1).standard_normal(len(nh))
x=lu_solve(model.film_systems["bar"], torch.as_tensor(h).cuda()).cpu().numpy()
xr=la.solve(-Afull[np.ix_(nh,nh)], h)
print('lu_solve no-holes system:', rel(x,xr))
# 3. g_transport
gt=solve_for_terminal_current_stream(device, info, ts, info.terminal_currents).cpu().numpy()
terms={k: np.where(sc.geometry.points_in_polygon(g[f"in_{k}_polygon"], sites[boundary]))[0] for k in ("source","drain")}
film = port.OracleFilm(name="bar", mesh=mesh, z0=0.0, Lambda=Lam, interior_indices=interior, hole_indices={"hole": hole}, boundary_ordered=boundary, terminals=terms)
port.factorize_film(film)
gto=port.solve_for_terminal_current_stream(film, {"source":10.0,"drain":-10.0})
print('g_transport:', rel(gt,gto), 'boundary part', rel(gt[boundary], gto[boundary]), 'interior', rel(gt[interior], gto[interior]))
sol = sc.solve(model=model)[0].film_solutions["bar"]
print('final stream', rel(sol.stream, g['out_current_stream']))
# ---- step-by-step replay of solve_for_terminal_current_stream on the device vs dense numpy ----
from superscreen_b200.solver.utils import stream_from_terminal_current
gg=np.zeros(n); bp=sites[boundary]; cur={"source":10.0,"drain":-10.0}
for t in device.terminals["bar"]:
    ixb=np.sort(t.contains_points(bp, index=True)); rem=boundary[ixb[-1]:]; ixt=boundary[ixb]
    st=stream_from_terminal_current(sites[ixt], -cur[t.name]); gg[ixt[:-1]]+=st; gg[rem]+=st[-1]
gg=gg-gg.max()+np.ptp(gg)/2
print('boundary g equal to oracle', rel(gg[boundary], gto[boundary]), 'nonzero outside boundary', np.abs(np.delete(gg, boundary)).max())
g_dev=torch.as_tensor(gg).cuda()
Ha=-apply_operator(info, g_dev, src_idx=ts.boundary.indices_dev)
Ha_ref=-(Afull@gg)
print('Ha_eff interior', rel(Ha.cpu().numpy()[interior], Ha_ref[interior]))
x1=lu_solve(ts.film_without_boundary, -Ha[ts.film_without_boundary.indices_dev])
x1r=la.solve(-Afull[np.ix_(interior,interior)], -Ha_ref[interior])
print('first solve', rel(x1.cpu().numpy(), x1r))
g_dev[ts.film_without_boundary.indices_dev]=x1
g2=gg.copy(); g2[interior]=x1r
print('g after step 2', rel(g_dev.cpu().numpy(), g2))
