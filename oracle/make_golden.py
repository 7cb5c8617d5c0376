"""Generates tests/golden/*.npz by executing the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Run from the repo root:  ``python -m oracle.make_golden``

Every array stored under ``out_*`` was produced by the reference's own functions
(superscreen 0.13.0 imported from /root/reference through ``oracle/reference_loader.py``):
``Mesh.from_triangulation`` / ``MeshOperators.from_mesh`` (device/mesh.py:110-155,361-394),
``FilmInfo`` / ``LambdaInfo`` (solver/utils.py:19-132), ``factorize_linear_systems`` and
``solve_film`` (solver/solve_film.py:151-282,440-574), ``biot_savart_film_to_film``
(solver/solve.py:28-73), ``_biot_savart_2d_z/_vector`` (sources/current.py:13-110).
The multi-film driver loop (solver/solve.py:454-549) needs pint + a real Device, so it is
composed here from those live functions in the reference's order (Jacobi: all film-to-film
fields from the previous iterate first, then all re-solves).  Arrays under ``in_*`` are the
inputs both the oracle port and the CUDA path must be fed.
"""
from __future__ import annotations

import itertools
import os
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sp

from oracle.reference_loader import load_reference
from superscreen_b200.geometry import box, circle, points_in_polygon
from superscreen_b200.synthetic import disk_mesh, make_mesh

MU_0 = 1.25663706212e-06
CONV = 1e-3 / MU_0  # mT -> uA/um
VORTEX_FLUX = 2.067833848461929e-15 / MU_0 * 1e12  # Phi_0/mu_0 in uA*um
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _csr(prefix, m, out):
    m = sp.csr_matrix(m).copy()
    m.sort_indices()
    out[f"{prefix}_indptr"] = m.indptr.astype(np.int64)
    out[f"{prefix}_indices"] = m.indices.astype(np.int64)
    out[f"{prefix}_data"] = m.data.astype(np.float64)


def _ref_mesh_outputs(sc, sites, elements, out, tag=""):
    from superscreen import fem
    from superscreen.device import utils as dutils
    from superscreen.device.mesh import Mesh

    mesh = Mesh.from_triangulation(sites, elements)
    ops = mesh.operators
    out[f"out{tag}_boundary_indices"] = mesh.boundary_indices
    edges, is_b = dutils.get_edges(elements)
    out[f"out{tag}_edges"] = edges.astype(np.int64)
    out[f"out{tag}_edge_is_boundary"] = is_b
    out[f"out{tag}_edge_lengths"] = mesh.edge_mesh.edge_lengths
    out[f"out{tag}_edge_centers"] = mesh.edge_mesh.centers
    out[f"out{tag}_edge_directions"] = mesh.edge_mesh.directions
    out[f"out{tag}_triangle_areas"] = mesh.triangle_areas
    out[f"out{tag}_vertex_areas"] = mesh.vertex_areas
    out[f"out{tag}_centroids"] = mesh.triangle_centroids
    adj = fem.adjacency_matrix(elements)
    _csr(f"out{tag}_adjacency", adj, out)
    lil = fem.adj_directed_tri_indices(elements, len(sites)).tolil()
    out[f"out{tag}_star_ptr"] = np.cumsum([0] + [len(r) for r in lil.rows]).astype(np.int64)
    out[f"out{tag}_star_heads"] = np.concatenate([np.asarray(r, dtype=np.int64) for r in lil.rows])
    out[f"out{tag}_star_tris"] = np.concatenate([np.asarray(d, dtype=np.int64) for d in lil.data]) - 1
    _csr(f"out{tag}_laplacian", ops.laplacian, out)
    _csr(f"out{tag}_gradient_x", ops.gradient_x, out)
    _csr(f"out{tag}_gradient_y", ops.gradient_y, out)
    _csr(f"out{tag}_gradient_tri_x", ops.gradient_tri_x, out)
    _csr(f"out{tag}_gradient_tri_y", ops.gradient_tri_y, out)
    out[f"out{tag}_C"] = type(ops).C_vector(sites)
    out[f"out{tag}_Q_diag"] = np.diag(ops.Q).copy()
    return mesh


def _film_info(sc, name, mesh, Lambda, interior, hole_indices, circ, vortices=()):
    from superscreen.solver.utils import FilmInfo, LambdaInfo

    n = len(mesh.sites)
    in_hole = np.zeros(n, dtype=bool)
    for ix in hole_indices.values():
        in_hole[ix] = True
    li = LambdaInfo(film=name, Lambda=Lambda[:, None])
    grad = None
    if li.inhomogeneous:
        grad = np.array([mesh.operators.gradient_x.toarray(), mesh.operators.gradient_y.toarray()])
    return FilmInfo(
        name=name, layer="layer", lambda_info=li, vortices=tuple(vortices),
        interior_indices=interior, boundary_indices=mesh.boundary_indices,
        hole_indices=hole_indices, in_hole=in_hole, circulating_currents=dict(circ),
        weights=mesh.operators.weights.astype(np.float64),
        kernel=mesh.operators.Q.astype(np.float64),
        laplacian=mesh.operators.laplacian.toarray().astype(np.float64), gradient=grad,
    )


def _store_solution(out, key, fs):
    out[f"out_{key}_stream"] = fs.stream
    out[f"out_{key}_J"] = fs.current_density
    out[f"out_{key}_applied_field"] = fs.applied_field
    out[f"out_{key}_self_field"] = fs.self_field
    if fs.field_from_other_films is not None:
        out[f"out_{key}_other"] = fs.field_from_other_films


def golden_ring(sc):
    """C1-like ring: film r=4, hole r=2, mesh disk r=4.4, Lambda=5, z0=0.5."""
    from superscreen.solution import Vortex
    from superscreen.solver.solve_film import factorize_linear_systems, solve_film
    from superscreen.sources.current import _biot_savart_2d_vector, _biot_savart_2d_z

    film_poly, hole_poly = circle(4.0, 64), circle(2.0, 40)
    sites, elements = disk_mesh(4.4, 650, embedded=[film_poly, hole_poly], seed=0)
    out = dict(in_sites=sites, in_elements=elements, in_film_polygon=film_poly,
               in_hole_polygon=hole_poly, in_z0=0.5)
    mesh = _ref_mesh_outputs(sc, sites, elements, out)
    n = len(sites)
    Lambda = np.full(n, 5.0)
    in_film = np.where(points_in_polygon(film_poly, sites))[0]
    interior = np.setdiff1d(in_film, mesh.boundary_indices)
    hole_ix = np.where(points_in_polygon(hole_poly, sites))[0]
    out.update(in_Lambda=Lambda, in_interior_indices=interior, in_hole_indices=hole_ix)
    device = SimpleNamespace(meshes={"ring": mesh}, terminals={})
    vortex = Vortex(x=3.0, y=0.2, film="ring", nPhi0=1)
    out["in_vortex"] = np.array([3.0, 0.2, 1.0])
    cases = {
        "field": dict(circ={}, H=np.full(n, 1.0) * CONV, vort=()),
        "circ": dict(circ={"hole": 1000.0}, H=np.zeros(n), vort=()),
        "vortex": dict(circ={}, H=np.zeros(n), vort=(vortex,)),
    }
    for key, c in cases.items():
        info = _film_info(sc, "ring", mesh, Lambda, interior, {"hole": hole_ix}, c["circ"], c["vort"])
        fsys, hsys, _ = factorize_linear_systems(device, {"ring": info})
        if key == "field":
            out["out_system_indices"] = fsys["ring"].indices
            out["out_A"] = fsys["ring"].A
            out["out_hole_A_rowsum"] = hsys["ring"]["hole"].A.sum(axis=1)
        fs = solve_film(device=device, applied_field=c["H"], film_info=info,
                        film_system=fsys["ring"], hole_systems=hsys["ring"],
                        field_conversion=CONV, vortex_flux=VORTEX_FLUX)
        _store_solution(out, key, fs)
        if key == "circ":
            J = fs.current_density
    # field evaluation from the circulating-current solution (SI inputs, tesla out)
    rng = np.random.default_rng(1)
    ev = np.column_stack([rng.uniform(-6, 6, 300), rng.uniform(-6, 6, 300), rng.uniform(0.8, 3.0, 300)])
    pos3 = np.column_stack([sites, np.full(n, 0.5)]) * 1e-6
    out["in_eval_positions"] = ev
    from scipy.constants import mu_0 as scipy_mu0

    out["in_scipy_mu_0"] = scipy_mu0
    out["out_Bz"] = _biot_savart_2d_z(ev * 1e-6, pos3, J * 1.0, mesh.vertex_areas * 1e-12)
    out["out_Bvec"] = _biot_savart_2d_vector(ev * 1e-6, pos3, J * 1.0, mesh.vertex_areas * 1e-12)
    np.savez_compressed(os.path.join(OUT, "ring.npz"), **out)
    return out


def golden_two_rings(sc):
    """Two stacked rings (z=0 and z=1), I_circ in the lower ring, iterations=3."""
    from superscreen.solver.solve import biot_savart_film_to_film
    from superscreen.solver.solve_film import factorize_linear_systems, solve_film

    out = {}
    films = {}
    specs = {"lower": dict(r_out=3.0, r_in=1.5, z0=0.0, nv=420, seed=1, center=(0.0, 0.0)),
             "upper": dict(r_out=2.0, r_in=1.0, z0=1.0, nv=360, seed=2, center=(0.3, -0.2))}
    meshes, infos = {}, {}
    for name, s in specs.items():
        fp = circle(s["r_out"], 48, center=s["center"])
        hp = circle(s["r_in"], 32, center=s["center"])
        sites, elements = disk_mesh(1.1 * s["r_out"], s["nv"], embedded=[fp, hp], seed=s["seed"], center=s["center"])
        tmp = {}
        mesh = _ref_mesh_outputs(sc, sites, elements, tmp)
        n = len(sites)
        Lambda = np.full(n, 0.08**2 / 0.2 if name == "upper" else 0.5)
        in_film = np.where(points_in_polygon(fp, sites))[0]
        interior = np.setdiff1d(in_film, mesh.boundary_indices)
        hole_ix = np.where(points_in_polygon(hp, sites))[0]
        out.update({f"in_{name}_sites": sites, f"in_{name}_elements": elements,
                    f"in_{name}_Lambda": Lambda, f"in_{name}_interior_indices": interior,
                    f"in_{name}_hole_indices": hole_ix, f"in_{name}_z0": s["z0"],
                    f"in_{name}_film_polygon": fp, f"in_{name}_hole_polygon": hp})
        meshes[name] = mesh
        circ = {f"{name}_hole": 1000.0} if name == "lower" else {}
        infos[name] = _film_info(sc, name, mesh, Lambda, interior, {f"{name}_hole": hole_ix}, circ)
        films[name] = s
    device = SimpleNamespace(meshes=meshes, terminals={})
    fsys, hsys, _ = factorize_linear_systems(device, infos)
    applied = {k: np.full(len(m.sites), 0.2) * CONV for k, m in meshes.items()}
    out["in_applied_mT"] = 0.2
    iterations = 3
    out["in_iterations"] = iterations

    def run(other):
        return {k: solve_film(device=device, applied_field=applied[k], film_info=infos[k],
                              film_system=fsys[k], hole_systems=hsys[k], field_conversion=CONV,
                              vortex_flux=VORTEX_FLUX,
                              field_from_other_films=None if other is None else other[k])
                for k in meshes}

    sols = [run(None)]
    for _ in range(iterations):
        other = {k: np.zeros(len(m.sites)) for k, m in meshes.items()}
        for src, dst in itertools.product(meshes, repeat=2):
            if src == dst:
                continue
            other[dst] += biot_savart_film_to_film(
                film1_sites=meshes[src].sites, film1_z0=films[src]["z0"],
                film1_areas=infos[src].weights, film1_J=sols[-1][src].current_density,
                film2_sites=meshes[dst].sites, film2_z0=films[dst]["z0"])
        sols.append(run(other))
    for it, sol in enumerate(sols):
        for k, fs in sol.items():
            _store_solution(out, f"it{it}_{k}", fs)
    np.savez_compressed(os.path.join(OUT, "two_rings.npz"), **out)
    return out


def golden_square_inhomogeneous(sc):
    """Square film with spatially varying Lambda (exercises the grad-Lambda term),
    plus the 'uniform' and 'inv_euclidean' Laplacian weightings."""
    from superscreen import fem
    from superscreen.solver.solve_film import factorize_linear_systems, solve_film

    sites, elements = make_mesh(box(6.0, points=4), target_vertices=520, seed=3)
    out = dict(in_sites=sites, in_elements=elements)
    mesh = _ref_mesh_outputs(sc, sites, elements, out)
    n = len(sites)
    Lambda = 0.3 * (1.0 + 0.5 * np.cos(sites[:, 0]) * np.sin(0.7 * sites[:, 1]))
    interior = np.setdiff1d(np.arange(n), mesh.boundary_indices)
    out.update(in_Lambda=Lambda, in_interior_indices=interior)
    device = SimpleNamespace(meshes={"sq": mesh}, terminals={})
    info = _film_info(sc, "sq", mesh, Lambda, interior, {}, {})
    assert info.lambda_info.inhomogeneous
    fsys, hsys, _ = factorize_linear_systems(device, {"sq": info})
    out["out_A"] = fsys["sq"].A
    H = (1.0 + 0.1 * sites[:, 0]) * CONV
    out["in_applied_field_solver_units"] = H
    fs = solve_film(device=device, applied_field=H, film_info=info, film_system=fsys["sq"],
                    hole_systems=hsys["sq"], field_conversion=CONV, vortex_flux=VORTEX_FLUX)
    _store_solution(out, "inhom", fs)
    for method in ("uniform", "inv_euclidean"):
        _csr(f"out_laplacian_{method}", fem.laplace_operator(sites, elements, mesh.vertex_areas, method), out)
    np.savez_compressed(os.path.join(OUT, "square_inhomogeneous.npz"), **out)
    return out


def golden_transport(sc):
    """Transport-terminal branch (reference solve_film.py:308-437,505-524,557-562): a 8 x 3 bar with a
    hole, current fed through terminals on the two short edges, plus a uniform field and a
    circulating current.  ``Device.boundary_vertices`` (matplotlib + shapely) is replaced by the
    package's own counter-clockwise boundary ordering, which is an input."""
    from superscreen.solver.solve_film import factorize_linear_systems, solve_film
    from superscreen_b200.mesh import boundary_vertices_ccw

    hole_poly = circle(0.6, 24, center=(0.5, 0.2))
    film_poly = box(8.0, 3.0, points=4)
    sites, elements = make_mesh(film_poly, target_vertices=700, embedded=[hole_poly], seed=4)
    out = dict(in_sites=sites, in_elements=elements, in_film_polygon=film_poly, in_hole_polygon=hole_poly)
    mesh = _ref_mesh_outputs(sc, sites, elements, out)
    n = len(sites)
    Lambda = np.full(n, 0.4)
    boundary = boundary_vertices_ccw(elements)
    in_film = np.where(points_in_polygon(film_poly, sites))[0]
    interior = np.setdiff1d(in_film, boundary)
    hole_ix = np.where(points_in_polygon(hole_poly, sites))[0]
    term_polys = {"source": box(0.2, 2.0, points=4, center=(-4.0, 0.0)),
                  "drain": box(0.2, 2.0, points=4, center=(4.0, 0.0))}
    terminals = [SimpleNamespace(name=k, contains_points=(lambda pts, index=True, poly=v:
                                                         np.where(points_in_polygon(poly, pts))[0]))
                 for k, v in term_polys.items()]
    out.update(in_Lambda=Lambda, in_boundary_ordered=boundary, in_interior_indices=interior, in_hole_indices=hole_ix,
               in_source_polygon=term_polys["source"], in_drain_polygon=term_polys["drain"])
    device = SimpleNamespace(meshes={"bar": mesh}, terminals={"bar": terminals})
    cases = {
        "current": dict(term={"source": 10.0, "drain": -10.0}, circ={}, H=np.zeros(n)),
        "mixed": dict(term={"source": 25.0, "drain": -25.0}, circ={"hole": 3.0}, H=np.full(n, 0.1) * CONV),
        "nocurrent": dict(term={"source": 0.0, "drain": 0.0}, circ={}, H=np.full(n, 0.1) * CONV),
    }
    for key, c in cases.items():
        info = _film_info(sc, "bar", mesh, Lambda, interior, {"hole": hole_ix}, c["circ"])
        info.boundary_indices = boundary
        info.terminal_currents = dict(c["term"])
        fsys, hsys, tsys = factorize_linear_systems(device, {"bar": info})
        if key == "current":
            out["out_system_indices"] = fsys["bar"].indices
            out["out_with_holes_indices"] = tsys["bar"].film_without_boundary.indices
        fs = solve_film(device=device, applied_field=c["H"], film_info=info, film_system=fsys["bar"],
                        hole_systems=hsys["bar"], field_conversion=CONV, vortex_flux=VORTEX_FLUX,
                        terminal_systems=tsys["bar"])
        _store_solution(out, key, fs)
    np.savez_compressed(os.path.join(OUT, "transport.npz"), **out)
    return out


def golden_smooth(sc):
    """Mesh.smooth (device/mesh.py:172-211) on the jittered ring mesh: sites after 1 and after 4 sweeps."""
    from superscreen.device.mesh import Mesh

    ring = np.load(os.path.join(OUT, "ring.npz"))
    sites, elements = ring["in_sites"], ring["in_elements"]
    out = {"in_sites": sites, "in_elements": elements}
    mesh = Mesh.from_triangulation(sites, elements, build_operators=False)
    out["out_sites_1"] = mesh.smooth(1, build_operators=False).sites
    m4 = mesh.smooth(4, build_operators=True)
    out["out_sites_4"] = m4.sites
    out["out_vertex_areas_4"] = m4.vertex_areas
    np.savez_compressed(os.path.join(OUT, "smooth.npz"), **out)
    return out


def main():
    sc = load_reference()
    os.makedirs(OUT, exist_ok=True)
    for fn in (golden_ring, golden_two_rings, golden_square_inhomogeneous, golden_transport, golden_smooth):
        o = fn(sc)
        print(fn.__name__, {k: v.shape for k, v in o.items() if hasattr(v, "shape") and k.startswith("in_") and v.ndim > 0})
    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write(
            "Golden vectors produced by `python -m oracle.make_golden` in the build container from the\n"
            f"unmodified reference superscreen {sc.__version__} (/root/reference), numpy {np.__version__}.\n"
            "`in_*` = inputs, `out_*` = reference outputs.  See oracle/make_golden.py.\n"
        )


if __name__ == "__main__":
    main()
