"""GPU parity tests: the CUDA path (through the C ABI) against the committed golden vectors of
the live reference and against the CPU oracle on identical seeded inputs.

Tolerances (BASELINE.json north_star / SURVEY.md 8d): integer structures bit-exact; local float
arithmetic (areas, CSR data, A) rel-L2 <= 1e-12; solution quantities (stream, J, fields,
fluxoids) rel-L2 <= 1e-8."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL_LOCAL = 1e-12
TOL_SOLUTION = 1e-8


@pytest.fixture(scope="module")
def sc():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import superscreen_b200 as sc

    return sc


def _check_csr(m, g, prefix, tol=TOL_LOCAL):
    """Canonical form on both sides: sorted indices, no explicit zeros.  (The reference assigns
    dense rows into a LIL matrix, which drops entries that are exactly 0.0, fem.py:400-401; the
    CUDA path always stores the full structural pattern adjacency + I.)"""
    m = sp.csr_matrix(m).copy()
    m.sort_indices()
    m.eliminate_zeros()
    r = sp.csr_matrix((g[prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]), shape=m.shape)
    r.sort_indices()
    r.eliminate_zeros()
    assert np.array_equal(m.indptr, r.indptr), prefix
    assert np.array_equal(m.indices, r.indices), prefix
    if tol is not None:
        assert rel_l2(m.data, r.data) <= tol, (prefix, rel_l2(m.data, r.data))


def _check_mesh(sc, g, sites, elements, tag=""):
    mesh = sc.Mesh.from_triangulation(sites, elements)
    # ---- integers: bit-exact ----
    assert mesh.boundary_indices.dtype == np.int64
    assert np.array_equal(mesh.boundary_indices, g[f"out{tag}_boundary_indices"])
    assert np.array_equal(mesh.edge_mesh.edges, g[f"out{tag}_edges"])
    assert np.array_equal(mesh.edge_mesh.is_boundary, g[f"out{tag}_edge_is_boundary"])
    _check_csr(mesh.adjacency_matrix(), g, f"out{tag}_adjacency", tol=None)
    ptr, heads, tris = mesh.directed_star()
    assert np.array_equal(ptr, g[f"out{tag}_star_ptr"])
    assert np.array_equal(heads, g[f"out{tag}_star_heads"])
    assert np.array_equal(tris, g[f"out{tag}_star_tris"])
    # ---- floats ----
    assert rel_l2(mesh.triangle_areas, g[f"out{tag}_triangle_areas"]) <= TOL_LOCAL
    assert rel_l2(mesh.vertex_areas, g[f"out{tag}_vertex_areas"]) <= TOL_LOCAL
    assert rel_l2(mesh.triangle_centroids, g[f"out{tag}_centroids"]) <= TOL_LOCAL
    assert rel_l2(mesh.edge_mesh.edge_lengths, g[f"out{tag}_edge_lengths"]) <= TOL_LOCAL
    assert rel_l2(mesh.edge_mesh.centers, g[f"out{tag}_edge_centers"]) <= TOL_LOCAL
    assert rel_l2(mesh.edge_mesh.directions, g[f"out{tag}_edge_directions"]) <= TOL_LOCAL
    ops = mesh.operators
    for name in ("laplacian", "gradient_x", "gradient_y", "gradient_tri_x", "gradient_tri_y"):
        _check_csr(getattr(ops, name), g, f"out{tag}_{name}")
    # C is ill-conditioned at mesh-boundary vertices on the bounding box (1/(a - x)^2 with
    # a - x ~ 0); those vertices never enter a system, compare on the rest
    interior = np.setdiff1d(np.arange(len(sites)), mesh.boundary_indices)
    assert rel_l2(ops.C[interior], g[f"out{tag}_C"][interior]) <= 1e-11
    assert rel_l2(ops.Q_diagonal[interior], g[f"out{tag}_Q_diag"][interior]) <= 1e-11
    return mesh


def _check_solution(g, key, fs, tol=TOL_SOLUTION):
    errs = dict(
        stream=rel_l2(fs.stream, g[f"out_{key}_stream"]),
        J=rel_l2(fs.current_density, g[f"out_{key}_J"]),
        self_field=rel_l2(fs.self_field, g[f"out_{key}_self_field"]),
    )
    if np.any(g[f"out_{key}_applied_field"]):
        errs["applied"] = rel_l2(fs.applied_field, g[f"out_{key}_applied_field"])
    if f"out_{key}_other" in g:
        errs["other"] = rel_l2(fs.field_from_other_films, g[f"out_{key}_other"])
    assert all(e <= tol for e in errs.values()), (key, errs)
    return errs


def _ring_device(sc, g):
    device = sc.Device(
        "ring", layers=[sc.Layer("layer", Lambda=5.0, z0=float(g["in_z0"]))],
        films=[sc.Polygon("ring", layer="layer", points=g["in_film_polygon"])],
        holes=[sc.Polygon("hole", layer="layer", points=g["in_hole_polygon"])],
    )
    return device


def test_mesh_operators_ring(sc, golden):
    g = golden("ring")
    _check_mesh(sc, g, g["in_sites"], g["in_elements"])


def test_weight_methods(sc, golden):
    g = golden("square_inhomogeneous")
    for method in ("uniform", "inv_euclidean"):
        lap = sc.fem.laplace_operator(g["in_sites"], g["in_elements"], weight_method=method)
        _check_csr(lap, g, f"out_laplacian_{method}")
    with pytest.raises(ValueError):
        sc.fem.laplace_operator(g["in_sites"], g["in_elements"], weight_method="nope")
    # the edge-weight matrices themselves (reference fem.py:124-256), against the oracle
    from oracle import port

    for method in ("uniform", "inv_euclidean", "half_cotangent"):
        w = sp.csr_matrix(sc.fem.calculate_weights(g["in_sites"], g["in_elements"], method))
        ref = sp.csr_matrix(port.calculate_weights(g["in_sites"], g["in_elements"], method))
        ref.eliminate_zeros()
        w.sort_indices(); ref.sort_indices()
        assert np.array_equal(w.indptr, ref.indptr) and np.array_equal(w.indices, ref.indices), method
        assert rel_l2(w.data, ref.data) <= TOL_LOCAL, (method, rel_l2(w.data, ref.data))
    dense = sc.fem.weights_half_cotangent(g["in_sites"], g["in_elements"], sparse=False)
    assert isinstance(dense, np.ndarray) and dense.shape == (len(g["in_sites"]),) * 2
    with pytest.raises(ValueError):
        sc.fem.calculate_weights(g["in_sites"], g["in_elements"], "nope")


def test_ring_solve_against_reference_golden(sc, golden):
    g = golden("ring")
    device = _ring_device(sc, g)
    mesh = _check_mesh(sc, g, g["in_sites"], g["in_elements"])
    device.set_meshes({"ring": mesh})
    model = sc.factorize_model(device=device, current_units="uA")
    info = model.film_info["ring"]
    assert np.array_equal(info.interior_indices, g["in_interior_indices"])
    assert np.array_equal(info.hole_indices["hole"], g["in_hole_indices"])
    system = model.film_systems["ring"]
    assert system.indices.dtype == np.int64
    assert np.array_equal(system.indices, g["out_system_indices"])
    assert rel_l2(system.A, g["out_A"]) <= TOL_LOCAL
    lu, piv = system.lu_piv
    assert lu.shape == g["out_A"].shape and piv.dtype == np.int32
    # uniform field
    sol = sc.solve(model=model, applied_field=sc.ConstantField(1.0), check_inversion=True)
    assert len(sol) == 1
    _check_solution(g, "field", sol[0].film_solutions["ring"])
    # circulating current
    model.set_circulating_currents({"hole": 1000.0})
    sol = sc.solve(model=model)[0]
    _check_solution(g, "circ", sol.film_solutions["ring"])
    # field evaluation kernels (SI in, tesla out) on the circulating-current solution
    from superscreen_b200.solution import biot_savart_2d

    ev = g["in_eval_positions"]
    J = sol.film_solutions["ring"].current_density
    Bz = biot_savart_2d(ev[:, 0], ev[:, 1], ev[:, 2], positions=g["in_sites"], current_densities=J, z0=0.5,
                        areas=mesh.vertex_areas, vector=False)
    Bv = biot_savart_2d(ev[:, 0], ev[:, 1], ev[:, 2], positions=g["in_sites"], current_densities=J, z0=0.5,
                        areas=mesh.vertex_areas, vector=True)
    assert rel_l2(Bz, g["out_Bz"]) <= TOL_SOLUTION
    assert rel_l2(Bv, g["out_Bvec"]) <= TOL_SOLUTION
    # the public API agrees with the kernel (mT out)
    f = sol.field_at_position(ev, units="mT", with_units=False, return_sum=True)
    assert rel_l2(f, g["out_Bz"] * 1e3) <= TOL_SOLUTION
    # vortex
    model.set_circulating_currents({})
    vx, vy, nphi = g["in_vortex"]
    model.set_vortices([sc.Vortex(x=float(vx), y=float(vy), film="ring", nPhi0=float(nphi))])
    sol = sc.solve(model=model)[0]
    _check_solution(g, "vortex", sol.film_solutions["ring"])


def test_two_rings_iterations(sc, golden):
    g = golden("two_rings")
    layers = [sc.Layer("lo", Lambda=float(g["in_lower_Lambda"][0]), z0=float(g["in_lower_z0"])),
              sc.Layer("up", Lambda=float(g["in_upper_Lambda"][0]), z0=float(g["in_upper_z0"]))]
    films = [sc.Polygon("lower", layer="lo", points=g["in_lower_film_polygon"]),
             sc.Polygon("upper", layer="up", points=g["in_upper_film_polygon"])]
    holes = [sc.Polygon("lower_hole", layer="lo", points=g["in_lower_hole_polygon"]),
             sc.Polygon("upper_hole", layer="up", points=g["in_upper_hole_polygon"])]
    device = sc.Device("two", layers=layers, films=films, holes=holes)
    device.set_meshes({k: (g[f"in_{k}_sites"], g[f"in_{k}_elements"]) for k in ("lower", "upper")})
    model = sc.factorize_model(device=device, current_units="uA", circulating_currents={"lower_hole": 1000.0})
    for k in ("lower", "upper"):
        assert np.array_equal(model.film_info[k].interior_indices, g[f"in_{k}_interior_indices"])
        assert np.array_equal(model.film_info[k].hole_indices[f"{k}_hole"], g[f"in_{k}_hole_indices"])
    iterations = int(g["in_iterations"])
    sols = sc.solve(model=model, applied_field=sc.ConstantField(float(g["in_applied_mT"])), iterations=iterations)
    assert len(sols) == iterations + 1
    for it, sol in enumerate(sols):
        for name in ("lower", "upper"):
            _check_solution(g, f"it{it}_{name}", sol.film_solutions[name])
    # function-level drop-in for solver/solve.py:28-73
    from superscreen_b200.solver import biot_savart_film_to_film

    lower, upper = device.meshes["lower"], device.meshes["upper"]
    out = biot_savart_film_to_film(
        film1_sites=lower.sites, film1_z0=0.0, film1_areas=lower.vertex_areas,
        film1_J=g[f"out_it{iterations - 1}_lower_J"], film2_sites=upper.sites, film2_z0=1.0)
    conv = sc.field_conversion_factor("mT", "uA", "um").magnitude
    assert rel_l2(out / conv, g[f"out_it{iterations}_upper_other"]) <= TOL_SOLUTION


def test_square_inhomogeneous_lambda(sc, golden):
    g = golden("square_inhomogeneous")
    sites, elements = g["in_sites"], g["in_elements"]
    mesh = _check_mesh(sc, g, sites, elements)
    from scipy.interpolate import NearestNDInterpolator

    lam = NearestNDInterpolator(sites, g["in_Lambda"])  # exact at the mesh sites
    from superscreen_b200.geometry import box

    device = sc.Device("sq", layers=[sc.Layer("layer", Lambda=lambda x, y: lam(x, y), z0=0.0)],
                       films=[sc.Polygon("sq", layer="layer", points=box(6.0, points=4))])
    device.set_meshes({"sq": mesh})
    model = sc.factorize_model(device=device, current_units="uA")
    assert model.film_info["sq"].lambda_info.inhomogeneous
    assert model.film_systems["sq"].sym_scale is None  # grad-Lambda term: general (unsymmetric) LU
    assert np.array_equal(model.film_systems["sq"].indices, g["in_interior_indices"])
    assert rel_l2(model.film_systems["sq"].A, g["out_A"]) <= TOL_LOCAL
    conv = sc.field_conversion_factor("mT", "uA", "um").magnitude
    H = g["in_applied_field_solver_units"] / conv
    field = lambda x, y, z: NearestNDInterpolator(sites, H)(x, y)
    sol = sc.solve(model=model, applied_field=field)[0]
    _check_solution(g, "inhom", sol.film_solutions["sq"])


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("metric", ["euclidean", "sqeuclidean"])
def test_cdist_matches_scipy(sc, dim, metric):
    """The reference's only pinned hot-path kernel test (test_distance.py:27-37)."""
    from scipy.spatial import distance

    rng = np.random.default_rng(1)
    XA, XB = rng.random((123, dim)), rng.random((77, dim))
    out = sc.distance.cdist(XA, XB, metric=metric)
    assert out.shape == (123, 77)
    assert np.allclose(out, distance.cdist(XA, XB, metric=metric), rtol=1e-14, atol=1e-15)
    with pytest.raises(ValueError):
        sc.distance.cdist(XA, XB, metric="invalid")


def test_q_matrix_function(sc):
    from oracle import port

    rng = np.random.default_rng(0)
    pts = rng.random((300, 2))
    assert rel_l2(sc.distance.q_matrix(pts), port.q_matrix(pts)) <= TOL_LOCAL


@pytest.mark.parametrize("n_target,seed", [(3000, 11), (6000, 12)])
def test_medium_square_against_oracle(sc, n_target, seed):
    """Seeded C2-shaped case at a size the oracle finishes in seconds; exercises multi-block LU
    (n_int spans many 128-blocks, plus identity padding)."""
    from oracle import port
    from superscreen_b200.geometry import box
    from superscreen_b200.synthetic import square_mesh

    sites, elements = square_mesh(10.0, n_target, seed=seed)
    device = sc.Device("sq", layers=[sc.Layer("layer", Lambda=0.1, z0=0.0)],
                       films=[sc.Polygon("film", layer="layer", points=box(10.0, points=4))])
    device.set_meshes({"film": (sites, elements)})
    sol = sc.solve(device, applied_field=sc.ConstantField(1.0), check_inversion=True)[0]
    fs = sol.film_solutions["film"]
    om = port.build_mesh(sites, elements)
    interior = np.setdiff1d(np.arange(len(sites)), om.boundary_indices)
    film = port.factorize_film(port.OracleFilm(name="film", mesh=om, z0=0.0, Lambda=np.full(len(sites), 0.1),
                                               interior_indices=interior, hole_indices={}))
    conv = port.field_conversion_mT_to_uA_per_um()
    ref = port.solve_film(film, np.full(len(sites), conv), {}, conv)
    errs = dict(stream=rel_l2(fs.stream, ref.stream), J=rel_l2(fs.current_density, ref.current_density),
                self_field=rel_l2(fs.self_field, ref.self_field), total=rel_l2(fs.total_field, ref.total_field))
    assert all(e <= TOL_SOLUTION for e in errs.values()), errs


def test_full_size_c2_properties(sc):
    """BASELINE config 2 (20k-vertex square): size-independent properties instead of the oracle:
    residual of the linear system through the matrix-free operator, linearity in the applied
    field, Meissner screening (total field << applied deep inside) and lumped-mass invariant."""
    import torch

    from superscreen_b200.geometry import box
    from superscreen_b200.solver.solve_film import apply_operator
    from superscreen_b200.synthetic import square_mesh

    sites, elements = square_mesh(10.0, 20164, seed=0)
    device = sc.Device("sq", layers=[sc.Layer("layer", Lambda=0.1, z0=0.0)],
                       films=[sc.Polygon("film", layer="layer", points=box(10.0, points=4))])
    device.set_meshes({"film": (sites, elements)})
    mesh = device.meshes["film"]
    assert abs(mesh.vertex_areas.sum() - mesh.triangle_areas.sum()) <= 1e-12 * 100.0
    assert abs(mesh.triangle_areas.sum() - 100.0) <= 1e-9
    model = sc.factorize_model(device=device, current_units="uA")
    system, info = model.film_systems["film"], model.film_info["film"]
    assert float(system.margin.min().item()) > 0, "system must be row-diagonally dominant"
    s1 = sc.solve(model=model, applied_field=sc.ConstantField(1.0))[0].film_solutions["film"]
    s2 = sc.solve(model=model, applied_field=sc.ConstantField(2.5))[0].film_solutions["film"]
    assert rel_l2(s2.stream, 2.5 * s1.stream) <= 1e-12
    # residual || (-A) g - h ||_inf / ||h||_inf  (reference reaches 4e-13..8e-12)
    d = mesh._data
    g = torch.as_tensor(s1.stream).to(d.device)
    ix = system.indices_dev
    conv = sc.field_conversion_factor("mT", "uA", "um").magnitude
    res = -apply_operator(info, g, src_idx=ix)[ix] - conv
    assert float(res.abs().max().item()) / conv <= 1e-10
    # screening: |total field| in the centre is far below the applied 1 mT for Lambda << size
    centre = np.linalg.norm(sites, axis=1) < 2.0
    assert np.abs(s1.total_field[centre]).mean() < 0.2


def test_non_delaunay_mesh_uses_refinement(sc, caplog):
    """Vertex-perturbed (non-Delaunay) mesh: cotangent weights of interior edges can be negative, so
    the row-dominance bound fails (SURVEY.md Q11 caveat).  The unpivoted LU + iterative refinement
    must still reproduce the oracle's pivoted LAPACK solution."""
    from oracle import port
    from superscreen_b200.geometry import box
    from superscreen_b200.synthetic import square_mesh

    sites, elements = square_mesh(10.0, 2500, seed=21)
    rng = np.random.default_rng(5)
    om0 = port.build_mesh(sites, elements, with_Q=False)
    interior = np.setdiff1d(np.arange(len(sites)), om0.boundary_indices)
    h = np.sqrt(100.0 / len(sites))
    moved = sites.copy()
    moved[interior] += 0.42 * h * (rng.random((len(interior), 2)) * 2 - 1)
    # keep only perturbations that leave every triangle counter-clockwise
    p = moved[elements]
    cross = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    bad_vertices = np.unique(elements[cross <= 0.05 * h * h])
    moved[bad_vertices] = sites[bad_vertices]
    p = moved[elements]
    cross = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    assert (cross > 0).all()
    device = sc.Device("sq", layers=[sc.Layer("layer", Lambda=2.0, z0=0.0)],
                       films=[sc.Polygon("film", layer="layer", points=box(10.0, points=4))])
    device.set_meshes({"film": (moved, elements)})
    model = sc.factorize_model(device=device, current_units="uA")
    system = model.film_systems["film"]
    sol = sc.solve(model=model, applied_field=sc.ConstantField(1.0), check_inversion=True)[0]
    fs = sol.film_solutions["film"]
    om = port.build_mesh(moved, elements)
    film = port.factorize_film(port.OracleFilm(name="film", mesh=om, z0=0.0, Lambda=np.full(len(sites), 2.0),
                                               interior_indices=interior, hole_indices={}))
    conv = port.field_conversion_mT_to_uA_per_um()
    ref = port.solve_film(film, np.full(len(sites), conv), {}, conv)
    assert rel_l2(fs.stream, ref.stream) <= TOL_SOLUTION, (rel_l2(fs.stream, ref.stream), system.refine)
    assert rel_l2(fs.current_density, ref.current_density) <= TOL_SOLUTION
    assert rel_l2(fs.total_field, ref.total_field) <= TOL_SOLUTION


def test_transport_terminals_against_reference_golden(sc, golden):
    """Transport-terminal branch against the live-reference golden."""
    from superscreen_b200.mesh import boundary_vertices_ccw

    g = golden("transport")
    sites, elements = g["in_sites"], g["in_elements"]
    terminals = {"bar": [sc.Polygon("source", points=g["in_source_polygon"]),
                         sc.Polygon("drain", points=g["in_drain_polygon"])]}
    device = sc.Device("bar", layers=[sc.Layer("layer", Lambda=float(g["in_Lambda"][0]), z0=0.0)],
                       films=[sc.Polygon("bar", layer="layer", points=g["in_film_polygon"])],
                       holes=[sc.Polygon("hole", layer="layer", points=g["in_hole_polygon"])], terminals=terminals)
    device.set_meshes({"bar": (sites, elements)})
    assert np.array_equal(device.boundary_vertices("bar"), g["in_boundary_ordered"])
    assert np.array_equal(boundary_vertices_ccw(elements), g["in_boundary_ordered"])
    cases = {
        "current": dict(term={"bar": {"source": "10 uA", "drain": "-10 uA"}}, circ=None, field=None),
        "mixed": dict(term={"bar": {"source": 25.0, "drain": -25.0}}, circ={"hole": 3.0}, field=sc.ConstantField(0.1)),
        "nocurrent": dict(term={"bar": {"source": 0.0, "drain": 0.0}}, circ=None, field=sc.ConstantField(0.1)),
    }
    for key, c in cases.items():
        model = sc.factorize_model(device=device, current_units="uA", terminal_currents=c["term"],
                                   circulating_currents=c["circ"])
        info = model.film_info["bar"]
        assert np.array_equal(info.interior_indices, g["in_interior_indices"])
        assert np.array_equal(model.film_systems["bar"].indices, g["out_system_indices"])
        ts = model.terminal_systems["bar"]
        assert np.array_equal(ts.film_without_boundary.indices, g["out_with_holes_indices"])
        assert ts.film_without_boundary_or_holes is model.film_systems["bar"]
        sol = sc.solve(model=model, applied_field=c["field"], check_inversion=True)[0]
        _check_solution(g, key, sol.film_solutions["bar"])
    # current conservation is enforced like the reference (solve.py:260-264)
    with pytest.raises(ValueError):
        sc.factorize_model(device=device, current_units="uA", terminal_currents={"bar": {"source": 1.0, "drain": 0.0}})


def test_static_c_vector_and_q_matrix(sc):
    """MeshOperators.C_vector / Q_matrix as static functions of (points, weights) (device/mesh.py:400-458)."""
    from oracle import port

    rng = np.random.default_rng(7)
    pts = rng.random((400, 2)) * np.array([3.0, 2.0])
    w = 0.01 + rng.random(400)
    C = sc.MeshOperators.C_vector(pts)
    Cref = port.C_vector(pts)
    inner = (np.abs(pts[:, 0] - pts[:, 0].mean()) < 0.45 * np.ptp(pts[:, 0])) & \
            (np.abs(pts[:, 1] - pts[:, 1].mean()) < 0.45 * np.ptp(pts[:, 1]))
    assert rel_l2(C[inner], Cref[inner]) <= 1e-12
    Q = sc.MeshOperators.Q_matrix(pts, w)
    Qref = port.Q_matrix(pts, w)
    off = ~np.eye(400, dtype=bool)
    assert rel_l2(Q[off], Qref[off]) <= TOL_LOCAL
    assert rel_l2(np.diag(Q)[inner], np.diag(Qref)[inner]) <= 1e-11


def test_symmetric_factorization_matches_general(sc, monkeypatch):
    """Constant-Lambda films are factored through the diagonally similar symmetric form
    S = D (-A) D^-1 (half the flops).  It must agree with the general factorization far inside
    the solution tolerance, be bit-reproducible run to run, and still expose (lu, piv) of -A."""
    import torch

    from superscreen_b200.geometry import box
    from superscreen_b200.synthetic import square_mesh

    sites, elements = square_mesh(10.0, 6000, seed=5)

    def run(mode):
        monkeypatch.setenv("SCB_SYMMETRIC", mode)
        device = sc.Device("sq", layers=[sc.Layer("layer", Lambda=0.1, z0=0.0)],
                           films=[sc.Polygon("film", layer="layer", points=box(10.0, points=4))])
        device.set_meshes({"film": (sites, elements)})
        model = sc.factorize_model(device=device, current_units="uA")
        sol = sc.solve(model=model, applied_field=sc.ConstantField(1.0))[0].film_solutions["film"]
        return model.film_systems["film"], sol

    sys_g, sol_g = run("0")
    sys_s, sol_s = run("1")
    sys_s2, sol_s2 = run("1")
    assert sys_g.sym_scale is None and sys_s.sym_scale is not None
    assert rel_l2(sol_s.stream, sol_g.stream) <= 1e-11
    assert rel_l2(sol_s.current_density, sol_g.current_density) <= 1e-10
    assert torch.equal(sys_s.lu, sys_s2.lu), "symmetric LU must be bit-reproducible"
    assert np.array_equal(sol_s.stream, sol_s2.stream)
    # host view: unit-lower L and U of the plain -A (no row exchanges)
    lu, piv = sys_s.lu_piv
    assert np.array_equal(piv, np.arange(len(piv), dtype=np.int32))
    Lf = np.tril(lu, -1) + np.eye(len(lu))
    assert rel_l2(Lf @ np.triu(lu), -sys_s.A) <= 1e-12



def test_mesh_smooth_against_reference_golden(sc, golden):
    """Mesh.smooth on the device (scb_mesh_smooth) against the reference's Mesh.smooth: the
    summation order is reproduced, so the smoothed coordinates are bit-identical."""
    g = golden("smooth")
    mesh = sc.Mesh.from_triangulation(g["in_sites"], g["in_elements"], build_operators=False)
    assert mesh.smooth(0) is mesh
    assert np.array_equal(mesh.smooth(1, build_operators=False).sites, g["out_sites_1"])
    m4 = mesh.smooth(4)
    assert np.array_equal(m4.sites, g["out_sites_4"])
    assert np.array_equal(m4.boundary_indices, mesh.boundary_indices)
    assert rel_l2(m4.vertex_areas, g["out_vertex_areas_4"]) <= TOL_LOCAL
    assert m4.operators is not None
