"""CPU-only tests: host logic, units, geometry, and the C-ABI surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

import superscreen_b200 as sc
from superscreen_b200 import _lib, units
from superscreen_b200.geometry import box, circle, points_in_polygon
from superscreen_b200.synthetic import disk_mesh, square_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g

    g.build()
    lib = _lib.load_library()
    header = open(os.path.join(ROOT, "include", "scb.h")).read()
    declared = set(re.findall(r"\b(scb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes found in include/scb.h"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in scb.h but not exported"
    assert set(_lib.EXPORTS) == declared
    assert lib.scb_version() >= 100
    assert lib.scb_getrf_dinv_bytes(256) >= (2 * 2 * 128 * 128 + 2 * 256 * 128) * 8
    assert lib.scb_mesh_workspace_elems(10, 20) > 0


def test_compute_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    sites, elements = square_mesh(4.0, 200)
    with pytest.raises(_lib.SCBError):
        sc.Mesh.from_triangulation(sites, elements)


def test_units():
    assert np.isclose(units.conversion_factor("mT", "T"), 1e-3)
    assert np.isclose(units.conversion_factor("uA / um", "A / m"), 1.0)
    assert np.isclose(units.conversion_factor("mT * um ** 2", "Phi_0"), 1e-15 / units.PHI_0)
    assert np.isclose(units.conversion_factor("Phi_0 / A", "pH"), units.PHI_0 * 1e12)
    assert np.isclose(sc.field_conversion_factor("mT", "uA", "um").magnitude, 1e-3 / units.MU_0)
    assert np.isclose(sc.field_conversion_factor("A / m", "uA", "um").magnitude, 1.0)
    assert np.isclose(sc.convert_field(1.0, "mT", old_units="uA / um", with_units=False), units.MU_0 * 1e3)
    assert np.isclose(sc.convert_field(np.ones(3), "uT", old_units="mT", with_units=False)[0], 1e3)
    assert np.isclose(units.to_quantity("1 mA", "uA").to("uA").magnitude, 1000.0)
    with pytest.raises(units.DimensionalityError):
        units.conversion_factor("mT", "um")


def test_point_in_polygon_and_mesh_generators():
    sq = box(2.0, points=4)
    q = np.array([[0, 0], [0.99, 0.99], [1.01, 0], [5, 5]])
    assert points_in_polygon(sq, q).tolist() == [True, True, False, False]
    sites, elements = disk_mesh(4.4, 1200, embedded=[circle(4, 64), circle(2, 40)], seed=0)
    p = sites[elements]
    cross = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    assert (cross > 0).all(), "all triangles must be CCW"
    assert np.unique(elements).size == len(sites)
    s2, e2 = disk_mesh(4.4, 1200, embedded=[circle(4, 64), circle(2, 40)], seed=0)
    assert np.array_equal(sites, s2) and np.array_equal(elements, e2), "generator must be deterministic"


def test_device_model_and_error_paths():
    layer = sc.Layer("base", london_lambda=0.5, thickness=0.05, z0=0.5)
    assert np.isclose(layer.Lambda, 5.0)
    film = sc.Polygon("ring", layer="base", points=circle(4, 64))
    hole = sc.Polygon("hole", layer="base", points=circle(2, 40))
    dev = sc.Device("d", layers=[layer], films=[film], holes=[hole])
    assert [h.name for h in dev.holes_by_film()["ring"]] == ["hole"]
    assert dev.solve_dtype == np.float64
    with pytest.raises(ValueError):
        sc.solve()  # neither model nor device
    with pytest.raises(ValueError):
        sc.solve(dev)  # no mesh
    with pytest.raises(TypeError):
        sc.solve(model=object())
    with pytest.raises(ValueError):
        sc.Device("bad", layers=[layer], films=[sc.Polygon("f", layer="nope", points=circle(1, 16))])
    with pytest.raises(ValueError):
        sc.Layer("x")


def test_index_sets_reproduce_golden(golden):
    """Mesh vertices lie exactly ON the film / hole polygons, so the point-in-polygon arithmetic must
    be reproducible bit for bit: the golden index sets are inputs of both the oracle and the CUDA path."""
    g = golden("ring")
    hole = np.where(points_in_polygon(g["in_hole_polygon"], g["in_sites"]))[0]
    assert np.array_equal(hole, g["in_hole_indices"])
    g = golden("two_rings")
    for k in ("lower", "upper"):
        hole = np.where(points_in_polygon(g[f"in_{k}_hole_polygon"], g[f"in_{k}_sites"]))[0]
        assert np.array_equal(hole, g[f"in_{k}_hole_indices"])
        film = np.where(points_in_polygon(g[f"in_{k}_film_polygon"], g[f"in_{k}_sites"]))[0]
        assert set(g[f"in_{k}_interior_indices"]).issubset(set(film))


def test_linear_tri_interpolate_matches_bruteforce():
    from oracle import port
    from superscreen_b200.solution import linear_tri_interpolate

    rng = np.random.default_rng(0)
    sites, el = disk_mesh(4.4, 1500, embedded=[circle(4, 64), circle(2, 40)])
    vals = np.column_stack([np.sin(sites[:, 0]), sites[:, 1] ** 2])
    pts = np.concatenate([circle(3.0, 201), rng.uniform(-5, 5, (200, 2))])
    a = linear_tri_interpolate(sites, el, vals, pts)
    b = port.linear_tri_interp(sites, el, vals, pts)
    m = np.isfinite(b).all(axis=1)
    assert np.array_equal(np.isfinite(a).all(axis=1), m)
    assert np.abs(a[m] - b[m]).max() <= 1e-13


def test_memoised_host_queries_and_cheap_copies():
    """The memoised point location / point-in-polygon queries and the constructor-free Polygon.copy
    used by the batched post-processing return exactly what the direct computation returns."""
    from superscreen_b200.solution import linear_tri_interpolate

    poly = sc.Polygon("p", layer="l", points=circle(2.0, 40))
    cp = poly.copy()
    assert cp is not poly and cp.name == "p" and cp.layer == "l"
    assert np.array_equal(cp.points, poly.points) and cp.points is not poly.points
    rng = np.random.default_rng(0)
    for m in (5, 300, 2000):  # below, inside and above the memoisation thresholds
        pts = rng.uniform(-3, 3, (m, 2))
        direct = points_in_polygon(poly.points, pts)
        assert np.array_equal(poly.contains_points(pts), direct)
        assert np.array_equal(cp.contains_points(pts), direct)          # served from the memo
        assert np.array_equal(poly.contains_points(pts, index=True), np.where(direct)[0])
    sites, elements = disk_mesh(3.0, 400, seed=2)
    xy = rng.uniform(-2, 2, (50, 2))
    v1, v2 = rng.standard_normal(len(sites)), rng.standard_normal((len(sites), 2))
    a = linear_tri_interpolate(sites, elements, v1, xy)
    b = linear_tri_interpolate(sites, elements, v1, xy)                  # second call: memoised location
    c = linear_tri_interpolate(sites, elements, v2, xy)
    assert np.array_equal(a, b) and c.shape == (50, 2)
    # linear functions are reproduced exactly (up to rounding) by piecewise-linear interpolation
    lin = 0.3 * sites[:, 0] - 1.7 * sites[:, 1] + 0.5
    assert np.allclose(linear_tri_interpolate(sites, elements, lin, xy), 0.3 * xy[:, 0] - 1.7 * xy[:, 1] + 0.5,
                       atol=1e-12)


def test_with_lambda_shares_meshes():
    """configs.with_lambda: a new Device with another Lambda around the same mesh objects (no GPU needed
    to build it when the meshes dict is given)."""
    from superscreen_b200 import configs

    device = sc.Device("square", layers=[sc.Layer("layer", Lambda=0.1, z0=0.25)],
                       films=[sc.Polygon("film", layer="layer", points=box(10.0, points=4))])
    sentinel = object()
    device.meshes = {"film": sentinel}
    other = configs.with_lambda(device, 0.4)
    assert other.layers["layer"].Lambda == 0.4 and other.layers["layer"].z0 == 0.25
    assert device.layers["layer"].Lambda == 0.1
    assert other.meshes["film"] is sentinel and list(other.films) == ["film"]


def test_solution_equality_helpers():
    """FilmSolution.is_close / == (tolerance-based like the reference) and Solution.equals."""
    from superscreen_b200.solution import FilmSolution, Solution

    rng = np.random.default_rng(0)
    n = 50
    base = dict(stream=rng.standard_normal(n), current_density=rng.standard_normal((n, 2)),
                applied_field=rng.standard_normal(n), self_field=rng.standard_normal(n))
    a, b = FilmSolution(**base), FilmSolution(**{k: v.copy() for k, v in base.items()})
    assert a == b and a.is_close(b)
    c = FilmSolution(**{**base, "stream": base["stream"] + 1e-2})
    assert a != c and not a.is_close(c)
    d = FilmSolution(**base, field_from_other_films=np.zeros(n))
    assert a != d and a != "not a solution"
    device = sc.Device("d", layers=[sc.Layer("l", Lambda=1.0, z0=0.0)],
                       films=[sc.Polygon("film", layer="l", points=box(2.0, points=4))])
    f = sc.ConstantField(1.0)
    kw = dict(device=device, applied_field_func=f, field_units="mT", current_units="uA")
    s1, s2 = Solution(film_solutions={"film": a}, **kw), Solution(film_solutions={"film": b}, **kw)
    assert s1.equals(s1) and s1.equals(s2) and not s1.equals(Solution(film_solutions={"film": c}, **kw))
    assert not s1.equals(Solution(film_solutions={"film": a}, **{**kw, "field_units": "uT"}))
    assert not s1.equals(42)


def test_polygon_affine_helpers_and_device_queries():
    """Small geometry conveniences mirrored from the reference (device/polygon.py:93-300,
    device/device.py:186-240): numpy affine maps instead of shapely."""
    sq = sc.Polygon("sq", layer="l", points=box(2.0, 1.0, points=4))
    assert abs(sq.area - 2.0) < 1e-14
    assert abs(sc.Polygon(points=circle(1.0, 400)).area - np.pi) < 1e-3
    t = sq.translate(dx=1.0, dy=-2.0)
    assert t is not sq and np.allclose(t.points.mean(axis=0) - sq.points.mean(axis=0), [1.0, -2.0])
    r = sq.rotate(90.0)
    assert np.allclose(sorted(np.ptp(r.points, axis=0)), [1.0, 2.0]) and abs(r.area - 2.0) < 1e-14
    assert np.allclose(np.ptp(r.points, axis=0), [1.0, 2.0])          # extents swapped
    c = sq.translate(3.0, 4.0).rotate(30.0, origin="centroid")
    assert np.allclose(c._origin("centroid"), [3.0, 4.0]) and np.allclose(c._origin("center"), [3.0, 4.0])
    s2 = sq.scale(xfact=-2.0, yfact=0.5)
    assert abs(s2.area - 2.0) < 1e-14 and np.allclose(np.ptp(s2.points, axis=0), [4.0, 0.5])
    from superscreen_b200.geometry import signed_area
    assert signed_area(s2.points) > 0                                   # still counter-clockwise after mirroring
    sq.translate(1.0, 0.0, inplace=True)
    assert np.allclose(sq.points[:, 0].min(), 0.0)
    assert sq.set_name("a").name == "a" and sq.set_layer("m").layer == "m"
    with pytest.raises(ValueError):
        sq.rotate(10.0, origin="nope")
    device = sc.Device("d", layers=[sc.Layer("l", Lambda=1.0, z0=0.0)],
                       films=[sc.Polygon("film", layer="l", points=box(4.0, points=4))],
                       holes=[sc.Polygon("hole", layer="l", points=circle(1.0, 20))])
    assert [p.name for p in device.get_polygons()] == ["film", "hole"]
    assert [p.name for p in device.get_polygons("hole")] == ["hole"]
    with pytest.raises(ValueError):
        device.get_polygons("nope")
    assert device.poly_points().shape == (4 + 20, 2) and device.poly_points(holes=False).shape == (4, 2)
    with pytest.raises(ValueError):
        device.mesh_stats_dict()


def test_c_abi_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call: bad sizes return a negative code and leave a
    message in scb_last_error() (error behaviour) -- checkable without a device."""
    lib = _lib.load_library()
    rc = lib.scb_getrf_nopiv(100, None, None, None, None)          # n_pad not a multiple of 128
    assert rc < 0 and b"128" in lib.scb_last_error()
    rc = lib.scb_getrs_nopiv(128, None, None, 0, None, None)       # nrhs must be positive
    assert rc < 0 and b"nrhs" in lib.scb_last_error()
    rc = lib.scb_cdist(4, 0, 1, None, 1, None, None, None)         # dim must be 2 or 3
    assert rc < 0 and b"dim" in lib.scb_last_error()
    rc = lib.scb_biot_savart(1, -1, None, 0, None, None, None, 0.0, 1.0, 1, None, None)
    assert rc < 0 and b"sizes" in lib.scb_last_error()
    with pytest.raises(_lib.SCBError):
        _lib.check(rc)


def test_meshgen_host_geometry_helpers():
    """Host-side pieces of the device mesh generator (no GPU): hull, mitre offset, ring resampling,
    duplicate removal, boundary conformity check; the generator itself refuses to run without a GPU."""
    from superscreen_b200 import _lib, meshgen
    from superscreen_b200.geometry import signed_area

    ell = np.array([[0, 0], [6, 0], [6, 2], [2, 2], [2, 5], [0, 5]], dtype=float)
    hull = meshgen.convex_hull_ring(ell)
    assert signed_area(hull) > 0 and len(hull) == 5 and {tuple(p) for p in hull} == {(0, 0), (6, 0), (6, 2), (2, 5), (0, 5)}
    sq = np.array([[0, 0], [2, 0], [2, 1], [0, 1]], dtype=float)
    off = meshgen.offset_convex_ring(sq, 0.5)
    assert np.allclose(sorted(map(tuple, off)), sorted([(-0.5, -0.5), (2.5, -0.5), (2.5, 1.5), (-0.5, 1.5)]))
    res = meshgen._resample_ring(sq, 0.5)
    assert len(res) == 12 and {tuple(p) for p in sq} <= {tuple(p) for p in res}
    assert np.allclose(np.linalg.norm(np.roll(res, -1, axis=0) - res, axis=1), 0.5)
    dup = np.array([[0, 0], [1, 0], [0, 0], [1, 1], [1, 0]], dtype=float)
    assert np.array_equal(meshgen.ensure_unique(dup), np.array([[0, 0], [1, 0], [1, 1]], dtype=float))
    # two triangles tiling the unit square: conforming to the square, not to a bigger one
    pts = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=float)
    tri = np.array([[0, 1, 2], [0, 2, 3]])
    assert meshgen.boundary_is_conforming(pts, tri, [pts])
    assert not meshgen.boundary_is_conforming(pts, tri, [2 * pts])
    assert abs(meshgen.min_triangle_angle(pts, tri) - 45.0) < 1e-9
    import torch

    if not torch.cuda.is_available():
        with pytest.raises(_lib.SCBError):
            meshgen.generate_mesh(ell, min_points=100)


def test_polygon_fluxoid_functional_and_geometry_caches():
    """`Solution.polygon_fluxoid` evaluates the supercurrent integral through a cached linear functional of
    the mesh currents: it must agree with the direct formula of the reference (solution.py:484-563:
    interpolate J at the polygon points, Lambda J . dl, trapezoid rule; flux = total field over the enclosed
    vertex areas) -- also when J is not finite somewhere (fallback path); `Device.holes_by_film` is memoised
    on the geometry and must follow a changed polygon."""
    from types import SimpleNamespace

    from superscreen_b200.solution import FilmSolution, Solution, _barycentric, _locate
    from superscreen_b200 import units as _u

    sites, elements = disk_mesh(3.0, 900, seed=4)
    # vertex areas: a third of the incident triangle areas (what Mesh.vertex_areas holds)
    p = sites[elements]
    tri_area = 0.5 * np.abs((p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1])
                            - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0]))
    w = np.zeros(len(sites))
    np.add.at(w, elements.ravel(), np.repeat(tri_area / 3.0, 3))
    device = sc.Device("d", layers=[sc.Layer("l", Lambda=0.7, z0=0.0)],
                       films=[sc.Polygon("film", layer="l", points=circle(3.0, 80))],
                       holes=[sc.Polygon("hole", layer="l", points=circle(0.5, 40))])
    device.meshes = {"film": SimpleNamespace(sites=sites, elements=elements, vertex_areas=w)}
    rng = np.random.default_rng(1)
    n = len(sites)
    fs = FilmSolution(stream=rng.standard_normal(n), current_density=rng.standard_normal((n, 2)),
                      applied_field=rng.standard_normal(n), self_field=rng.standard_normal(n),
                      field_from_other_films=rng.standard_normal(n))
    sol = Solution(device=device, film_solutions={"film": fs}, applied_field_func=sc.ConstantField(0.0),
                   field_units="mT", current_units="uA", _device_is_copy=True)
    polygon = circle(1.5, 101)

    def direct(film_solution):
        poly = sc.Polygon(points=polygon)
        pts = poly.points
        ix = np.where(poly.contains_points(sites))[0]
        total = film_solution.applied_field + film_solution.self_field + film_solution.field_from_other_films
        flux = np.sum(total[ix] * w[ix])
        ok, tri, bw = _locate(sites, elements, pts, 16)
        J = np.einsum("qk,qkc->qc", bw, film_solution.current_density[elements[tri]])
        J[~(ok & device.films["film"].contains_points(pts))] = 0
        J[~np.isfinite(J).all(axis=1)] = 0
        int_J = np.trapezoid(0.7 * np.sum(J[:-1] * np.diff(pts, axis=0), axis=1))
        to_phi0 = _u.conversion_factor("mT * um ** 2", "Phi_0")
        si = _u.MU_0 * int_J * _u.conversion_factor("(uA / um) * um ** 2", "A * m") * _u.conversion_factor("Wb", "Phi_0")
        return flux * to_phi0, si

    got = sol.polygon_fluxoid(polygon, film="film", units="Phi_0", with_units=False)
    ref = direct(fs)
    assert abs(got.flux_part - ref[0]) <= 1e-12 * abs(ref[0])
    assert abs(got.supercurrent_part - ref[1]) <= 1e-12 * abs(ref[1])
    again = sol.polygon_fluxoid(polygon, film="film", units="Phi_0", with_units=False)  # cached geometry + conversions
    assert again.flux_part == got.flux_part and again.supercurrent_part == got.supercurrent_part
    # a non-finite current density at one corner takes the reference's zeroing path
    bad = FilmSolution(stream=fs.stream, current_density=fs.current_density.copy(), applied_field=fs.applied_field,
                       self_field=fs.self_field, field_from_other_films=fs.field_from_other_films)
    ok, tri, _ = _locate(sites, elements, sc.Polygon(points=polygon).points, 16)
    bad.current_density[elements[tri[3], 0]] = np.nan
    sol_bad = Solution(device=device, film_solutions={"film": bad}, applied_field_func=sc.ConstantField(0.0),
                       field_units="mT", current_units="uA", _device_is_copy=True)
    got_bad = sol_bad.polygon_fluxoid(polygon, film="film", units="Phi_0", with_units=False)
    ref_bad = direct(bad)
    assert np.isfinite(got_bad.supercurrent_part)
    assert abs(got_bad.supercurrent_part - ref_bad[1]) <= 1e-12 * abs(ref_bad[1])
    # holes_by_film: memoised on the geometry, follows a moved hole
    assert [h.name for h in device.holes_by_film()["film"]] == ["hole"]
    assert [h.name for h in device.holes_by_film()["film"]] == ["hole"]
    device.holes["hole"] = sc.Polygon("hole", layer="l", points=circle(0.5, 40) + np.array([10.0, 0.0]))  # outside
    assert device.holes_by_film()["film"] == []


def test_default_fluxoid_polygons_of_a_ring():
    """`make_fluxoid_polygons` (the default of `Device.mutual_inductance_matrix` / `find_fluxoid_solution`, reference
    fluxoid.py:12-52): a hole is grown by half the distance to the nearest other polygon of its layer -- for the
    ring of the reference's tests (hole r = 2 in a film r = 4) the circle of radius 3, counter-clockwise, and it
    contains the hole and lies in the film."""
    from superscreen_b200.fluxoid import make_fluxoid_polygons

    device = sc.Device("ring", layers=[sc.Layer("base", Lambda=1.0, z0=0.0)],
                       films=[sc.Polygon("ring", layer="base", points=circle(4.0, 200))],
                       holes=[sc.Polygon("hole", layer="base", points=circle(2.0, 200))])
    polys = make_fluxoid_polygons(device)
    assert list(polys) == ["hole"]
    p = polys["hole"]
    r = np.linalg.norm(p, axis=1)
    assert np.allclose(r, 3.0, atol=2e-3)
    area2 = np.sum(p[:-1, 0] * p[1:, 1] - p[1:, 0] * p[:-1, 1]) if np.array_equal(p[0], p[-1]) else \
        np.sum(p[:, 0] * np.roll(p[:, 1], -1) - np.roll(p[:, 0], -1) * p[:, 1])
    assert area2 > 0  # counter-clockwise
    assert points_in_polygon(p, device.holes["hole"].points).all()
    assert device.films["ring"].contains_points(p).all()
    assert len(make_fluxoid_polygons(device, holes="hole", interp_points=64)["hole"]) in (64, 65)


def test_polygon_flux():
    """`Solution.polygon_flux` (reference solution.py:430-482): total field times vertex areas over the mesh vertices
    inside a named polygon of the device, in flux units; unknown names raise ValueError."""
    from types import SimpleNamespace

    from superscreen_b200.solution import FilmSolution, Solution
    from superscreen_b200 import units as _u

    sites, elements = disk_mesh(3.0, 700, seed=6)
    p = sites[elements]
    tri_area = 0.5 * np.abs((p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1])
                            - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0]))
    w = np.zeros(len(sites))
    np.add.at(w, elements.ravel(), np.repeat(tri_area / 3.0, 3))
    device = sc.Device("d", layers=[sc.Layer("l", Lambda=0.7, z0=0.0)],
                       films=[sc.Polygon("film", layer="l", points=circle(3.0, 80))],
                       holes=[sc.Polygon("hole", layer="l", points=circle(0.8, 40))])
    device.meshes = {"film": SimpleNamespace(sites=sites, elements=elements, vertex_areas=w)}
    rng = np.random.default_rng(3)
    n = len(sites)
    fs = FilmSolution(stream=np.zeros(n), current_density=np.zeros((n, 2)), applied_field=rng.standard_normal(n),
                      self_field=rng.standard_normal(n))
    sol = Solution(device=device, film_solutions={"film": fs}, applied_field_func=sc.ConstantField(0.0),
                   field_units="uT", current_units="uA", _device_is_copy=True)
    for name in ("film", "hole"):
        ix = points_in_polygon(device.get_polygons(include_terminals=False)[[q.name for q in device.get_polygons(
            include_terminals=False)].index(name)].points, sites)
        ref_uT_um2 = np.sum(fs.total_field[ix] * w[ix])
        got = sol.polygon_flux(name, with_units=False)                      # default units: field * length^2
        assert abs(got - ref_uT_um2) <= 1e-12 * abs(ref_uT_um2)
        got_phi0 = sol.polygon_flux(name, units="Phi_0", with_units=False)
        assert abs(got_phi0 - ref_uT_um2 * _u.conversion_factor("uT * um ** 2", "Phi_0")) <= 1e-12 * abs(got_phi0)
        assert sol.polygon_flux(name, units="Phi_0").magnitude == got_phi0
    with pytest.raises(ValueError):
        sol.polygon_flux("nope")


def test_polygon_resample_and_hdf5_methods():
    """`Polygon.resample` (uniform in arc length along the closed boundary, reference device/polygon.py:483-506)
    and the `to_hdf5` / `from_hdf5` methods of Layer / Polygon (reference group layouts) on an in-memory group."""
    from superscreen_b200.io import MemoryGroup

    sq = sc.Polygon("sq", layer="l", points=box(2.0, points=4))
    rs = sq.resample(41)
    assert rs.name == "sq" and rs.layer == "l"
    p = rs.points
    closed = p if np.allclose(p[0], p[-1]) else np.concatenate([p, p[:1]])
    seg = np.linalg.norm(np.diff(closed, axis=0), axis=1)
    assert np.allclose(np.abs(closed).max(axis=1), 1.0)          # all points on the boundary of the 2 x 2 box
    assert np.allclose(seg[seg > 1e-12], 8.0 / 40, atol=1e-12)   # perimeter 8 in 40 equal steps
    assert np.array_equal(sq.resample(0).points, sq.points) and sq.resample(0) is not sq
    assert len(sq.resample().points) == len(sq.points)
    g = MemoryGroup()
    sq.to_hdf5(g.create_group("poly"))
    back = sc.Polygon.from_hdf5(g["poly"])
    assert back.name == "sq" and back.layer == "l" and np.array_equal(back.points, sq.points)
    layer = sc.Layer("base", london_lambda=0.5, thickness=0.05, z0=0.25)
    layer.to_hdf5(g.create_group("layer"))
    lb = sc.Layer.from_hdf5(g["layer"])
    assert lb.name == "base" and lb.z0 == 0.25 and lb.thickness == 0.05 and lb.london_lambda == 0.5
    assert np.isclose(lb.Lambda, 0.5**2 / 0.05)


def test_device_rigid_transformations():
    """Device.scale / rotate / mirror_layers / translate / translation (reference device/device.py:256-381) on a
    device without meshes: new devices with transformed polygons and layers, the original untouched."""
    device = sc.Device("d", layers=[sc.Layer("a", Lambda=1.0, z0=0.5), sc.Layer("b", Lambda=2.0, z0=-1.0)],
                       films=[sc.Polygon("film", layer="a", points=box(2.0, points=4)),
                              sc.Polygon("plate", layer="b", points=box(4.0, points=4))],
                       holes=[sc.Polygon("hole", layer="a", points=circle(0.3, 24))])
    orig = {p.name: p.points.copy() for p in device.get_polygons()}
    moved = device.translate(1.5, -2.0, dz=0.25)
    assert moved is not device
    for p in moved.get_polygons():
        assert np.allclose(p.points, orig[p.name] + np.array([1.5, -2.0]))
    assert moved.layers["a"].z0 == 0.75 and moved.layers["b"].z0 == -0.75 and device.layers["a"].z0 == 0.5
    for p in device.get_polygons():
        assert np.array_equal(p.points, orig[p.name])
    rot = device.rotate(90.0)
    assert np.allclose(rot.films["film"].points, orig["film"] @ np.array([[0.0, 1.0], [-1.0, 0.0]]))
    sca = device.scale(xfact=-2.0, yfact=0.5, origin=(1.0, 0.0))
    want = (orig["hole"] - [1.0, 0.0]) * [-2.0, 0.5] + [1.0, 0.0]  # (a mirrored polygon is re-oriented counter-clockwise)
    srt = lambda a: a[np.lexsort((np.round(a[:, 1], 9), np.round(a[:, 0], 9)))]
    assert np.allclose(srt(np.unique(np.round(sca.holes["hole"].points, 12), axis=0)),
                       srt(np.unique(np.round(want, 12), axis=0)))
    with pytest.raises(TypeError):
        device.rotate(10.0, origin=[0, 0])
    mir = device.mirror_layers(about_z=1.0)
    assert mir.layers["a"].z0 == 0.5 and mir.layers["b"].z0 == 2.0
    with device.translation(3.0, 0.0, dz=1.0):
        assert np.allclose(device.films["film"].points, orig["film"] + np.array([3.0, 0.0]))
        assert device.layers["b"].z0 == 0.0
    assert np.allclose(device.films["film"].points, orig["film"]) and device.layers["b"].z0 == -1.0
    assert device.translate(1.0, 1.0, inplace=True) is device


def test_polygon_on_boundary_and_named_set_operations():
    """Polygon.on_boundary (reference device/polygon.py:164-190) and the `name` argument of union / difference."""
    sq = sc.Polygon("sq", layer="l", points=box(2.0, points=4))
    pts = np.array([[1.0, 0.3], [0.9995, -0.2], [0.99, 0.0], [0.0, 0.0], [1.0005, 1.0005], [1.01, 0.0], [-1.0, -1.0]])
    got = sq.on_boundary(pts, radius=1e-3)
    assert got.tolist() == [True, True, False, False, True, False, True]
    assert sq.on_boundary(pts, radius=1e-3, index=True).tolist() == [0, 1, 4, 6]
    assert sq.on_boundary(pts, radius=0.02).tolist() == [True, True, True, False, True, True, True]
    hole = sc.Polygon("h", layer="l", points=circle(0.4, 32))
    ring = sq.difference(hole, name="ring")
    assert ring.name == "ring" and ring.layer == "l"
    assert ring.contains_points([[0.8, 0.0], [0.0, 0.0]]).tolist() == [True, False]
    both = sq.union(sc.Polygon("far", layer="l", points=box(1.0, points=4) + np.array([5.0, 0.0])), name="both")
    assert both.name == "both" and both.contains_points([[5.0, 0.0], [3.0, 0.0]]).tolist() == [True, False]
    assert sq.union(hole).name == "sq"
    with pytest.raises(NotImplementedError):
        sq.difference(hole, symmetric=True)


def test_applied_field_sources_known_answers():
    """Known answers of the applied-field sources (these run without the reference): the on-axis field of a z dipole
    mu_0 m / (2 pi d^3), the flux n Phi_0 / 2 of a monopole through the half space above it, and the total flux
    n Phi_0 carried by a Pearl vortex."""
    from scipy.constants import mu_0

    import superscreen_b200.sources as S

    d = 2.5e-6
    B = S.dipole_field((0.0, 0.0, d), r0=(0, 0, 0), moment=(0, 0, 3e-15))
    assert np.allclose(B, [0, 0, mu_0 * 3e-15 / (2 * np.pi * d**3)], rtol=1e-14, atol=0)
    Bz = S.DipoleField(dipole_positions=(0, 0, 0), dipole_moments=(0, 0, 1e8), component="z", length_units="um",
                       moment_units="mu_B")(np.array([0.0]), np.array([0.0]), np.array([2.5]))
    assert np.allclose(Bz, mu_0 * 1e8 * 9.2740100783e-24 / (2 * np.pi * d**3), rtol=1e-12)
    # monopole: integral of mu_0 H_z over the plane z = h above it is n Phi_0 / 2 (half of the solid angle... of 2 n Phi_0)
    g = np.linspace(-400, 400, 1601)
    X, Y = np.meshgrid(g, g)
    hz = S.MonopoleField(r0=(0, 0, 0), nPhi0=3)(X.ravel(), Y.ravel(), np.full(X.size, 1.0))
    assert abs(hz.sum() * (g[1] - g[0]) ** 2 - 3.0) < 0.02  # n Phi_0 (2 pi / 2 pi), minus what leaves the window
    xs = np.linspace(-40, 40, 256)  # (an even count puts a sample at k = 0: the grid sum is the zero-frequency term)
    pv = S.PearlVortexField(r0=(0, 0, 0), Lambda=0.5, nPhi0=2, xs=xs, ys=xs)
    Xs, Ys = np.meshgrid(xs, xs)
    total = pv(Xs.ravel(), Ys.ravel(), np.full(Xs.size, 0.3)).sum() * (xs[1] - xs[0]) ** 2
    assert abs(total - 2.0) < 1e-9


def test_parameter_arithmetic():
    """Parameters compose with + - * / ** among themselves and with real numbers (reference parameter.py:146-273)."""
    import superscreen_b200.sources as S

    x, y, z = np.linspace(-1, 1, 7), np.linspace(2, 3, 7), np.full(7, 0.5)
    c, m = S.ConstantField(0.25), S.MonopoleField(r0=(0, 0, -1.0), nPhi0=2)
    f = 2 * c + m / 4 - 1
    assert isinstance(f, S.CompositeParameter) and isinstance(f, S.Parameter)
    assert np.allclose(f(x, y, z), 2 * c(x, y, z) + m(x, y, z) / 4 - 1)
    assert np.allclose((c ** 2)(x, y, z), 0.0625) and np.allclose((2 ** c)(x, y, z), 2 ** 0.25)
    assert np.allclose((1 - c)(x, y, z), 0.75) and np.allclose((1 / c)(x, y, z), 4.0) and np.allclose((c * m)(x, y, z),
                                                                                                  0.25 * m(x, y, z))
    lam = S.Parameter(lambda x, y, a=1.0: a * (1 + x**2), a=0.5)  # a 2-D parameter (e.g. Lambda(x, y))
    assert np.allclose((lam + 1)(x, y), 1 + 0.5 * (1 + x**2))
    with pytest.raises(TypeError):
        S.CompositeParameter(1, 2, "+")
    with pytest.raises(TypeError):
        c + "a"
    with pytest.raises(ValueError):
        S.CompositeParameter(c, 2, "%")


def test_host_section_timing(monkeypatch):
    """`SCB_HOST_TIMING`: the NVTX ranges of the path also accumulate host wall time per section name (film names in
    brackets are folded together); off by default and free of cost."""
    import time

    from superscreen_b200 import _lib

    _lib.host_times.clear()
    with _lib.nvtx_range("scb.section[a]"):
        pass
    assert _lib.host_times == {}
    monkeypatch.setattr(_lib, "_HOST_TIMING", True)
    for name in ("scb.section[a]", "scb.section[b]", "scb.other"):
        with _lib.nvtx_range(name):
            time.sleep(0.002)
    assert set(_lib.host_times) == {"scb.section", "scb.other"}
    assert _lib.host_times["scb.section"][0] == 2 and _lib.host_times["scb.section"][1] >= 0.003
    report = _lib.host_timing_report()
    assert "scb.section" in report and "2 x" in report and _lib.host_times == {}
