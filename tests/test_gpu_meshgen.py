"""GPU tests of the device-side mesh generation: scb_delaunay against
scipy.spatial.Delaunay (identical triangle sets), generate_mesh contract (reference device/utils.py:17-136:
input points are vertices, boundary conforming, min_points / max_edge_length honoured), Device.make_mesh
-> solve against the oracle."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sc():
    import torch

    assert torch.cuda.is_available()
    import superscreen_b200 as sc

    return sc


def _canon(tri):
    t = np.sort(np.asarray(tri, dtype=np.int64), axis=1)
    return set(map(tuple, t))


def _scipy_region_triangles(points, rings):
    """scipy's Delaunay triangulation of the same points, cut down to the region like generate_mesh does."""
    from scipy.spatial import Delaunay

    from superscreen_b200.geometry import points_in_polygon

    tri = Delaunay(points).simplices.astype(np.int64)
    p = points[tri]
    cross = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    cen = p.mean(axis=1)
    inside = np.zeros(len(tri), dtype=bool)
    for r in rings:
        inside ^= points_in_polygon(r, cen)
    # (Qhull emits zero-area triangles over collinear boundary points -- circumradius ~1e13 --: not part of a mesh)
    longest2 = np.max(np.sum((p - np.roll(p, -1, axis=1)) ** 2, axis=2), axis=1)
    return tri[inside & (np.abs(cross) > 1e-9 * longest2)]


@pytest.mark.parametrize("n,seed", [(50, 0), (300, 1), (5000, 2), (60000, 3)])
def test_delaunay_matches_scipy(sc, n, seed):
    import torch
    from scipy.spatial import Delaunay

    from superscreen_b200 import meshgen

    rng = np.random.default_rng(seed)
    pts = rng.random((n, 2)) * np.array([3.0, 1.0]) + np.array([-1.0, 5.0])
    tri = meshgen.delaunay(pts).cpu().numpy()
    ref = Delaunay(pts).simplices
    # identical triangle sets, apart from hull slivers flatter than any mesh could use (either side may keep them)
    a, b = _canon(tri), _canon(ref)
    diff = a ^ b
    for t in diff:
        p = pts[list(t)]
        area2 = abs((p[1, 0] - p[0, 0]) * (p[2, 1] - p[0, 1]) - (p[1, 1] - p[0, 1]) * (p[2, 0] - p[0, 0]))
        sides = [np.linalg.norm(p[i] - p[j]) for i, j in ((0, 1), (1, 2), (2, 0))]
        circumradius = np.prod(sides) / (2.0 * area2)
        # (scb_delaunay does not produce triangles whose circumradius exceeds 1e4 domain extents)
        assert circumradius > 1e3 * 3.0, f"triangle {t} differs and is not a degenerate hull sliver"
    assert len(diff) <= 4
    # layout contract: counter-clockwise, smallest vertex first, ordered by it; deterministic
    p = pts[tri]
    cross = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    assert (cross > 0).all()
    assert (tri[:, 0] < tri[:, 1]).all() and (tri[:, 0] < tri[:, 2]).all()
    assert (np.diff(tri[:, 0]) >= 0).all()
    again = meshgen.delaunay(torch.as_tensor(pts).cuda()).cpu().numpy()
    assert np.array_equal(tri, again)


def _check_mesh(sc, points, triangles, rings, fixed, min_points, max_edge_length, holes):
    from superscreen_b200 import meshgen
    from superscreen_b200.geometry import signed_area

    assert len(points) >= min_points
    # every input polygon point is a vertex
    have = set(map(tuple, points))
    assert all(tuple(p) in have for p in fixed)
    # orientation, edge lengths
    p = points[triangles]
    cross = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    assert (cross > 0).all()
    lengths = np.concatenate([np.linalg.norm(p[:, i] - p[:, (i + 1) % 3], axis=1) for i in range(3)])
    assert lengths.max() <= max_edge_length
    # the triangles tile the region exactly and the mesh boundary is the polygon boundary
    area = abs(signed_area(rings[0])) - sum(abs(signed_area(r)) for r in rings[1:])
    assert abs(0.5 * cross.sum() - area) <= 1e-9 * area
    assert meshgen.boundary_is_conforming(points, triangles, rings)
    # a valid manifold for the FEM operators (raises otherwise); Euler characteristic 1 - #holes
    mesh = sc.Mesh.from_triangulation(points, triangles)
    d = mesh._data
    assert len(points) - d.n_edges + len(triangles) == 1 - holes
    # it IS the Delaunay triangulation of its vertices (restricted to the region)
    ref = _scipy_region_triangles(points, rings)
    a, b = _canon(triangles), _canon(ref)
    assert a == b, f"{len(a ^ b)} of {len(a)} triangles differ from scipy's"
    assert meshgen.min_triangle_angle(points, triangles) > 10.0
    return mesh


def test_generate_mesh_annulus_with_hole_cut_out(sc):
    from superscreen_b200 import meshgen
    from superscreen_b200.geometry import circle

    outer, hole = circle(4.0, 100), circle(2.0, 60)
    points, triangles = meshgen.generate_mesh(outer, hole_coords=[hole], min_points=3000, max_edge_length=0.3)
    _check_mesh(sc, points, triangles, [outer, hole], np.concatenate([outer, hole]), 3000, 0.3, holes=1)
    again = meshgen.generate_mesh(outer, hole_coords=[hole], min_points=3000, max_edge_length=0.3)
    assert np.array_equal(points, again[0]) and np.array_equal(triangles, again[1])
    other_seed = meshgen.generate_mesh(outer, hole_coords=[hole], min_points=3000, max_edge_length=0.3, seed=5)
    assert not np.array_equal(points[-100:], other_seed[0][-100:])


def test_generate_mesh_nonconvex_polygon_and_refinement(sc):
    from superscreen_b200 import meshgen

    ell = np.array([[0, 0], [6, 0], [6, 2], [2, 2], [2, 5], [0, 5]], dtype=float)
    coarse = meshgen.generate_mesh(ell)
    _check_mesh(sc, *coarse, [ell], ell, 0, np.inf, holes=0)
    fine = meshgen.generate_mesh(ell, min_points=8000)
    _check_mesh(sc, *fine, [ell], ell, 8000, np.inf, holes=0)
    assert len(fine[0]) < 2 * 8000  # the refinement loop does not overshoot wildly
    edge = meshgen.generate_mesh(ell, max_edge_length=0.12)
    _check_mesh(sc, *edge, [ell], ell, 0, 0.12, holes=0)
    hull = meshgen.generate_mesh(ell, convex_hull=True, min_points=500)
    hull_ring = meshgen.convex_hull_ring(ell)
    _check_mesh(sc, *hull, [hull_ring], ell, 500, np.inf, holes=0)
    with pytest.raises(ValueError):
        meshgen.generate_mesh(ell, convex_hull=True, boundary=ell)


def test_device_make_mesh_then_solve_against_oracle(sc):
    from oracle import port
    from superscreen_b200.geometry import circle

    film_poly, hole_poly = circle(4.0, 64), circle(2.0, 40)
    device = sc.Device(
        "ring", layers=[sc.Layer("base", Lambda=5.0, z0=0.5)],
        films=[sc.Polygon("ring", layer="base", points=film_poly)],
        holes=[sc.Polygon("hole", layer="base", points=hole_poly)],
    )
    device.make_mesh(min_points=2500, buffer_factor=0.05)
    mesh = device.meshes["ring"]
    sites, elements = mesh.sites, mesh.elements
    assert len(sites) >= 2500
    have = set(map(tuple, sites))
    assert all(tuple(p) in have for p in device.films["ring"].points)
    assert all(tuple(p) in have for p in device.holes["hole"].points)
    # the mesh covers the buffered region: 5 % of the 8 um extent on every side
    assert sites[:, 0].max() >= 4.0 + 0.39 and sites[:, 0].min() <= -4.0 - 0.39
    sol = sc.solve(device, applied_field=sc.ConstantField(1.0), circulating_currents={"hole": "1 mA"},
                   field_units="mT", current_units="uA")[-1]
    fs = sol.film_solutions["ring"]
    info = sc.solver.utils.make_film_info(device=device, vortices=[], circulating_currents={"hole": 1000.0},
                                          terminal_currents={})["ring"]
    film = port.OracleFilm(name="ring", mesh=port.build_mesh(sites, elements), z0=0.5, Lambda=np.full(len(sites), 5.0),
                           interior_indices=info.interior_indices, hole_indices=info.hole_indices)
    port.factorize_film(film)
    conv = sc.field_conversion_factor("mT", "uA", "um").magnitude
    ref = port.solve_film(film, np.full(len(sites), 1.0) * conv, {"hole": 1000.0}, conv)
    assert rel_l2(fs.stream, ref.stream) <= 1e-8
    assert rel_l2(fs.current_density, ref.current_density) <= 1e-8
    # physics: the hole carries the circulating current (reference tests/test_solve.py:161-183: +-5 %)
    smoothed = sc.Polygon("p", points=circle(3.0, 80)).make_mesh(min_points=400, smooth=3)
    assert len(smoothed.sites) >= 400
