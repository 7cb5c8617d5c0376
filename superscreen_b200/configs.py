"""The five synthetic benchmark / parity configurations of BASELINE.json.

Each builder returns a ``Device`` with seeded synthetic meshes attached plus whatever the
configuration needs (fluxoid polygons ...).  Sizes can be scaled down for the parity tests.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from .device import Device, Layer, Polygon
from .geometry import box, circle
from .synthetic import disk_mesh, make_mesh, square_mesh


def c1_ring(n_vertices: int = 2000, seed: int = 0):
    """C1: single-layer ring (film r=4, hole r=2, london_lambda=0.5, thickness=0.05 -> Lambda=5,
    z0=0.5; geometry of reference test_solve.py:12-36), mesh disk r=4.4."""
    film_poly, hole_poly = circle(4.0, 100), circle(2.0, 100)
    device = Device("ring", layers=[Layer("base", london_lambda=0.5, thickness=0.05, z0=0.5)],
                    films=[Polygon("ring", layer="base", points=film_poly)],
                    holes=[Polygon("ring_hole", layer="base", points=hole_poly)])
    device.set_meshes({"ring": disk_mesh(4.4, n_vertices, embedded=[film_poly, hole_poly], seed=seed)})
    return device, {"ring_hole": circle(3.0, 201)}


def c2_square(n_vertices: int = 20164, seed: int = 0, Lambda: float = 0.1, side: float = 10.0):
    """C2 / C5: box(side) film, buffer 0 (mesh boundary == film boundary)."""
    device = Device("square", layers=[Layer("layer", Lambda=Lambda, z0=0.0)],
                    films=[Polygon("film", layer="layer", points=box(side, points=4))])
    device.set_meshes({"film": square_mesh(side, n_vertices, seed=seed)})
    return device


def c3_susceptometer(n_vertices: int = 4000, seed: int = 0):
    """C3: scanning-SQUID-like stack: 4 films on 2 layers (z=0: field-coil ring r 2-3 + its shield
    plate; z=0.5: pickup ring r 0.5-1 + shield), Lambda = 0.08^2/0.2, circulating current in the
    field coil; M = Phi_pickup / I_fieldcoil."""
    Lam = 0.08**2 / 0.2
    layers = [Layer("W1", Lambda=Lam, z0=0.0), Layer("BE", Lambda=Lam, z0=0.5)]
    fc, fc_hole = circle(3.0, 80), circle(2.0, 60)
    fc_shield = box(5.0, 2.0, points=40, center=(0.0, -4.5))
    pl, pl_hole = circle(1.0, 60, center=(0.0, 0.0)), circle(0.5, 40, center=(0.0, 0.0))
    pl_shield = box(3.0, 2.0, points=40, center=(0.0, -3.0))
    films = [Polygon("fc", layer="W1", points=fc), Polygon("fc_shield", layer="W1", points=fc_shield),
             Polygon("pl", layer="BE", points=pl), Polygon("pl_shield", layer="BE", points=pl_shield)]
    holes = [Polygon("fc_center", layer="W1", points=fc_hole), Polygon("pl_center", layer="BE", points=pl_hole)]
    device = Device("susceptometer", layers=layers, films=films, holes=holes)
    meshes = {
        "fc": disk_mesh(3.3, n_vertices, embedded=[fc, fc_hole], seed=seed),
        "fc_shield": make_mesh(box(5.5, 2.2, points=4, center=(0.0, -4.5)), target_vertices=n_vertices,
                               embedded=[fc_shield], seed=seed + 1),
        "pl": disk_mesh(1.1, n_vertices, embedded=[pl, pl_hole], seed=seed + 2),
        "pl_shield": make_mesh(box(3.3, 2.2, points=4, center=(0.0, -3.0)), target_vertices=n_vertices,
                               embedded=[pl_shield], seed=seed + 3),
    }
    device.set_meshes(meshes)
    polygons = {"fc_center": circle(2.5, 201), "pl_center": circle(0.75, 201)}
    return device, polygons


def c4_ring_array(n_rings: int = 8, n_vertices: int = 5000, seed: int = 0, pitch: float = 12.0):
    """C4: copies of the C1 ring on a 2 x 4 grid (pitch 12 um), one layer; 8 x 8 mutual-inductance
    matrix, one film factorization per GPU."""
    films, holes, meshes, polygons = [], [], {}, {}
    for k in range(n_rings):
        c = (pitch * (k % 4), pitch * (k // 4))
        fp, hp = circle(4.0, 100, center=c), circle(2.0, 100, center=c)
        films.append(Polygon(f"ring{k}", layer="base", points=fp))
        holes.append(Polygon(f"hole{k}", layer="base", points=hp))
        meshes[f"ring{k}"] = disk_mesh(4.4, n_vertices, embedded=[fp, hp], seed=seed + k, center=c)
        polygons[f"hole{k}"] = circle(3.0, 201, center=c)
    device = Device("ring_array", layers=[Layer("base", london_lambda=0.5, thickness=0.05, z0=0.5)],
                    films=films, holes=holes)
    device.set_meshes(meshes)
    return device, polygons


def c5_large(n_vertices: int = 60000, seed: int = 0, Lambda: float = 0.1):
    """C5: large single film (box(10)), 64 uniform fields linspace(0.1, 6.4, 64) mT against one LU,
    field_at_position on a grid over [-7.5, 7.5]^2 at z = 1 um."""
    device = c2_square(n_vertices, seed=seed, Lambda=Lambda)
    fields = np.linspace(0.1, 6.4, 64)
    return device, fields


C5_LAMBDA_SWEEP = (0.05, 0.1, 0.2, 0.4)


def with_lambda(device, Lambda: float):
    """A copy of a single-layer device with another Lambda that shares the (device-resident) meshes:
    only the system assembly and the LU are redone by ``factorize_model`` (C5 Lambda sweep)."""
    layers = [Layer(layer.name, Lambda=Lambda, z0=layer.z0) for layer in device.layers.values()]
    new = Device(device.name, layers=layers, films=list(device.films.values()), holes=list(device.holes.values()))
    new.meshes = dict(device.meshes)
    return new


def evaluation_grid(n: int = 1000, half_width: float = 7.5, z: float = 1.0) -> np.ndarray:
    xs = np.linspace(-half_width, half_width, n)
    X, Y = np.meshgrid(xs, xs)
    return np.column_stack([X.ravel(), Y.ravel(), np.full(X.size, z)])


# ----------------------------------------------------------------------------------------------
# IBM scanning-SQUID susceptometers (the only physically validated numbers the reference publishes:
# field-coil <-> pickup-loop mutual inductance 69 +- 7 Phi_0/A ("small", 100 nm pickup loop) and
# 166 +- 4 Phi_0/A ("medium", 300 nm), docs/notebooks/scanning-squid.ipynb cell 3; Rev. Sci. Instrum.
# 87, 093702 (2016) Table 1).  Coordinates restated from docs/notebooks/squids/ibm/{small,medium}.py,
# layer stack from docs/notebooks/squids/ibm/layers.py (align="middle"), `with_terminals=False` route of
# docs/notebooks/squids/mutuals.py:52-55 (closed field coil, circulating current in `fc_center`).
# ----------------------------------------------------------------------------------------------
def ibm_squid_layers(london_lambda: float = 0.08, z0: float = 0.0, d_BE: float = 0.16, d_I1: float = 0.15,
                     d_W1: float = 0.10, d_I2: float = 0.13, d_W2: float = 0.20):
    z_W2 = z0 + d_W2 / 2
    z_W1 = z_W2 + d_I2 + d_W1 / 2
    z_BE = z_W1 + d_I1 + d_BE / 2
    return [Layer("W2", london_lambda=london_lambda, thickness=d_W2, z0=z_W2),
            Layer("W1", london_lambda=london_lambda, thickness=d_W1, z0=z_W1),
            Layer("BE", london_lambda=london_lambda, thickness=d_BE, z0=z_BE)]


def _keyhole(radius: float, half_width: float, y_bottom: float, points: int = 160) -> np.ndarray:
    """Counter-clockwise ring: a circle of ``radius`` about the origin whose bottom opens into a slot
    ``|x| <= half_width`` reaching down to ``y_bottom``."""
    a0 = np.arcsin(half_width / radius)
    t = np.linspace(-np.pi / 2 + a0, 3 * np.pi / 2 - a0, points)
    arc = radius * np.stack([np.cos(t), np.sin(t)], axis=1)   # from the right slot edge around to the left one
    ys = np.linspace(arc[-1, 1], y_bottom, 40)[1:]
    left = np.stack([np.full(len(ys), -half_width), ys], axis=1)
    bottom = np.stack([np.linspace(-half_width, half_width, 12)[1:-1], np.full(10, y_bottom)], axis=1)
    right = np.stack([np.full(len(ys), half_width), ys[::-1]], axis=1)
    return np.concatenate([arc, left, bottom, right], axis=0)


def ibm_susceptometer(size: str = "small", n_vertices: int = 4000, seed: int = 0, attach_meshes: bool = True):
    """IBM susceptometer with a closed field coil -> (device, {"pl_center": fluxoid ring}).

    Every film gets its own mesh: a jittered-hex Delaunay mesh of the film's bounding box (8 % margin)
    with the rings of the film and of its holes embedded as vertices (`synthetic.make_mesh`)."""
    if size == "small":
        pl_length, ri_pl, ro_pl, ri_fc, ro_fc = 2.5, 0.1, 0.3, 0.5, 1.0125
        pl_center = Polygon("pl_center", layer="W1", points=box(0.20, pl_length, center=(0, -pl_length / 2 + ri_pl)))
        pl = Polygon("pl", layer="W1", points=box(2 * ro_pl, pl_length + ro_pl,
                                                  center=(0, -(pl_length + 0.3) / 2 + 3 * ri_pl))).union(
            np.array([[-0.30, -1.10], [-0.385, -1.7], [-0.64, -2.57], [+0.62, -2.57], [+0.35, -1.67], [+0.30, -1.15]]))
        pl_shield1 = Polygon("pl_shield1", layer="W2", points=np.array(
            [[+0.35, -ri_pl], [-0.35, -ri_pl], [-0.98, -2.65], [-1.05, -2.80], [+1.05, -2.80], [+0.98, -2.65]]))
        pl_shield2 = Polygon("pl_shield2", layer="BE", points=np.array(
            [[+0.5, -1.5 - ri_pl], [-0.5, -1.5 - ri_pl], [-0.84, -2.70], [+0.84, -2.70]]))
        fc = Polygon("fc", layer="BE", points=circle(ro_fc, center=(0, 0.01))).union(np.array(
            [[2.30, -0.35], [2.00, -0.04], [1.19, 0.54], [0.60, 0.80], [0.40, -0.9], [1.1, -1.30], [1.35, -1.9]]))
        fc_shield = Polygon("fc_shield", layer="W1", points=np.array(
            [[2.5, -0.45], [2.15, -0.15], [2.00, -0.04], [1.31, 0.43], [0.81, -0.08], [0.66, -1.23], [1.25, -2.65]]))
        fc_center = Polygon("fc_center", layer="BE", points=circle(ri_fc)).union(np.array(
            [[1.7, -0.47], [0.95, 0.02], [0.6, 0.11], [0.4, 0.28], [0.33, -0.34], [0.69, -0.44], [1.4, -0.9]]))
        # fluxoid contour half way between the slot-shaped hole and the outline of the pickup loop
        ring = box(0.40, 2.65, points=240, center=(0.0, -1.125))
        name = "ibm_100nm"
    elif size == "medium":
        pl_length, ri_pl, ro_pl, ri_fc, ro_fc = 2.2, 0.3, 0.5, 1.0, 1.5
        pl_center = Polygon("pl_center", layer="W1", points=circle(ri_pl)).union(
            box(0.2, pl_length, center=(0, -pl_length / 2 - 0.9 * ri_pl)))
        pl = Polygon("pl", layer="W1", points=circle(ro_pl)).union(
            np.array([[+0.3, -0.4], [-0.3, -0.4], [-0.87, -2.8], [+0.85, -2.8]]))
        pl_shield2 = Polygon("pl_shield2", layer="BE", points=np.array(
            [[+0.75, -(2.3 - ri_pl)], [-0.75, -(2.3 - ri_pl)], [-0.99, -3.0], [+0.96, -3.0]]))
        pl_shield1 = Polygon("pl_shield1", layer="W2", points=np.array(
            [[+0.3, -0.4], [-0.3, -0.4], [-1.0, -2.7], [-1.2, -3.2], [+1.2, -3.2], [+1.0, -2.7]]))
        fc_center = Polygon("fc_center", layer="BE", points=circle(ri_fc)).union(np.array(
            [[2.2, -1.2], [1.7, -0.45], [0.97, 0.0], [0.8, -0.5], [1.23, -0.78], [1.4, -0.9], [1.85, -1.55]]))
        fc = Polygon("fc", layer="BE", points=circle(ro_fc)).union(np.array(
            [[3.0, -1.05], [2.0, 0.0], [1.68, 0.2], [1.2, 0.52], [0.85, -1.18], [1.12, -1.35], [1.55, -2.35]]))
        fc_shield = Polygon("fc_shield", layer="W1", points=np.array(
            [[3.25, -1.25], [2.96, -0.9], [2.0, 0.0], [1.67, 0.19], [1.11, -0.37], [0.9, -1.4], [1.5, -2.9]]))
        ring = _keyhole(0.40, 0.17, -(pl_length + 0.9 * ri_pl) - 0.08)
        name = "ibm_300nm"
    else:
        raise ValueError(f"Unknown susceptometer size {size!r} (small | medium).")
    films = [fc, fc_shield, pl_shield1, pl_shield2, pl]
    holes = [pl_center, fc_center]
    device = Device(name, layers=ibm_squid_layers(), films=films, holes=holes, length_units="um")
    holes_by_film = device.holes_by_film()
    meshes = {}
    for k, film in enumerate(films):
        pts = film.points
        lo, hi = pts.min(axis=0), pts.max(axis=0)
        c, ext = 0.5 * (lo + hi), (hi - lo) * 1.08
        rings = [r[:-1] for r in film.rings]
        for hole in holes_by_film[film.name]:
            rings += [r[:-1] for r in hole.rings]
        meshes[film.name] = make_mesh(box(ext[0], ext[1], points=4, center=tuple(c)), target_vertices=n_vertices,
                                      embedded=rings, seed=seed + k)
    if not attach_meshes:  # (host-only use: the triangulations without the device-built operators)
        return device, {"pl_center": ring}, meshes
    device.set_meshes(meshes)
    return device, {"pl_center": ring}
