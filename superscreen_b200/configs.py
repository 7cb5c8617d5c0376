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
