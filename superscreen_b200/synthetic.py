"""Seeded synthetic meshes for the benchmark / parity configurations.

Mesh generation is an *input producer* and is out of scope for the hot path (the reference
uses meshpy/Triangle, which is not installable here).  These generators build a jittered
hexagonal point lattice inside a convex outline, add the exact polygon-boundary points of
the outline and of every embedded polygon (film boundary, holes), triangulate with
``scipy.spatial.Delaunay`` and flip every triangle to counter-clockwise order.  The same
``(sites, elements)`` arrays are then fed to the CUDA path and to the oracle.
"""
from __future__ import annotations

from typing import Iterable, Optional, Sequence, Tuple

import numpy as np
from scipy.spatial import Delaunay, cKDTree

from .geometry import box, circle, orient_ccw, points_in_polygon


def _resample_closed(poly: np.ndarray, spacing: float) -> np.ndarray:
    """Resamples a closed polygon so consecutive points are ~``spacing`` apart,
    keeping the original vertices."""
    poly = orient_ccw(poly)
    out = []
    nxt = np.roll(poly, -1, axis=0)
    for a, b in zip(poly, nxt):
        length = float(np.linalg.norm(b - a))
        k = max(1, int(round(length / spacing)))
        s = np.arange(k)[:, None] / k
        out.append(a[None, :] * (1 - s) + b[None, :] * s)
    return np.concatenate(out, axis=0)


def make_mesh(
    outline: np.ndarray,
    *,
    target_vertices: int,
    embedded: Sequence[np.ndarray] = (),
    seed: int = 0,
    jitter: float = 0.25,
) -> Tuple[np.ndarray, np.ndarray]:
    """Delaunay mesh of a convex ``outline`` with ~``target_vertices`` vertices.

    Args:
        outline: (k, 2) convex polygon; its (resampled) points become the mesh boundary.
        target_vertices: approximate number of vertices wanted.
        embedded: polygons (film boundary inside a buffered outline, holes) whose
            resampled points must be mesh vertices.
        seed: seed for the interior jitter.
        jitter: jitter amplitude in units of the lattice spacing (interior points only).

    Returns:
        ``sites`` (n, 2) float64 and ``elements`` (m, 3) int64, all triangles CCW.
    """
    outline = orient_ccw(outline)
    x0, y0 = outline.min(axis=0)
    x1, y1 = outline.max(axis=0)
    from .geometry import signed_area

    area = abs(signed_area(outline))
    # hexagonal lattice: area per point = h^2 * sqrt(3)/2
    h = np.sqrt(area / (target_vertices * np.sqrt(3.0) / 2.0))
    rng = np.random.default_rng(seed)
    constraint_polys = [_resample_closed(outline, h)] + [
        _resample_closed(p, h) for p in embedded
    ]
    fixed = np.concatenate(constraint_polys, axis=0)
    # lattice
    ny = int(np.ceil((y1 - y0) / (h * np.sqrt(3.0) / 2.0))) + 2
    nx = int(np.ceil((x1 - x0) / h)) + 2
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    xs = x0 + (ii + 0.5 * (jj % 2)) * h - 0.25 * h
    ys = y0 + jj * (h * np.sqrt(3.0) / 2.0) - 0.25 * h
    lattice = np.stack([xs.ravel(), ys.ravel()], axis=1)
    lattice = lattice + jitter * h * (rng.random(lattice.shape) * 2.0 - 1.0)
    keep = points_in_polygon(outline, lattice)
    lattice = lattice[keep]
    # drop lattice points too close to any constrained boundary point (avoids slivers)
    tree = cKDTree(fixed)
    d, _ = tree.query(lattice)
    lattice = lattice[d > 0.6 * h]
    sites = np.concatenate([fixed, lattice], axis=0)
    # unique (exact duplicates only)
    _, first = np.unique(np.round(sites / (1e-9 * max(h, 1e-300))).astype(np.int64), axis=0, return_index=True)
    sites = np.ascontiguousarray(sites[np.sort(first)], dtype=np.float64)
    tri = Delaunay(sites).simplices.astype(np.int64)
    # remove degenerate triangles and flip to CCW
    p = sites[tri]
    cross = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (
        p[:, 1, 1] - p[:, 0, 1]
    ) * (p[:, 2, 0] - p[:, 0, 0])
    good = np.abs(cross) > 1e-12 * h * h
    tri, cross = tri[good], cross[good]
    flip = cross < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    # drop unreferenced vertices (can happen after removing degenerate hull slivers)
    used = np.zeros(len(sites), dtype=bool)
    used[tri.ravel()] = True
    if not used.all():
        remap = -np.ones(len(sites), dtype=np.int64)
        remap[used] = np.arange(int(used.sum()))
        sites = np.ascontiguousarray(sites[used])
        tri = remap[tri]
    return sites, np.ascontiguousarray(tri, dtype=np.int64)


def square_mesh(side: float, target_vertices: int, seed: int = 0):
    """C2/C5: ``box(side)`` film, buffer 0 (mesh boundary == film boundary)."""
    return make_mesh(box(side, points=4), target_vertices=target_vertices, seed=seed)


def disk_mesh(radius: float, target_vertices: int, embedded: Iterable[np.ndarray] = (),
              seed: int = 0, center=(0.0, 0.0), boundary_points: Optional[int] = None):
    npts = boundary_points or max(24, int(2.2 * np.sqrt(np.pi * target_vertices)))
    return make_mesh(circle(radius, points=npts, center=center),
                     target_vertices=target_vertices, embedded=list(embedded), seed=seed)
