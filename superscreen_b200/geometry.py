"""Small host-side geometry helpers used by the solve hot path and its tests.

Only what the path touches is provided (the reference's shapely/matplotlib geometry
subsystem is out of scope): polygon factories, an even-odd
point-in-polygon test standing in for ``matplotlib.path.Path.contains_points``
(reference: superscreen/device/polygon.py:138-162, superscreen/fem.py:32-56) and
``path_vectors`` (reference: superscreen/geometry.py:160-182).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def circle(radius: float, points: int = 100, center=(0.0, 0.0)) -> np.ndarray:
    """Counter-clockwise polygon approximating a circle (open: first != last)."""
    t = np.linspace(0.0, 2.0 * np.pi, points, endpoint=False)
    xy = radius * np.stack([np.cos(t), np.sin(t)], axis=1)
    return xy + np.asarray(center, dtype=float)


def ellipse(a: float, b: float, points: int = 100, center=(0.0, 0.0)) -> np.ndarray:
    t = np.linspace(0.0, 2.0 * np.pi, points, endpoint=False)
    xy = np.stack([a * np.cos(t), b * np.sin(t)], axis=1)
    return xy + np.asarray(center, dtype=float)


def box(width: float, height: float = None, points: int = 101, center=(0.0, 0.0)) -> np.ndarray:
    """Counter-clockwise rectangle with ``points`` vertices spread along the perimeter."""
    if height is None:
        height = width
    per_side = max(1, int(points) // 4)
    x0, y0 = -width / 2.0, -height / 2.0
    s = np.arange(per_side) / per_side
    bottom = np.stack([x0 + width * s, np.full(per_side, y0)], axis=1)
    right = np.stack([np.full(per_side, x0 + width), y0 + height * s], axis=1)
    top = np.stack([x0 + width - width * s, np.full(per_side, y0 + height)], axis=1)
    left = np.stack([np.full(per_side, x0), y0 + height - height * s], axis=1)
    xy = np.concatenate([bottom, right, top, left], axis=0)
    return xy + np.asarray(center, dtype=float)


def close_curve(points: np.ndarray) -> np.ndarray:
    """Appends the first point if the curve is not closed (reference geometry.py:185-196)."""
    points = np.asarray(points, dtype=float)
    if not np.array_equal(points[0], points[-1]):
        points = np.concatenate([points, points[:1]], axis=0)
    return points


def signed_area(points: np.ndarray) -> float:
    p = np.asarray(points, dtype=float)
    x, y = p[:, 0], p[:, 1]
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def orient_ccw(points: np.ndarray) -> np.ndarray:
    """Open polygon ring, counter-clockwise."""
    p = np.asarray(points, dtype=float)
    if len(p) > 1 and np.array_equal(p[0], p[-1]):
        p = p[:-1]
    if signed_area(p) < 0:
        p = p[::-1]
    return np.ascontiguousarray(p)


def path_vectors(path: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Edge lengths and unit normals of a path (reference geometry.py:160-182)."""
    dr = np.diff(np.asarray(path, dtype=float), axis=0)
    normals = np.stack([dr[:, 1], -dr[:, 0]], axis=1)  # == np.cross(dr, [0, 0, 1])[:, :2]
    edge_lengths = np.linalg.norm(dr, axis=1)
    unit_normals = normals / edge_lengths[:, np.newaxis]
    return edge_lengths, unit_normals


def points_in_polygon(poly_points: np.ndarray, query_points: np.ndarray) -> np.ndarray:
    """Even-odd rule point-in-polygon, vectorised over the query points.

    Stands in for ``matplotlib.path.Path.contains_points`` (radius=0).  Points exactly on
    an edge are implementation-defined in matplotlib; here an edge is
    half-open in y, which makes the result deterministic.  Index sets derived from this
    are *inputs* to both the CUDA path and the oracle.
    """
    poly = np.asarray(poly_points, dtype=float)
    if len(poly) > 1 and np.array_equal(poly[0], poly[-1]):
        poly = poly[:-1]
    q = np.atleast_2d(np.asarray(query_points, dtype=float))
    x1, y1 = poly[:, 0], poly[:, 1]
    x2, y2 = np.roll(x1, -1), np.roll(y1, -1)
    dx, dy = x2 - x1, y2 - y1
    inside = np.zeros(len(q), dtype=bool)
    chunk = max(1, (1 << 22) // max(1, len(poly)))  # bound the (points x edges) temporaries
    for s in range(0, len(q), chunk):
        x, y = q[s:s + chunk, 0][:, None], q[s:s + chunk, 1][:, None]
        cond = (y1[None, :] > y) != (y2[None, :] > y)
        # same operation order as the scalar formula (x2-x1)*(y-y1)/(y2-y1)+x1: mesh vertices that
        # lie exactly ON a polygon vertex/edge must classify reproducibly (golden index sets)
        with np.errstate(divide="ignore", invalid="ignore"):
            xint = dx[None, :] * (y - y1[None, :]) / dy[None, :] + x1[None, :]
            crossing = cond & (x < xint)
        inside[s:s + chunk] = np.count_nonzero(crossing, axis=1) & 1
    return inside
