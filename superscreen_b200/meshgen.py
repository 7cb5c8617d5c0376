"""Mesh generation on the device: the stand-in for the reference's meshpy / Triangle call.

Reference: ``superscreen/device/utils.py:17-136`` (``generate_mesh``: polygon points + boundary facets
-> Delaunay mesh, refined until ``min_points`` / ``max_edge_length`` hold), called from
``Polygon.make_mesh`` (``device/polygon.py:192-224``) and ``Device.make_mesh``
(``device/device.py:383-471``).  meshpy is not installable here and its output cannot be reproduced
vertex by vertex (parity unpinned by construction: the mesh is an *input* of the solve path); what is
kept is the call signature and the contract --

* every input polygon point is a mesh vertex, the boundary ring(s) are mesh edges,
* the triangulation is the Delaunay triangulation of its vertices restricted to the region
  (checked against ``scipy.spatial.Delaunay`` in the tests: identical triangle sets),
* refinement: the point density is increased until ``len(points) >= min_points`` and the longest
  (interior) edge is ``<= max_edge_length`` (same loop as ``device/utils.py:113-135``).

The work runs on the GPU through the C ABI: ``scb_lattice_points`` (jittered hexagonal point cloud
clipped to the region and kept clear of the polygon points), ``scb_delaunay`` (one thread per point
clips its Voronoi cell in a uniform grid) and ``scb_points_in_rings`` (triangle centroids vs region).
torch is used for device memory and stream compaction only.
"""
from __future__ import annotations

import logging
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .geometry import signed_area

logger = logging.getLogger("superscreen_b200.meshgen")


def ensure_unique(coords: np.ndarray) -> np.ndarray:
    """Removes repeated points, keeping first occurrences in order (reference geometry.py ensure_unique)."""
    coords = np.asarray(coords, dtype=np.float64)
    _, ix = np.unique(coords, return_index=True, axis=0)
    return coords[np.sort(ix)]


def _open_ring(ring: np.ndarray) -> np.ndarray:
    ring = np.asarray(ring, dtype=np.float64)
    if len(ring) > 1 and np.array_equal(ring[0], ring[-1]):
        ring = ring[:-1]
    return ring


def _resample_ring(ring: np.ndarray, spacing: float) -> np.ndarray:
    """Inserts equally spaced points on every segment longer than ``spacing`` (original vertices kept)."""
    nxt = np.roll(ring, -1, axis=0)
    out = []
    for a, b in zip(ring, nxt):
        k = max(1, int(np.ceil(np.linalg.norm(b - a) / spacing - 1e-9)))
        s = np.arange(k)[:, None] / k
        out.append(a[None, :] * (1.0 - s) + b[None, :] * s)
    return np.concatenate(out, axis=0)


def convex_hull_ring(points: np.ndarray) -> np.ndarray:
    """Counter-clockwise convex hull (monotone chain)."""
    pts = np.unique(np.asarray(points, dtype=np.float64), axis=0)
    if len(pts) < 3:
        return pts

    def half(seq):
        h: List[np.ndarray] = []
        for p in seq:
            while len(h) >= 2 and ((h[-1][0] - h[-2][0]) * (p[1] - h[-2][1])
                                   - (h[-1][1] - h[-2][1]) * (p[0] - h[-2][0])) <= 0:
                h.pop()
            h.append(p)
        return h

    lower, upper = half(pts), half(pts[::-1])
    return np.array(lower[:-1] + upper[:-1])


def offset_convex_ring(ring: np.ndarray, distance: float) -> np.ndarray:
    """Moves every edge of a counter-clockwise convex ring outward by ``distance`` (mitre joins)."""
    ring = np.asarray(ring, dtype=np.float64)
    nxt = np.roll(ring, -1, axis=0)
    d = nxt - ring
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    normal = np.stack([d[:, 1], -d[:, 0]], axis=1)  # outward for a CCW ring
    a = ring + distance * normal                   # a point on every offset edge
    out = []
    for k in range(len(ring)):
        j = k - 1
        # intersection of offset edge j (a[j], d[j]) with offset edge k (a[k], d[k])
        det = d[j, 0] * (-d[k, 1]) - (-d[k, 0]) * d[j, 1]
        if abs(det) < 1e-12:
            out.append(a[k])
            continue
        rhs = a[k] - a[j]
        t = (rhs[0] * (-d[k, 1]) - (-d[k, 0]) * rhs[1]) / det
        out.append(a[j] + t * d[j])
    return np.array(out)


def _dev():
    import torch

    if not torch.cuda.is_available():
        raise _lib.SCBError("mesh generation runs on the GPU (libsc_b200): no CUDA device is available")
    return torch, torch.device(f"cuda:{torch.cuda.current_device()}")


def _upload_rings(rings: Sequence[np.ndarray]):
    torch, dev = _dev()
    ptr = np.zeros(len(rings) + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([len(r) for r in rings])
    verts = np.ascontiguousarray(np.concatenate(rings, axis=0), dtype=np.float64)
    return torch.as_tensor(ptr).to(dev), torch.as_tensor(verts).to(dev)


def points_in_rings(points, rings: Sequence[np.ndarray]):
    """Even-odd rule of (device or host) ``points`` against the closed ``rings``; returns a device bool tensor."""
    torch, dev = _dev()
    L = _lib.lib()
    pts = points if isinstance(points, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(points, dtype=np.float64))
    pts = pts.to(dev, torch.float64).contiguous()
    ring_ptr, ring_v = _upload_rings([_open_ring(r) for r in rings])
    inside = torch.empty(len(pts), dtype=torch.uint8, device=dev)
    _lib.check(L.scb_points_in_rings(len(pts), _lib.ptr(pts), len(rings), _lib.ptr(ring_ptr), _lib.ptr(ring_v),
                                     _lib.ptr(inside), _lib.stream_ptr()))
    return inside.bool()


def delaunay(points, cell: Optional[float] = None):
    """Delaunay triangulation of ``points`` ((n, 2), host array or device tensor) on the device.

    Returns a device ``int64`` tensor ``(m, 3)``: counter-clockwise triangles, smallest vertex first,
    ordered by that vertex and then counter-clockwise around it (a deterministic function of the points).
    Raises ``ValueError`` if a Voronoi cell overflowed its fixed-size buffers (wildly non-uniform point
    sets; not the quasi-uniform clouds this module produces)."""
    torch, dev = _dev()
    L = _lib.lib()
    pts = points if isinstance(points, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(points, dtype=np.float64))
    pts = pts.to(dev, torch.float64).contiguous()
    n = int(pts.shape[0])
    if n < 3:
        raise ValueError("at least three points are needed")
    lo, hi = pts.min(dim=0).values.cpu().numpy(), pts.max(dim=0).values.cpu().numpy()
    ext = np.maximum(hi - lo, 1e-300)
    if cell is None:  # about two points per grid cell
        cell = float(np.sqrt(2.0 * ext[0] * ext[1] / n)) if ext[0] * ext[1] > 0 else float(ext.max() / n)
        cell = max(cell, float(ext.max()) / 4096.0)
    x0, y0 = float(lo[0] - 0.5 * cell), float(lo[1] - 0.5 * cell)
    ncx, ncy = int(np.ceil(ext[0] / cell)) + 2, int(np.ceil(ext[1] / cell)) + 2
    max_tri = 2 * n
    tri = torch.empty(max_tri, 3, dtype=torch.int64, device=dev)
    info = torch.zeros(4, dtype=torch.int64, device=dev)
    _lib.check(L.scb_delaunay(n, _lib.ptr(pts), x0, y0, float(cell), ncx, ncy, max_tri, _lib.ptr(tri), _lib.ptr(info),
                              _lib.stream_ptr()))
    n_tri, cell_ovf, emit_ovf, _ = (int(v) for v in info.cpu().numpy())
    if cell_ovf or emit_ovf or n_tri > max_tri:
        raise ValueError(f"Delaunay triangulation failed: {cell_ovf} Voronoi cells / {emit_ovf} triangle lists "
                         f"overflowed their buffers ({n_tri} triangles for {n} points)")
    return tri[:n_tri]


def _edge_lengths(points_t, tri_t):
    """(unique undirected edges (e, 2), lengths (e,), is_boundary (e,)) of a triangulation, on the device."""
    import torch

    e = torch.cat([tri_t[:, [0, 1]], tri_t[:, [1, 2]], tri_t[:, [2, 0]]], dim=0)
    e = torch.sort(e, dim=1).values
    key = e[:, 0] * int(points_t.shape[0]) + e[:, 1]
    uniq, counts = torch.unique(key, return_counts=True)
    a, b = uniq // int(points_t.shape[0]), uniq % int(points_t.shape[0])
    length = (points_t[a] - points_t[b]).norm(dim=1)
    return torch.stack([a, b], dim=1), length, counts == 1


def _triangulate_region(points_t, rings: Sequence[np.ndarray], h: float):
    """Delaunay triangulation of the points cut down to the region: triangles whose centroid lies outside
    the rings (even-odd) and near-degenerate triangles are dropped."""
    import torch

    tri = delaunay(points_t)
    p = points_t[tri]
    cross = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    keep = cross > 1e-9 * h * h
    keep &= points_in_rings(p.mean(dim=1), rings)
    return tri[keep]


def generate_mesh(
    poly_coords: np.ndarray,
    hole_coords: Optional[List[np.ndarray]] = None,
    min_points: Optional[int] = None,
    max_edge_length: Optional[float] = None,
    convex_hull: bool = False,
    boundary: Optional[np.ndarray] = None,
    preserve_boundary: bool = False,
    min_angle: float = 32.5,
    seed: int = 0,
    jitter: float = 0.25,
    max_iterations: int = 60,
    embedded: Optional[List[np.ndarray]] = None,
    **kwargs,
) -> Tuple[np.ndarray, np.ndarray]:
    """Delaunay mesh for a set of polygon vertex coordinates (reference ``device/utils.py:17-136``).

    Args:
        poly_coords: ``(n, 2)`` polygon coordinates; all of them become mesh vertices.
        hole_coords: rings to cut out of the mesh.
        min_points: minimum number of mesh vertices.
        max_edge_length: maximum length of the mesh edges (interior edges if ``preserve_boundary``).
        convex_hull: mesh the convex hull of the points instead of the polygon.
        boundary: ``(m, 2)`` ring (a subset of ``poly_coords``) that bounds the mesh; default: the
            polygon ``poly_coords`` itself.
        preserve_boundary: do not add mesh vertices on the boundary.
        min_angle: accepted for signature compatibility; the point cloud is a jittered hexagonal
            lattice (angles around 60 degrees), no angle-driven Steiner insertion is done.  The smallest
            angle of the result is logged at debug level.
        seed, jitter: seed and amplitude (in units of the lattice spacing) of the lattice jitter.
        embedded: rings inside the region (a film outline inside a buffered boundary, holes that are meshed
            over) whose segments are resampled at the lattice spacing and added as mesh vertices, so that
            the polygon is resolved by mesh edges at every refinement level (the reference passes such
            polygons as bare points).

    Returns:
        ``points (n, 2) float64`` and ``triangles (m, 3) int64`` (counter-clockwise).
    """
    import torch

    torch_, dev = _dev()
    L = _lib.lib()
    poly = ensure_unique(_open_ring(poly_coords))
    holes = [ensure_unique(_open_ring(hc)) for hc in (hole_coords or [])]
    if convex_hull:
        if boundary is not None:
            raise ValueError("Cannot have both boundary is not None and convex_hull = True.")
        outer = convex_hull_ring(np.concatenate([poly] + holes, axis=0))
    elif boundary is not None:
        b = ensure_unique(_open_ring(boundary))
        bset = set(map(tuple, b))
        outer = np.array([p for p in poly if tuple(p) in bset])
        if len(outer) < 3:
            raise ValueError("boundary must consist of at least three of the polygon points.")
    else:
        outer = poly
    if signed_area(outer) < 0:
        outer = outer[::-1]
    rings = [outer] + holes
    area = abs(signed_area(outer)) - sum(abs(signed_area(hh)) for hh in holes)
    if not area > 0:
        raise ValueError("The region to be meshed has no area.")
    extent = np.ptp(np.concatenate(rings, axis=0), axis=0)

    def ring_spacing(r):
        return np.linalg.norm(np.roll(r, -1, axis=0) - r, axis=1)

    # initial point spacing: the polygon's own resolution (what Triangle's quality mesh grades to), capped
    # like the reference's first refinement step (max_volume = dx dy / 100); then from the targets
    h = float(np.median(np.concatenate([ring_spacing(r) for r in rings])))
    h = min(h, float(np.sqrt(extent[0] * extent[1] / 100.0 * 4.0 / np.sqrt(3.0))))
    if min_points:
        h = min(h, float(np.sqrt(area / (min_points * np.sqrt(3.0) / 2.0))))
    if max_edge_length is not None and max_edge_length > 0:
        h = min(h, float(max_edge_length) / 1.6)
    else:
        max_edge_length = np.inf
    min_points = int(min_points or 0)
    fixed_interior = np.concatenate([poly] + holes, axis=0)

    ring_ptr, ring_v = _upload_rings(rings)
    result = None
    for it in range(1, max_iterations + 1):
        # fixed points: every polygon point + (unless preserve_boundary) Steiner points on long ring segments
        if preserve_boundary:
            ring_pts = rings
        else:
            ring_pts = [_resample_ring(r, h) for r in rings]
        emb_pts = [_resample_ring(_open_ring(r), h) for r in (embedded or [])]
        fixed = ensure_unique(np.concatenate([fixed_interior] + list(ring_pts) + emb_pts, axis=0))
        fixed_t = torch.as_tensor(np.ascontiguousarray(fixed)).to(dev)
        lo = np.min(outer, axis=0)
        ny = int(np.ceil(extent[1] / (h * np.sqrt(3.0) / 2.0))) + 2
        nx = int(np.ceil(extent[0] / h)) + 2
        lattice = torch.empty(nx * ny, 2, dtype=torch.float64, device=dev)
        keep = torch.empty(nx * ny, dtype=torch.uint8, device=dev)
        _lib.check(L.scb_lattice_points(nx, ny, float(lo[0] - 0.25 * h), float(lo[1] - 0.25 * h), float(h), float(jitter),
                                        int(seed) + 7919 * (it - 1), len(rings), _lib.ptr(ring_ptr), _lib.ptr(ring_v),
                                        len(fixed), _lib.ptr(fixed_t), 0.6 * float(h), _lib.ptr(lattice), _lib.ptr(keep),
                                        _lib.stream_ptr()))
        points_t = torch.cat([fixed_t, lattice[keep.bool()]], dim=0).contiguous()
        tri_t = _triangulate_region(points_t, rings, h)
        edges, lengths, is_boundary = _edge_lengths(points_t, tri_t)
        used = torch.zeros(len(points_t), dtype=torch.bool, device=dev)
        used[tri_t.reshape(-1)] = True
        n_used = int(used.sum().item())
        sel = lengths[~is_boundary] if preserve_boundary else lengths
        max_length = float(sel.max().item()) if len(sel) else 0.0
        logger.debug(f"Iteration {it}: made mesh with {n_used} points and {len(tri_t)} triangles with maximum "
                     f"edge length {max_length:.2e} (target {max_edge_length:.2e}), lattice spacing {h:.3e}.")
        result = (points_t, tri_t, used, n_used)
        if n_used >= min_points and max_length <= max_edge_length:
            break
        # the reference shrinks Triangle's max_volume by min(0.98, sqrt(max_edge / max_len)) per pass; here
        # the lattice spacing is scaled directly (edge lengths ~ h, vertex count ~ 1 / h^2)
        shrink = 0.98
        if np.isfinite(max_edge_length) and max_length > max_edge_length:
            shrink = min(shrink, 0.98 * float(max_edge_length / max_length))
        if n_used < min_points:
            shrink = min(shrink, 0.98 * float(np.sqrt(max(n_used, 1) / min_points)))
        h *= shrink
    else:
        raise RuntimeError(f"generate_mesh did not reach min_points={min_points} / max_edge_length={max_edge_length} "
                           f"in {max_iterations} iterations")
    points_t, tri_t, used, n_used = result
    if n_used < len(points_t):  # drop unreferenced points (cannot be polygon points of a valid region)
        remap = torch.cumsum(used.to(torch.int64), dim=0) - 1
        tri_t = remap[tri_t]
        points_t = points_t[used]
    points = points_t.cpu().numpy()
    triangles = np.ascontiguousarray(tri_t.cpu().numpy(), dtype=np.int64)
    if logger.isEnabledFor(logging.DEBUG):
        logger.debug(f"smallest angle {min_triangle_angle(points, triangles):.1f} degrees (min_angle={min_angle})")
    return points, triangles


def min_triangle_angle(points: np.ndarray, triangles: np.ndarray) -> float:
    """Smallest interior angle (degrees) over all triangles."""
    p = points[triangles]
    best = 180.0
    for k in range(3):
        a, b, c = p[:, k], p[:, (k + 1) % 3], p[:, (k + 2) % 3]
        u, v = b - a, c - a
        cosang = np.einsum("ij,ij->i", u, v) / (np.linalg.norm(u, axis=1) * np.linalg.norm(v, axis=1))
        best = min(best, float(np.degrees(np.arccos(np.clip(cosang, -1.0, 1.0))).min()))
    return best


def boundary_is_conforming(points: np.ndarray, triangles: np.ndarray, rings: Sequence[np.ndarray]) -> bool:
    """True if the boundary edges of the triangulation are exactly the polyline(s) ``rings``
    (every boundary edge lies on a ring segment and the boundary length equals the ring length)."""
    e = np.concatenate([triangles[:, [0, 1]], triangles[:, [1, 2]], triangles[:, [2, 0]]], axis=0)
    e.sort(axis=1)
    uniq, counts = np.unique(e, axis=0, return_counts=True)
    bedges = uniq[counts == 1]
    blen = float(np.linalg.norm(points[bedges[:, 0]] - points[bedges[:, 1]], axis=1).sum())
    rlen = sum(float(np.linalg.norm(np.roll(_open_ring(r), -1, axis=0) - _open_ring(r), axis=1).sum()) for r in rings)
    if abs(blen - rlen) > 1e-9 * rlen:
        return False
    mid = 0.5 * (points[bedges[:, 0]] + points[bedges[:, 1]])
    dmin = np.full(len(mid), np.inf)
    for r in rings:
        r = _open_ring(r)
        a, b = r, np.roll(r, -1, axis=0)
        ab = b - a
        t = np.clip(np.einsum("mij,ij->mi", mid[:, None, :] - a[None], ab) / np.einsum("ij,ij->i", ab, ab)[None], 0, 1)
        d = np.linalg.norm(mid[:, None, :] - (a[None] + t[..., None] * ab[None]), axis=2).min(axis=1)
        dmin = np.minimum(dmin, d)
    return bool(dmin.max() <= 1e-9 * max(rlen, 1e-300))
