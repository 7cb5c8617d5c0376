"""Function-level FEM API (reference superscreen/fem.py): same names and return types, computed
on the device by ``scb_mesh_analyze`` / ``scb_mesh_build`` through a throw-away ``Mesh``."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .geometry import points_in_polygon
from .mesh import Mesh


def _mesh(points, triangles, weight_method="half_cotangent") -> Mesh:
    return Mesh.from_triangulation(points, triangles, weight_method=weight_method)


def triangle_areas(points, triangles) -> np.ndarray:
    """reference fem.py:10-29"""
    return _mesh(points, triangles).triangle_areas


def vertex_areas(points, triangles, tri_areas=None) -> np.ndarray:
    """reference device/utils.py:251-273"""
    return _mesh(points, triangles).vertex_areas


def centroids(points, triangles) -> np.ndarray:
    """reference fem.py:59-67"""
    return _mesh(points, triangles).triangle_centroids


def in_polygon(poly_points, query_points, radius: float = 0):
    """reference fem.py:32-56"""
    out = points_in_polygon(np.atleast_2d(poly_points), np.atleast_2d(query_points)).squeeze()
    return out.item() if out.ndim == 0 else out


def adjacency_matrix(triangles, sparse: bool = True):
    """reference fem.py:70-98"""
    n = int(np.max(triangles)) + 1
    pts = np.stack([np.arange(n, dtype=float), np.zeros(n)], axis=1)
    adj = _mesh(pts, triangles).adjacency_matrix()
    return adj if sparse else adj.toarray()


def adj_directed_tri_indices(triangles, num_sites: int):
    """reference fem.py:101-121 (returned as CSC like the reference)."""
    import scipy.sparse as sp

    pts = np.stack([np.arange(num_sites, dtype=float), np.zeros(num_sites)], axis=1)
    indptr, heads, tris = _mesh(pts, triangles).directed_star()
    return sp.csr_array((tris + 1, heads, indptr), shape=(num_sites, num_sites)).tocsc()


def _weights_from_laplacian(points, triangles, method: str, sparse: bool):
    """Off-diagonal weights W recovered from the device Laplacian L = M^-1 (W - diag(sum W)):
    W_ij = m_i L_ij for i != j (one rounding away from the reference's direct evaluation)."""
    import scipy.sparse as sp

    mesh = _mesh(points, triangles, method)
    lap = sp.csr_array(mesh.operators.laplacian)
    w = sp.csr_array(sp.diags(mesh.vertex_areas) @ lap)
    w.setdiag(0.0)
    w.eliminate_zeros()
    return w if sparse else w.toarray()


def weights_inv_euclidean(points, triangles, sparse: bool = True):
    """reference fem.py:124-162"""
    return _weights_from_laplacian(points, triangles, "inv_euclidean", sparse)


def weights_half_cotangent(points, triangles, sparse: bool = True):
    """reference fem.py:165-224"""
    return _weights_from_laplacian(points, triangles, "half_cotangent", sparse)


def calculate_weights(points, triangles, method: str, sparse: bool = True):
    """reference fem.py:227-256"""
    method = method.lower()
    if method == "uniform":
        return adjacency_matrix(triangles, sparse=sparse).astype(float)
    if method == "inv_euclidean":
        return weights_inv_euclidean(points, triangles, sparse=sparse)
    if method == "half_cotangent":
        return weights_half_cotangent(points, triangles, sparse=sparse)
    raise ValueError(
        f"Unknown method ({method}). "
        f"Supported methods are 'uniform', 'inv_euclidean', and 'half_cotangent'."
    )


def laplace_operator(points, triangles, masses: Optional[np.ndarray] = None, weight_method: str = "half_cotangent"):
    """reference fem.py:259-296.  The device builds ``inv(M) @ L`` with the lumped vertex areas as
    ``M``; user-supplied ``masses`` replace them by a row rescaling ``diag(w / masses)``."""
    import scipy.sparse as sp

    mesh = _mesh(points, triangles, weight_method)
    lap = mesh.operators.laplacian
    if masses is None:
        return lap
    masses = np.asarray(masses, dtype=np.float64)
    if masses.shape != (len(mesh.sites),):
        raise ValueError(f"masses must have shape ({len(mesh.sites)},), got {masses.shape}.")
    return sp.csr_array(sp.diags(mesh.vertex_areas / masses, format="csr") @ lap)


def gradient_triangles(points, triangles, areas=None) -> Tuple:
    """reference fem.py:299-347.  ``areas`` (pre-computed triangle areas) replace the device's own
    areas in the ``1 / (2 area)`` factor by a row rescaling."""
    import scipy.sparse as sp

    mesh = _mesh(points, triangles)
    ops = mesh.operators
    Gx, Gy = ops.gradient_tri_x, ops.gradient_tri_y
    if areas is None:
        return Gx, Gy
    areas = np.asarray(areas, dtype=np.float64)
    if areas.shape != (len(mesh.elements),):
        raise ValueError(f"areas must have shape ({len(mesh.elements)},), got {areas.shape}.")
    scale = sp.diags(mesh.triangle_areas / areas, format="csr")
    return sp.csr_array(scale @ Gx), sp.csr_array(scale @ Gy)


def gradient_vertices(points, triangles, gradient_tri=None, areas=None) -> Tuple:
    """reference fem.py:350-402.  The device evaluates the star average of its own triangle
    gradients; pre-computed ``gradient_tri`` / ``areas`` are accepted only when they agree with
    those (they are an optimisation hint in the reference, not a different operator)."""
    mesh = _mesh(points, triangles)
    ops = mesh.operators
    if areas is not None and not np.allclose(np.asarray(areas, dtype=float), mesh.triangle_areas, rtol=1e-12,
                                             atol=0.0):
        raise NotImplementedError("gradient_vertices: `areas` that differ from the mesh's triangle areas "
                                  "are not supported.")
    if gradient_tri is not None:
        for given, own in zip(gradient_tri, (ops.gradient_tri_x, ops.gradient_tri_y)):
            diff = abs(given - own)
            if diff.shape != own.shape or (diff.nnz and diff.max() > 1e-12 * abs(own).max()):
                raise NotImplementedError("gradient_vertices: a `gradient_tri` other than "
                                          "gradient_triangles(points, triangles) is not supported.")
    return ops.gradient_x, ops.gradient_y
