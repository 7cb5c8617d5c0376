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
    """reference fem.py:259-296 (``masses`` are always the lumped vertex areas)."""
    return _mesh(points, triangles, weight_method).operators.laplacian


def gradient_triangles(points, triangles, areas=None) -> Tuple:
    """reference fem.py:299-347"""
    ops = _mesh(points, triangles).operators
    return ops.gradient_tri_x, ops.gradient_tri_y


def gradient_vertices(points, triangles, gradient_tri=None, areas=None) -> Tuple:
    """reference fem.py:350-402"""
    ops = _mesh(points, triangles).operators
    return ops.gradient_x, ops.gradient_y
