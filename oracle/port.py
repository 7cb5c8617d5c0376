"""CPU oracle: a numpy/scipy(/numba) restatement of the reference's solve hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``superscreen_b200/`` imports this module; it is
used by ``tests/``, by ``__graft_entry__.smoke()`` as the checker and by ``bench.py`` for the
``cpu_baseline`` / ``--impl reference`` legs.  Every function cites the reference
(loganbvh/superscreen 0.13.0) file:line it restates.

Parity pinning: the reference ships no golden vectors for this path.
This port is pinned instead (a) against the UNMODIFIED reference executed live in the build
container (``tests/test_oracle_vs_reference.py``, via ``oracle/reference_loader.py``) and
(b) against fixtures generated from that live reference and committed under ``tests/golden/``
(``oracle/make_golden.py`` is the generating script).  The LU itself lives in an un-vendored
dependency (scipy.linalg -> LAPACK dgetrf/dgetrs in the OpenBLAS bundled with scipy; the
reference pins no version, container has scipy 1.18.1 / OpenBLAS 0.3.30); the oracle calls the
same scipy entry points at the reference's call sites (solver/solve_film.py:279,530).

Where the reference uses Python ``for`` loops (``vertex_areas`` device/utils.py:268-273,
``gradient_vertices`` fem.py:386-401) the port is vectorised; this changes floating-point
summation order at the 1e-16 level only and makes the CPU baseline *faster* than the
reference, never slower.  The dense parts (q_matrix, Q_matrix, _build_system_2d,
lu_factor, lu_solve, biot_savart_*) follow the reference operation for operation.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.linalg as la
import scipy.sparse as sp

try:  # the reference's kernels are numba-jitted (distance.py:87, solve.py:28, current.py:13)
    import numba

    _HAVE_NUMBA = True
except Exception:  # pragma: no cover
    numba = None
    _HAVE_NUMBA = False

MU_0 = 1.25663706212e-06  # N/A^2 (pint's mu_0, CODATA 2018); see SURVEY.md Q7
PHI_0 = 2.067833848461929e-15  # Wb
ONE_OVER_4PI = 1.0 / (4.0 * np.pi)


# ----------------------------------------------------------------------------------------
# mesh integer structures (a5)
# ----------------------------------------------------------------------------------------
def get_edges(triangles: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Unique sorted edges and boundary flags (reference device/utils.py:139-152)."""
    edges = np.concatenate([triangles[:, e] for e in [(0, 1), (1, 2), (2, 0)]])
    edges = np.sort(edges, axis=1)
    edges, counts = np.unique(edges, return_counts=True, axis=0)
    return edges, counts == 1


def find_boundary_indices(elements: np.ndarray) -> np.ndarray:
    """Ascending boundary vertex ids (reference device/mesh.py:157-170)."""
    edges, is_boundary = get_edges(elements)
    return np.unique(edges[is_boundary].ravel())


def adjacency_matrix(triangles: np.ndarray) -> sp.csr_array:
    """Symmetric 0/1 adjacency (reference fem.py:70-98)."""
    edges = np.concatenate(
        [triangles[:, [0, 1]], triangles[:, [1, 2]], triangles[:, [2, 0]]]
    )
    row, col = edges[:, 0], edges[:, 1]
    n = int(max(row.max(), col.max())) + 1
    adj = sp.csr_array((np.ones_like(row, dtype=int), (row, col)), shape=(n, n))
    adj = adj + adj.T
    return (adj > 0).astype(int)


def adj_directed_tri_indices(triangles: np.ndarray, num_sites: int) -> sp.csc_array:
    """Directed-edge -> (triangle id + 1) map (reference fem.py:101-121)."""
    t0, t1, t2 = triangles[:, 0], triangles[:, 1], triangles[:, 2]
    i = np.column_stack([t0, t1, t2]).ravel()
    j = np.column_stack([t1, t2, t0]).ravel()
    data = np.repeat(np.arange(1, triangles.shape[0] + 1), 3)
    return sp.csc_array((data, (i, j)), shape=(num_sites, num_sites))


def directed_star(triangles: np.ndarray, num_sites: int):
    """Row-wise form of ``adj_directed_tri_indices(...).tolil()`` used by
    ``gradient_vertices`` (reference fem.py:386-391): for vertex i, the triangles that own a
    directed edge i->j, ordered by head vertex j.  Returns CSR-like (indptr, heads, tris)."""
    t = np.asarray(triangles, dtype=np.int64)
    tail = t.ravel()
    head = t[:, [1, 2, 0]].ravel()
    tri = np.repeat(np.arange(len(t), dtype=np.int64), 3)
    order = np.lexsort((head, tail))
    tail, head, tri = tail[order], head[order], tri[order]
    indptr = np.zeros(num_sites + 1, dtype=np.int64)
    np.add.at(indptr, tail + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr, head, tri


# ----------------------------------------------------------------------------------------
# mesh float structures (a4, a6-a8)
# ----------------------------------------------------------------------------------------
def smooth_sites(sites: np.ndarray, elements: np.ndarray, iterations: int) -> np.ndarray:
    """Laplacian smoothing of the vertex positions (reference device/mesh.py:172-211): each sweep
    moves every vertex to the mean of its neighbours (two bincount passes over the sorted edge
    list, then a division by the neighbour count) and restores the boundary vertices."""
    edges, _ = get_edges(elements)
    n = len(sites)
    boundary = find_boundary_indices(elements)
    num_neighbors = np.bincount(edges.ravel(), minlength=n)
    sites = np.asarray(sites, dtype=np.float64)
    for _ in range(iterations):
        new = np.zeros((n, 2))
        new += np.array([np.bincount(edges[:, 0], v, minlength=n) for v in sites[edges[:, 1]].T]).T
        new += np.array([np.bincount(edges[:, 1], v, minlength=n) for v in sites[edges[:, 0]].T]).T
        new /= num_neighbors[:, np.newaxis]
        new[boundary] = sites[boundary]
        sites = new
    return sites


def triangle_areas(points: np.ndarray, triangles: np.ndarray) -> np.ndarray:
    """Signed triangle areas (reference device/utils.py:230-248)."""
    xy = points[triangles]
    s = xy[:, [2, 0]] - xy[:, [1, 2]]
    return np.linalg.det(s) * 0.5


def vertex_areas(points, triangles, tri_areas=None) -> np.ndarray:
    """Lumped-mass vertex areas (reference device/utils.py:251-273).  ``np.add.at`` applies
    the additions in (triangle, local vertex) order == the reference's Python loop order."""
    if tri_areas is None:
        tri_areas = triangle_areas(points, triangles)
    v = np.zeros(len(points), dtype=float)
    np.add.at(v, np.asarray(triangles).ravel(), np.repeat(tri_areas / 3, 3))
    return v


def edge_mesh(sites: np.ndarray, elements: np.ndarray) -> Dict[str, np.ndarray]:
    """EdgeMesh.from_mesh (reference device/edge_mesh.py:38-63)."""
    edges, is_boundary = get_edges(elements)
    coords = sites[edges]
    directions = np.diff(coords, axis=1).squeeze()
    return dict(
        edges=edges,
        is_boundary=is_boundary,
        boundary_edge_indices=np.where(is_boundary)[0],
        centers=coords.mean(axis=1),
        directions=directions,
        edge_lengths=np.linalg.norm(directions, axis=1),
    )


def _half_cot(points, triangles, k):
    """0.5*cot(angle at local vertex k) (reference fem.py:188-222)."""
    a, b = (k + 1) % 3, (k + 2) % 3
    v1 = points[triangles[:, a]] - points[triangles[:, k]]
    v2 = points[triangles[:, b]] - points[triangles[:, k]]
    return 0.5 / np.tan(
        np.arccos(
            np.sum(v1 * v2, axis=1) / (la.norm(v1, axis=1) * la.norm(v2, axis=1))
        )
    )


def calculate_weights(points, triangles, method: str) -> sp.csr_array:
    """Edge weight matrix W (reference fem.py:124-256)."""
    method = method.lower()
    n = points.shape[0]
    t = triangles
    if method == "uniform":
        return adjacency_matrix(t).astype(float)
    if method == "inv_euclidean":
        # assignment (not accumulation): each undirected edge gets 1/length (fem.py:148-160)
        rows, cols, vals = [], [], []
        for a, b in ((0, 1), (0, 2), (1, 2)):
            inv = 1 / la.norm(points[t[:, b]] - points[t[:, a]], axis=1)
            rows += [t[:, a], t[:, b]]
            cols += [t[:, b], t[:, a]]
            vals += [inv, inv]
        rows, cols, vals = map(np.concatenate, (rows, cols, vals))
        # duplicates carry identical values up to rounding; "last write wins" in the
        # reference -> keep one deterministic representative
        key = rows * n + cols
        _, first = np.unique(key[::-1], return_index=True)
        sel = len(key) - 1 - first
        return sp.csr_array((vals[sel], (rows[sel], cols[sel])), shape=(n, n))
    if method == "half_cotangent":
        rows, cols, vals = [], [], []
        for k in range(3):
            a, b = [(1, 2), (0, 2), (0, 1)][k]
            w = _half_cot(points, t, k)
            rows += [t[:, a], t[:, b]]
            cols += [t[:, b], t[:, a]]
            vals += [w, w]
        rows, cols, vals = map(np.concatenate, (rows, cols, vals))
        return sp.csr_array((vals, (rows, cols)), shape=(n, n))  # sums duplicates
    raise ValueError(
        f"Unknown method ({method}). "
        f"Supported methods are 'uniform', 'inv_euclidean', and 'half_cotangent'."
    )


def laplace_operator(points, triangles, masses=None, weight_method="half_cotangent") -> sp.csr_array:
    """inv(M) @ (W - diag(sum W)) (reference fem.py:259-296)."""
    if masses is None:
        masses = vertex_areas(points, triangles)
    W = calculate_weights(points, triangles, weight_method).tocsr()
    W.setdiag(0)
    W.sort_indices()
    rowsum = np.asarray(W.sum(axis=1)).ravel()
    L = (W - sp.diags(rowsum, format="csr")).tocsr()
    lap = (sp.diags(1 / masses, format="csr") @ L).tocsr()
    lap.sort_indices()
    return lap


def gradient_triangles(points, triangles, areas=None):
    """Per-triangle gradient operators Gx, Gy (reference fem.py:299-347)."""
    if areas is None:
        areas = triangle_areas(points, triangles)
    xy = points[triangles]
    edges = np.roll(xy, 2, axis=1) - np.roll(xy, 1, axis=1)
    rot = np.empty_like(edges)
    rot[:, :, 0] = +edges[:, :, 1]
    rot[:, :, 1] = -edges[:, :, 0]
    tri_data = (rot / (2 * areas[:, None, None])).reshape(-1, 2).T
    shape = (triangles.shape[0], points.shape[0])
    row = np.repeat(np.arange(len(triangles)), 3)
    col = triangles.ravel()
    Gx = sp.csr_array((tri_data[0], (row, col)), shape=shape, dtype=float)
    Gy = sp.csr_array((tri_data[1], (row, col)), shape=shape, dtype=float)
    return Gx, Gy


def gradient_vertices(points, triangles, gradient_tri=None, areas=None):
    """Vertex gradient operators (reference fem.py:350-402), vectorised.

    Quirk Q1: each adjacent triangle is weighted by the angle at its
    own *local vertex 0* (fem.py:393-398), not by the angle at vertex i."""
    if gradient_tri is None:
        Gx, Gy = gradient_triangles(points, triangles, areas=areas)
    else:
        Gx, Gy = gradient_tri
    n = len(points)
    indptr, _, tris = directed_star(triangles, n)
    v1 = points[triangles[:, 1]] - points[triangles[:, 0]]
    v2 = points[triangles[:, 2]] - points[triangles[:, 0]]
    theta = np.arccos(
        np.einsum("ij, ij -> i", v1, v2) / (la.norm(v1, axis=1) * la.norm(v2, axis=1))
    )
    th = theta[tris]
    rows = np.repeat(np.arange(n), np.diff(indptr))
    tot = np.zeros(n)
    np.add.at(tot, rows, th)
    omega = th / tot[rows]
    S = sp.csr_array((omega, (rows, tris)), shape=(n, len(triangles)))
    gx = (S @ Gx).tocsr()
    gy = (S @ Gy).tocsr()
    gx.sort_indices()
    gy.sort_indices()
    return gx, gy


# ----------------------------------------------------------------------------------------
# kernel matrix (a1-a3)
# ----------------------------------------------------------------------------------------
if _HAVE_NUMBA:

    @numba.njit(fastmath=True, parallel=True)
    def q_matrix(points):  # reference distance.py:87-115
        n = points.shape[0]
        out = np.empty((n, n), dtype=points.dtype)
        c = 1 / (4 * np.pi)
        for i in numba.prange(n):
            for j in range(n):
                if i == j:
                    out[i, j] = 0.0
                else:
                    out[i, j] = c * (
                        (points[i, 0] - points[j, 0]) ** 2
                        + (points[i, 1] - points[j, 1]) ** 2
                    ) ** (-1.5)
        return out

    @numba.njit(fastmath=True, parallel=True)
    def biot_savart_film_to_film(film1_sites, film1_z0, film1_areas, film1_J, film2_sites, film2_z0):
        # reference solver/solve.py:28-73
        c = 1 / (4 * np.pi)
        out = np.empty(film2_sites.shape[0], dtype=film1_J.dtype)
        dz2 = (film2_z0 - film1_z0) ** 2
        for i in numba.prange(film2_sites.shape[0]):
            tmp = 0.0
            for j in range(film1_sites.shape[0]):
                dx = film2_sites[i, 0] - film1_sites[j, 0]
                dy = film2_sites[i, 1] - film1_sites[j, 1]
                tmp += (
                    c
                    * film1_areas[j]
                    * (film1_J[j, 0] * dy - film1_J[j, 1] * dx)
                    * (dx**2 + dy**2 + dz2) ** (-1.5)
                )
            out[i] = tmp
        return out

    @numba.njit(fastmath=True, parallel=True)
    def biot_savart_2d_z_kernel(eval_positions, positions, current_densities, areas, pref0):
        # reference sources/current.py:13-57 (pref0 = mu_0/4pi)
        Jx = current_densities[:, 0]
        Jy = current_densities[:, 1]
        out = np.empty(len(eval_positions), dtype=np.float64)
        for i in numba.prange(eval_positions.shape[0]):
            a = 0.0
            b = 0.0
            for k in range(positions.shape[0]):
                dx = eval_positions[i, 0] - positions[k, 0]
                dy = eval_positions[i, 1] - positions[k, 1]
                dz = eval_positions[i, 2] - positions[k, 2]
                pref = pref0 * areas[k] * (dx * dx + dy * dy + dz * dz) ** (-1.5)
                a += pref * Jx[k] * dy
                b += pref * Jy[k] * dx
            out[i] = a - b
        return out

    @numba.njit(fastmath=True, parallel=True)
    def biot_savart_2d_vector_kernel(eval_positions, positions, current_densities, areas, pref0):
        # reference sources/current.py:60-110
        Jx = current_densities[:, 0]
        Jy = current_densities[:, 1]
        out = np.empty((len(eval_positions), 3), dtype=np.float64)
        for i in numba.prange(eval_positions.shape[0]):
            a = 0.0
            b = 0.0
            c = 0.0
            d = 0.0
            for k in range(positions.shape[0]):
                dx = eval_positions[i, 0] - positions[k, 0]
                dy = eval_positions[i, 1] - positions[k, 1]
                dz = eval_positions[i, 2] - positions[k, 2]
                pref = pref0 * areas[k] * (dx * dx + dy * dy + dz * dz) ** (-1.5)
                a += pref * Jx[k] * dy
                b += pref * Jy[k] * dx
                c += pref * Jx[k] * dz
                d += pref * Jy[k] * dz
            out[i, 0] = d
            out[i, 1] = -c
            out[i, 2] = a - b
        return out

else:  # pragma: no cover - numpy fallbacks, chunked to bound memory

    def q_matrix(points):
        n = points.shape[0]
        out = np.empty((n, n), dtype=points.dtype)
        for s in range(0, n, 1024):
            d = points[s : s + 1024, None, :] - points[None, :, :]
            r2 = d[..., 0] ** 2 + d[..., 1] ** 2
            with np.errstate(divide="ignore"):
                out[s : s + 1024] = ONE_OVER_4PI * r2 ** (-1.5)
        np.fill_diagonal(out, 0.0)
        return out

    def biot_savart_film_to_film(film1_sites, film1_z0, film1_areas, film1_J, film2_sites, film2_z0):
        out = np.empty(film2_sites.shape[0])
        dz2 = (film2_z0 - film1_z0) ** 2
        for s in range(0, len(out), 1024):
            dx = film2_sites[s : s + 1024, None, 0] - film1_sites[None, :, 0]
            dy = film2_sites[s : s + 1024, None, 1] - film1_sites[None, :, 1]
            out[s : s + 1024] = np.sum(
                ONE_OVER_4PI * film1_areas * (film1_J[:, 0] * dy - film1_J[:, 1] * dx)
                * (dx**2 + dy**2 + dz2) ** (-1.5), axis=1)
        return out

    def biot_savart_2d_vector_kernel(eval_positions, positions, current_densities, areas, pref0):
        out = np.empty((len(eval_positions), 3))
        Jx, Jy = current_densities[:, 0], current_densities[:, 1]
        for s in range(0, len(out), 1024):
            d = eval_positions[s : s + 1024, None, :] - positions[None, :, :]
            pref = pref0 * areas * np.sum(d * d, axis=2) ** (-1.5)
            out[s : s + 1024, 0] = np.sum(pref * Jy * d[..., 2], axis=1)
            out[s : s + 1024, 1] = -np.sum(pref * Jx * d[..., 2], axis=1)
            out[s : s + 1024, 2] = np.sum(pref * (Jx * d[..., 1] - Jy * d[..., 0]), axis=1)
        return out

    def biot_savart_2d_z_kernel(eval_positions, positions, current_densities, areas, pref0):
        return biot_savart_2d_vector_kernel(eval_positions, positions, current_densities, areas, pref0)[:, 2]


def C_vector(points: np.ndarray) -> np.ndarray:
    """Edge-correction vector (reference device/mesh.py:400-432)."""
    x = points[:, 0]
    y = points[:, 1]
    x = x - x.mean()
    y = y - y.mean()
    a = np.ptp(x) / 2
    b = np.ptp(y) / 2
    with np.errstate(divide="ignore"):
        C = sum(
            np.sqrt((a - p * x) ** (-2) + (b - q * y) ** (-2))
            for p, q in itertools.product((-1, 1), repeat=2)
        )
    C[np.isinf(C)] = 1e30
    C /= 4 * np.pi
    return C


def Q_matrix(points: np.ndarray, weights: np.ndarray) -> np.ndarray:
    """Kernel matrix Q (reference device/mesh.py:434-458)."""
    q = q_matrix(points)
    C = C_vector(points)
    diag = -(C + np.einsum("ij, j -> i", q, weights)) / weights
    np.fill_diagonal(q, diag)
    return -q


# ----------------------------------------------------------------------------------------
# mesh operator bundle
# ----------------------------------------------------------------------------------------
@dataclass
class OracleMesh:
    """What Mesh.from_triangulation + MeshOperators.from_mesh hold
    (reference device/mesh.py:110-155,361-394)."""

    sites: np.ndarray
    elements: np.ndarray
    boundary_indices: np.ndarray
    triangle_areas: np.ndarray
    vertex_areas: np.ndarray
    triangle_centroids: np.ndarray
    laplacian: sp.csr_array
    gradient_tri_x: sp.csr_array
    gradient_tri_y: sp.csr_array
    gradient_x: sp.csr_array
    gradient_y: sp.csr_array
    Q: Optional[np.ndarray] = None

    @property
    def weights(self):
        return self.vertex_areas


def build_mesh(sites, elements, with_Q: bool = True, weight_method="half_cotangent") -> OracleMesh:
    sites = np.asarray(sites, dtype=np.float64)
    elements = np.asarray(elements, dtype=np.int64)
    ta = triangle_areas(sites, elements)
    va = vertex_areas(sites, elements, ta)
    Gx, Gy = gradient_triangles(sites, elements, ta)
    gx, gy = gradient_vertices(sites, elements, gradient_tri=(Gx, Gy))
    lap = laplace_operator(sites, elements, va, weight_method=weight_method)
    return OracleMesh(
        sites=sites,
        elements=elements,
        boundary_indices=find_boundary_indices(elements),
        triangle_areas=ta,
        vertex_areas=va,
        triangle_centroids=sites[elements].mean(axis=1),
        laplacian=lap,
        gradient_tri_x=Gx,
        gradient_tri_y=Gy,
        gradient_x=gx,
        gradient_y=gy,
        Q=Q_matrix(sites, va) if with_Q else None,
    )


# ----------------------------------------------------------------------------------------
# per-film linear systems (a9-a12)
# ----------------------------------------------------------------------------------------
def lambda_inhomogeneous(Lambda: np.ndarray) -> bool:
    """reference solver/utils.py:44-47"""
    return bool(
        np.ptp(Lambda) / max(np.min(np.abs(Lambda)), np.finfo(float).eps) > 1e-6
    )


def grad_Lambda_term_sparse(mesh: OracleMesh, Lambda: np.ndarray) -> sp.csr_array:
    """T_jk = sum_d (grad_d Lambda)_j (grad_d)_jk (reference solve_film.py:181-185),
    kept sparse (the reference densifies to (2, n, n))."""
    gLx = mesh.gradient_x @ Lambda
    gLy = mesh.gradient_y @ Lambda
    return (sp.diags(gLx) @ mesh.gradient_x + sp.diags(gLy) @ mesh.gradient_y).tocsr()


def build_system_2d(Q, weights, Lambda, laplacian_dense, grad_Lambda_term, ix1d):
    """reference solve_film.py:296-305 (Lambda is (n,), broadcast over columns, Q3)."""
    ix2d = np.ix_(ix1d, ix1d)
    gl = grad_Lambda_term[ix2d] if isinstance(grad_Lambda_term, np.ndarray) else 0
    return Q[ix2d] * weights[ix1d] - Lambda[ix1d] * laplacian_dense[ix2d] - gl


def build_system_1d(Q, weights, Lambda, laplacian_dense, grad_Lambda_term, ix):
    """reference solve_film.py:285-293"""
    gl = grad_Lambda_term[:, ix] if isinstance(grad_Lambda_term, np.ndarray) else 0
    return Q[:, ix] * weights[ix] - Lambda[ix] * laplacian_dense[:, ix] - gl


@dataclass
class OracleFilm:
    """FilmInfo + LinearSystem for one film (reference solver/utils.py:96-132,
    solve_film.py:18-35)."""

    name: str
    mesh: OracleMesh
    z0: float
    Lambda: np.ndarray  # (n,)
    interior_indices: np.ndarray
    hole_indices: Dict[str, np.ndarray]
    indices: np.ndarray = None  # film system indices (interior minus holes)
    A: np.ndarray = None
    lu_piv: tuple = None
    hole_A: Dict[str, np.ndarray] = field(default_factory=dict)
    inhomogeneous: bool = False
    film_polygon: Optional[np.ndarray] = None
    # transport terminals (reference solve_film.py:220-263): CCW-ordered boundary vertex ids, per
    # terminal the ascending positions (into boundary_ordered) of the boundary vertices it contains
    boundary_ordered: Optional[np.ndarray] = None
    terminals: Optional[Dict[str, np.ndarray]] = None
    boundary_A: Optional[np.ndarray] = None
    indices_with_holes: Optional[np.ndarray] = None
    lu_piv_with_holes: tuple = None


def factorize_film(film: OracleFilm, keep_A: bool = True) -> OracleFilm:
    """factorize_linear_systems for a film without terminals
    (reference solve_film.py:151-218,264-282)."""
    mesh = film.mesh
    Q = mesh.Q if mesh.Q is not None else Q_matrix(mesh.sites, mesh.vertex_areas)
    w = mesh.vertex_areas
    lap = mesh.laplacian.toarray()  # reference solver/utils.py:292
    film.inhomogeneous = lambda_inhomogeneous(film.Lambda)
    glt = 0
    if film.inhomogeneous:
        glt = grad_Lambda_term_sparse(mesh, film.Lambda).toarray()
    film.hole_A = {
        name: build_system_1d(Q, w, film.Lambda, lap, glt, ix)
        for name, ix in film.hole_indices.items()
    }
    ix = film.interior_indices
    if film.terminals is not None:
        # reference solve_film.py:220-263: boundary slab + LU of the interior INCLUDING holes
        film.boundary_A = build_system_1d(Q, w, film.Lambda, lap, glt, film.boundary_ordered)
        film.indices_with_holes = ix
        film.lu_piv_with_holes = la.lu_factor(-build_system_2d(Q, w, film.Lambda, lap, glt, ix))
    if film.hole_indices:
        ix = np.setdiff1d(ix, np.concatenate(list(film.hole_indices.values())))
    if film.terminals is not None:
        ix = np.setdiff1d(ix, film.boundary_ordered)
    film.indices = ix
    A = build_system_2d(Q, w, film.Lambda, lap, glt, ix)
    film.lu_piv = la.lu_factor(-A)
    film.A = A if keep_A else None
    return film


def path_vectors(path: np.ndarray):
    """reference geometry.py:160-182 (edge lengths and unit normals)"""
    dr = np.diff(path, axis=0)
    normals = np.stack([dr[:, 1], -dr[:, 0]], axis=1)  # cross(dr, z)
    edge_lengths = la.norm(dr, axis=1)
    return edge_lengths, normals / edge_lengths[:, np.newaxis]


def stream_from_terminal_current(points: np.ndarray, current: float) -> np.ndarray:
    """reference solver/utils.py:440-488"""
    from scipy import integrate

    edge_lengths, unit_normals = path_vectors(points)
    J = current * unit_normals / np.sum(edge_lengths)
    zhat_cross_J = J[:, [1, 0]]
    zhat_cross_J[:, 0] *= -1
    dl = np.diff(points, axis=0)
    integrand = np.sum(zhat_cross_J * dl, axis=1)
    g = integrate.cumulative_trapezoid(integrand, initial=0)
    return g * current / g[-1]


def solve_for_terminal_current_stream(film: "OracleFilm", terminal_currents: Dict[str, float]) -> np.ndarray:
    """reference solve_film.py:308-390"""
    mesh = film.mesh
    points, weights = mesh.sites, mesh.vertex_areas
    npoints = len(points)
    if not any(terminal_currents.values()):
        return np.zeros(npoints)
    boundary_indices = film.boundary_ordered
    g = np.zeros(npoints)
    for name, ix_boundary in film.terminals.items():
        current = terminal_currents[name]
        ix_boundary = np.sort(ix_boundary)
        remaining_boundary = boundary_indices[ix_boundary[-1]:]
        ix_terminal = boundary_indices[ix_boundary]
        stream = stream_from_terminal_current(points[ix_terminal], -current)
        g[ix_terminal[:-1]] += stream
        g[remaining_boundary] += stream[-1]
    g = g - np.max(g) + np.ptp(g) / 2
    Ha_eff = -(film.boundary_A @ g[boundary_indices])
    ixa = film.indices_with_holes
    g[ixa] = la.lu_solve(film.lu_piv_with_holes, -Ha_eff[ixa])
    if len(film.hole_indices) == 0:
        return g
    Ha_eff = np.zeros(npoints)
    for name, ix in film.hole_indices.items():
        g[ix] = np.average(g[ix], weights=weights[ix])
        Ha_eff += -(film.hole_A[name] @ g[ix])
    Ha_eff += -(film.boundary_A @ g[boundary_indices])
    ix = film.indices
    g[ix] = la.lu_solve(film.lu_piv, -Ha_eff[ix])
    return g


def boundary_effective_field(sites, centers, lengths, normals, stream) -> np.ndarray:
    """reference solve_film.py:393-412 (vectorised)"""
    out = np.zeros(len(sites))
    for s in range(0, len(sites), 2048):
        dr = sites[s:s + 2048, None, :] - centers[None, :, :]
        r3 = np.sum(dr * dr, axis=2) ** 1.5
        out[s:s + 2048] = np.sum(stream / r3 * np.sum(dr * -normals[None, :, :], axis=2) * lengths, axis=1)
    return out / (4 * np.pi)


def biot_savart_within_film(sites, centroids, areas, J_tri) -> np.ndarray:
    """reference solve_film.py:415-437 (vectorised)"""
    out = np.zeros(len(sites))
    for s in range(0, len(sites), 1024):
        dx = sites[s:s + 1024, None, 0] - centroids[None, :, 0]
        dy = sites[s:s + 1024, None, 1] - centroids[None, :, 1]
        pref = areas * (dx * dx + dy * dy) ** (-1.5)
        out[s:s + 1024] = np.sum(pref * J_tri[:, 0] * dy, axis=1) - np.sum(pref * J_tri[:, 1] * dx, axis=1)
    return out / (4 * np.pi)


# ----------------------------------------------------------------------------------------
# solve (a13, a15, a16)
# ----------------------------------------------------------------------------------------
@dataclass
class OracleFilmSolution:
    stream: np.ndarray
    current_density: np.ndarray
    applied_field: np.ndarray
    self_field: np.ndarray
    field_from_other_films: Optional[np.ndarray] = None

    @property
    def total_field(self):
        t = self.applied_field + self.self_field
        if self.field_from_other_films is not None:
            t = t + self.field_from_other_films
        return t


def solve_film(
    film: OracleFilm,
    applied_field: np.ndarray,
    circulating_currents: Dict[str, float],
    field_conversion: float,
    vortices: Sequence[Tuple[float, float, float]] = (),
    vortex_flux: float = PHI_0 / MU_0 * 1e12,
    field_from_other_films: Optional[np.ndarray] = None,
    terminal_currents: Optional[Dict[str, float]] = None,
) -> OracleFilmSolution:
    """reference solve_film.py:440-574 (transport-terminal branch when film.terminals is set)."""
    mesh = film.mesh
    w = mesh.vertex_areas
    Q = mesh.Q
    Hz = applied_field
    if field_from_other_films is not None:
        Hz = Hz + field_from_other_films
    g = np.zeros_like(Hz)
    Ha_eff = np.zeros_like(Hz)
    for name, ix in film.hole_indices.items():
        current = circulating_currents.get(name, 0)
        g[ix] += current
        Ha_eff += -(film.hole_A[name] @ g[ix])
    if film.terminals is not None:
        # reference solve_film.py:505-524
        g_transport = solve_for_terminal_current_stream(film, terminal_currents or {})
        g += g_transport
        b = film.boundary_ordered
        boundary_sites = mesh.sites[b]
        boundary_stream = g_transport[b]
        centers = 0.5 * (boundary_sites + np.roll(boundary_sites, -1, axis=0))
        boundary_stream = 0.5 * (boundary_stream + np.roll(boundary_stream, -1, axis=0))
        closed = np.concatenate([boundary_sites, boundary_sites[:1]], axis=0)
        lengths, normals = path_vectors(closed)
        Ha_eff += boundary_effective_field(mesh.sites, centers, lengths, normals, boundary_stream)
    ix = film.indices
    h = Hz[ix] - Ha_eff[ix]
    gf = la.lu_solve(film.lu_piv, h)
    g[ix] += gf
    K = None
    for (vx, vy, nphi0) in vortices:
        if K is None:
            K = -la.lu_solve(film.lu_piv, np.eye(len(ix)))
        j_film = np.argmin(la.norm(mesh.sites[ix] - (vx, vy), axis=1))
        j_dev = np.argmin(la.norm(mesh.sites - (vx, vy), axis=1))
        g[ix] += vortex_flux * nphi0 * K[:, j_film] / w[j_dev]
    J = np.array([mesh.gradient_y @ g, -(mesh.gradient_x @ g)]).T
    if film.terminals is not None:
        # reference solve_film.py:557-562
        J_tri = np.array([mesh.gradient_tri_y @ g, -(mesh.gradient_tri_x @ g)]).T
        screening = biot_savart_within_film(mesh.sites, mesh.triangle_centroids, mesh.triangle_areas, J_tri)
    else:
        screening = Q @ (w * g)
    other = None
    if field_from_other_films is not None:
        other = field_from_other_films / field_conversion
    return OracleFilmSolution(
        stream=g,
        current_density=J,
        applied_field=applied_field / field_conversion,
        self_field=screening / field_conversion,
        field_from_other_films=other,
    )


def field_conversion_mT_to_uA_per_um() -> float:
    """reference solver/utils.py:407-437 for the default units (mT -> uA/um)."""
    return 1e-3 / MU_0  # A/m == uA/um


def solve(
    films: List[OracleFilm],
    applied_field: Callable[[np.ndarray, np.ndarray, np.ndarray], np.ndarray],
    circulating_currents: Optional[Dict[str, float]] = None,
    vortices: Optional[Dict[str, Sequence[Tuple[float, float, float]]]] = None,
    iterations: int = 0,
    field_conversion: Optional[float] = None,
    vortex_flux: float = PHI_0 / MU_0 * 1e12,
) -> List[Dict[str, OracleFilmSolution]]:
    """Driver + film-to-film Jacobi loop (reference solver/solve.py:406-549)."""
    circulating_currents = circulating_currents or {}
    vortices = vortices or {}
    conv = field_conversion_mT_to_uA_per_um() if field_conversion is None else field_conversion
    applied = {}
    for f in films:
        s = f.mesh.sites
        z0 = f.z0 * np.ones(len(s))
        applied[f.name] = np.squeeze(applied_field(s[:, 0], s[:, 1], z0) * conv).astype(np.float64)

    def run(other):
        return {
            f.name: solve_film(
                f, applied[f.name], circulating_currents, conv,
                vortices=vortices.get(f.name, ()), vortex_flux=vortex_flux,
                field_from_other_films=None if other is None else other[f.name],
            )
            for f in films
        }

    sols = [run(None)]
    if len(films) < 2 or iterations < 1:
        return sols
    for _ in range(iterations):
        other = {f.name: np.zeros(len(f.mesh.sites)) for f in films}
        for src, dst in itertools.product(films, repeat=2):
            if src is dst:
                continue
            other[dst.name] += biot_savart_film_to_film(
                src.mesh.sites, float(src.z0), src.mesh.vertex_areas,
                np.ascontiguousarray(sols[-1][src.name].current_density),
                dst.mesh.sites, float(dst.z0),
            )
        sols.append(run(other))
    return sols


# ----------------------------------------------------------------------------------------
# field evaluation (a17) and fluxoid (a19)
# ----------------------------------------------------------------------------------------
def biot_savart_2d(x, y, z, *, positions, current_densities, z0, areas,
                   to_meter=1e-6, to_amp_per_meter=1.0, vector=True, mu_0=MU_0):
    """reference sources/current.py:113-196 with explicit unit factors (um, uA/um)."""
    x, y, z = np.atleast_1d(x, y, z)
    if z.shape[0] == 1:
        z = z * np.ones_like(x)
    ev = np.array([x, y, z]).T * to_meter
    positions, current_densities = np.atleast_2d(positions, current_densities)
    J = current_densities * to_amp_per_meter
    pos = positions * to_meter
    zz = z0 * np.ones(len(pos)) * to_meter
    ar = areas * to_meter**2
    pos = np.concatenate([pos, zz[:, None]], axis=1)
    pref0 = mu_0 / (4 * np.pi)
    if vector:
        return biot_savart_2d_vector_kernel(ev, pos, J, ar, pref0)
    return biot_savart_2d_z_kernel(ev, pos, J, ar, pref0)


def linear_tri_interp(sites, elements, values, xy):
    """Barycentric linear interpolation (stands in for matplotlib LinearTriInterpolator,
    reference solution.py:272-312).  Brute force, for polygon-sized query sets only.
    Returns NaN outside the mesh."""
    xy = np.atleast_2d(xy)
    p = sites[elements]
    a, b, c = p[:, 0], p[:, 1], p[:, 2]
    det = (b[:, 1] - c[:, 1]) * (a[:, 0] - c[:, 0]) + (c[:, 0] - b[:, 0]) * (a[:, 1] - c[:, 1])
    out = np.full((len(xy),) + values.shape[1:], np.nan)
    for q, (x, y) in enumerate(xy):
        l1 = ((b[:, 1] - c[:, 1]) * (x - c[:, 0]) + (c[:, 0] - b[:, 0]) * (y - c[:, 1])) / det
        l2 = ((c[:, 1] - a[:, 1]) * (x - c[:, 0]) + (a[:, 0] - c[:, 0]) * (y - c[:, 1])) / det
        l3 = 1 - l1 - l2
        m = np.minimum(np.minimum(l1, l2), l3)
        t = int(np.argmax(m))
        if m[t] < -1e-12:
            continue
        v = values[elements[t]]
        out[q] = l1[t] * v[0] + l2[t] * v[1] + l3[t] * v[2]
    return out


def polygon_fluxoid(film: OracleFilm, sol: OracleFilmSolution, polygon: np.ndarray,
                    in_polygon_fn, field_units_to_tesla=1e-3, length_to_m=1e-6,
                    current_to_A=1e-6, mu_0=MU_0):
    """Fluxoid = flux part + supercurrent part, in Phi_0
    (reference solution.py:484-563).  ``polygon`` must be closed and CCW (the reference
    re-wraps the coordinates in a Polygon, device/polygon.py:67-77)."""
    mesh = film.mesh
    ix = in_polygon_fn(polygon, mesh.sites)
    flux = np.einsum("i, i ->", sol.total_field[ix], mesh.vertex_areas[ix])
    flux_part = flux * field_units_to_tesla * length_to_m**2 / PHI_0
    J_poly = linear_tri_interp(mesh.sites, mesh.elements, sol.current_density, polygon)
    if film.film_polygon is not None:
        J_poly[~in_polygon_fn(film.film_polygon, polygon)] = 0
    J_poly[~np.isfinite(J_poly).all(axis=1)] = 0
    Lambda_poly = linear_tri_interp(mesh.sites, mesh.elements, film.Lambda, polygon) \
        if lambda_inhomogeneous(film.Lambda) else np.full(len(polygon), film.Lambda[0])
    dl = np.diff(polygon, axis=0)
    int_J = np.trapezoid(Lambda_poly[:-1] * np.sum(J_poly[:-1] * dl, axis=1))
    # [uA/um * um^2] -> A m ; mu_0 * that = Wb
    super_part = mu_0 * int_J * (current_to_A / length_to_m) * length_to_m**2 / PHI_0
    return flux_part, super_part
