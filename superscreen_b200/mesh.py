"""Mesh + MeshOperators built on the B200 (reference: superscreen/device/mesh.py:110-155,
331-458, device/edge_mesh.py:38-63, device/utils.py:139-166,230-273, fem.py).

``Mesh.from_triangulation(sites, elements)`` uploads the triangulation once and runs
``scb_mesh_analyze`` + ``scb_mesh_build`` + ``scb_kernel_diagonal``; every array the reference
exposes (``boundary_indices``, ``vertex_areas``, ``operators.laplacian`` ...) is available as a
lazily downloaded numpy / scipy object, while the solver consumes the device-resident copies.
The dense kernel matrix ``Q`` (n x n) is never formed on the solve path; ``operators.Q``
materialises it on the device only when a caller asks for it.
"""
from __future__ import annotations

from typing import Dict, Optional, Union

import numpy as np

from . import _lib

WEIGHT_METHODS = {"half_cotangent": 0, "uniform": 1, "inv_euclidean": 2}


def _torch():
    import torch

    return torch


def boundary_vertices_ccw(elements: np.ndarray) -> np.ndarray:
    """Boundary vertex indices of a consistently CCW-oriented triangulation, ordered counter-
    clockwise along the (single) outer boundary, starting at the smallest boundary vertex index.
    Stands in for ``device.utils.boundary_vertices`` (reference device/utils.py:205-227, which
    needs matplotlib + shapely and starts at a shapely-defined vertex): the ordering is an INPUT of
    the transport-terminal branch, handed identically to the oracle / the reference."""
    el = np.asarray(elements, dtype=np.int64)
    a = el.ravel()
    b = el[:, [1, 2, 0]].ravel()
    n = int(el.max()) + 1
    key = a * n + b
    rev = b * n + a
    is_boundary = ~np.isin(key, rev)           # directed edges whose reverse is in no triangle
    ea, eb = a[is_boundary], b[is_boundary]
    nxt = dict(zip(ea.tolist(), eb.tolist()))
    if len(nxt) != len(ea):
        raise ValueError("mesh boundary is not a simple loop (a boundary vertex has two outgoing edges)")
    start = int(ea.min())
    loop = [start]
    v = nxt[start]
    while v != start:
        loop.append(v)
        v = nxt[v]
        if len(loop) > len(ea):
            raise ValueError("mesh boundary is not a single closed loop")
    if len(loop) != len(ea):
        raise ValueError("mesh has more than one boundary loop")
    return np.asarray(loop, dtype=np.int64)


class DeviceMeshData:
    """Device-resident arrays of one mesh (owned by torch, filled through the C ABI)."""

    def __init__(self, sites: np.ndarray, elements: np.ndarray, weight_method: str = "half_cotangent",
                 device: Optional[str] = None):
        torch = _torch()
        L = _lib.lib()
        if weight_method not in WEIGHT_METHODS:
            raise ValueError(
                f"Unknown method ({weight_method}). "
                f"Supported methods are 'uniform', 'inv_euclidean', and 'half_cotangent'."
            )
        dev = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        self.device = dev
        n, m = int(sites.shape[0]), int(elements.shape[0])
        self.n, self.m = n, m
        self.weight_method = weight_method
        with torch.cuda.device(dev):
            s = _lib.stream_ptr()
            if isinstance(sites, torch.Tensor):  # already resident (bench: device-resident leg)
                self.sites = sites.to(dev, torch.float64).contiguous()
                self.elements = elements.to(dev, torch.int64).contiguous()
            else:
                self.sites = torch.as_tensor(np.ascontiguousarray(sites, dtype=np.float64)).to(dev, non_blocking=True)
                self.elements = torch.as_tensor(np.ascontiguousarray(elements, dtype=np.int64)).to(dev, non_blocking=True)
            ws = torch.empty(int(L.scb_mesh_workspace_elems(n, m)), dtype=torch.int32, device=dev)
            counts = torch.empty(4, dtype=torch.int64, device=dev)
            flags = torch.empty(1, dtype=torch.int32, device=dev)
            _lib.check(L.scb_mesh_analyze(n, m, _lib.ptr(self.elements), _lib.ptr(ws), _lib.ptr(counts),
                                          _lib.ptr(flags), s))
            host = torch.cat([counts, flags.to(torch.int64)]).cpu().numpy()
            nnz_adj, n_edges, n_bv, n_be, bad = (int(v) for v in host)
            if bad:
                raise ValueError(
                    "The triangulation is not a consistently oriented (counter-clockwise) "
                    "manifold mesh: duplicated directed edges or degenerate triangles found."
                )
            self.n_edges, self.n_boundary_vertices, self.n_boundary_edges = n_edges, n_bv, n_be
            f64 = dict(dtype=torch.float64, device=dev)
            i32 = dict(dtype=torch.int32, device=dev)
            i64 = dict(dtype=torch.int64, device=dev)
            nnz_op = nnz_adj + n
            t = {
                "triangle_areas": torch.empty(m, **f64), "vertex_areas": torch.empty(n, **f64),
                "centroids": torch.empty(m, 2, **f64), "C": torch.empty(n, **f64),
                "adj_indptr": torch.empty(n + 1, **i32), "adj_indices": torch.empty(nnz_adj, **i32),
                "edges": torch.empty(n_edges, 2, **i64),
                "edge_is_boundary": torch.empty(n_edges, dtype=torch.uint8, device=dev),
                "boundary_indices": torch.empty(n_bv, **i64),
                "star_indptr": torch.empty(n + 1, **i32), "star_heads": torch.empty(3 * m, **i32),
                "star_tris": torch.empty(3 * m, **i32),
                "edge_centers": torch.empty(n_edges, 2, **f64), "edge_directions": torch.empty(n_edges, 2, **f64),
                "edge_lengths": torch.empty(n_edges, **f64),
                "op_indptr": torch.empty(n + 1, **i32), "op_indices": torch.empty(nnz_op, **i32),
                "laplacian": torch.empty(nnz_op, **f64), "gradient_x": torch.empty(nnz_op, **f64),
                "gradient_y": torch.empty(nnz_op, **f64),
                "gtri_indices": torch.empty(3 * m, **i32), "gtri_x": torch.empty(3 * m, **f64),
                "gtri_y": torch.empty(3 * m, **f64),
            }
            out = _lib.MeshOut(**{k: _lib.ptr(v) for k, v in t.items()})
            _lib.check(L.scb_mesh_build(n, m, _lib.ptr(self.sites), _lib.ptr(self.elements), _lib.ptr(ws),
                                        WEIGHT_METHODS[weight_method], out, s))
            self.t: Dict[str, "torch.Tensor"] = t
            # Q_ii * w_i = C_i + sum_j q_ij w_j  (device/mesh.py:456), an n^2 pair sum
            self.qdw = torch.empty(n, **f64)
            _lib.check(L.scb_kernel_diagonal(n, _lib.ptr(self.sites), _lib.ptr(t["vertex_areas"]),
                                             _lib.ptr(t["C"]), _lib.ptr(self.qdw), s))
            self.gtri_indptr = torch.arange(0, 3 * m + 1, 3, **i32)
        self._host: Dict[str, np.ndarray] = {}

    def host(self, name: str) -> np.ndarray:
        if name not in self._host:
            src = self.qdw if name == "qdw" else self.t[name]
            self._host[name] = src.cpu().numpy()
        return self._host[name]


class EdgeMesh:
    """reference device/edge_mesh.py:9-63 (arrays downloaded lazily)."""

    def __init__(self, data: DeviceMeshData):
        self._d = data

    centers = property(lambda self: self._d.host("edge_centers"))
    edges = property(lambda self: self._d.host("edges"))
    directions = property(lambda self: self._d.host("edge_directions"))
    edge_lengths = property(lambda self: self._d.host("edge_lengths"))
    is_boundary = property(lambda self: self._d.host("edge_is_boundary").astype(bool))

    @property
    def boundary_edge_indices(self) -> np.ndarray:
        return np.where(self.is_boundary)[0].astype(np.int64)


class MeshOperators:
    """reference device/mesh.py:331-458.  Sparse operators are scipy CSR views of the device
    arrays; ``Q`` is materialised on demand only."""

    def __init__(self, data: DeviceMeshData):
        self._d = data
        self._cache = {}

    def _csr(self, key: str, shape):
        import scipy.sparse as sp

        if key not in self._cache:
            d = self._d
            if key.startswith("gradient_tri"):
                data = d.host("gtri_x" if key.endswith("x") else "gtri_y")
                m = sp.csr_array((data, d.host("gtri_indices"), d.gtri_indptr.cpu().numpy()), shape=shape)
            else:
                m = sp.csr_array((d.host(key), d.host("op_indices"), d.host("op_indptr")), shape=shape)
            self._cache[key] = m
        return self._cache[key]

    @property
    def weights(self) -> np.ndarray:
        return self._d.host("vertex_areas")

    @property
    def laplacian(self):
        return self._csr("laplacian", (self._d.n, self._d.n))

    @property
    def gradient_x(self):
        return self._csr("gradient_x", (self._d.n, self._d.n))

    @property
    def gradient_y(self):
        return self._csr("gradient_y", (self._d.n, self._d.n))

    @property
    def gradient_tri_x(self):
        return self._csr("gradient_tri_x", (self._d.m, self._d.n))

    @property
    def gradient_tri_y(self):
        return self._csr("gradient_tri_y", (self._d.m, self._d.n))

    @property
    def C(self) -> np.ndarray:
        return self._d.host("C")

    @property
    def Q_diagonal(self) -> np.ndarray:
        """diag(Q) without forming Q."""
        return self._d.host("qdw") / self.weights

    def Q_device(self):
        """Dense Q (n x n) on the device: Q_ij = -q_ij, Q_ii = qdw_i / w_i (device/mesh.py:454-458)."""
        torch = _torch()
        L = _lib.lib()
        d = self._d
        n = d.n
        n_pad = -(-n // 128) * 128
        with torch.cuda.device(d.device):
            M = torch.empty(n_pad, n_pad, dtype=torch.float64, device=d.device)
            ix = torch.arange(n, dtype=torch.int64, device=d.device)
            pos = torch.empty(n, dtype=torch.int32, device=d.device)
            zeros = torch.zeros(n, dtype=torch.float64, device=d.device)
            _lib.check(L.scb_system_assemble(
                n, _lib.ptr(d.sites), _lib.ptr(d.t["vertex_areas"]), _lib.ptr(d.qdw), None, _lib.ptr(zeros),
                _lib.ptr(d.t["op_indptr"]), _lib.ptr(d.t["op_indices"]), _lib.ptr(d.t["laplacian"]), None,
                n, _lib.ptr(ix), _lib.ptr(pos), n_pad, _lib.ptr(M), None, None, _lib.stream_ptr()))
            # M = -(Q * w[None, :])
            return -(M[:n, :n] / d.t["vertex_areas"][None, :])

    @property
    def Q(self) -> np.ndarray:
        if "Q" not in self._cache:
            self._cache["Q"] = self.Q_device().cpu().numpy()
        return self._cache["Q"]

    @staticmethod
    def from_mesh(mesh: "Mesh") -> "MeshOperators":
        return MeshOperators(mesh._data)

    @staticmethod
    def C_vector(points: np.ndarray) -> np.ndarray:
        """Edge vector C for arbitrary points (reference device/mesh.py:400-432)."""
        torch = _torch()
        L = _lib.lib()
        pts = np.ascontiguousarray(points, dtype=np.float64)
        dev = torch.device(f"cuda:{torch.cuda.current_device()}")
        with torch.cuda.device(dev):
            p = torch.as_tensor(pts).to(dev)
            scratch = torch.empty(8, dtype=torch.float64, device=dev)
            C = torch.empty(len(pts), dtype=torch.float64, device=dev)
            _lib.check(L.scb_c_vector(len(pts), _lib.ptr(p), _lib.ptr(scratch), _lib.ptr(C), _lib.stream_ptr()))
            return C.cpu().numpy()

    @staticmethod
    def Q_matrix(points: np.ndarray, weights: np.ndarray) -> np.ndarray:
        """Dense kernel matrix Q for arbitrary points / weights (reference device/mesh.py:434-458):
        Q = -q off the diagonal, Q_ii = (C_i + sum_j q_ij w_j) / w_i."""
        from .distance import q_matrix

        q = q_matrix(points)
        C = MeshOperators.C_vector(points)
        w = np.asarray(weights, dtype=np.float64)
        diag = (C + q @ w) / w
        Q = -q
        np.fill_diagonal(Q, diag)
        return Q

    def copy(self) -> "MeshOperators":
        return self


class Mesh:
    """A triangular mesh whose operators live on the GPU (reference device/mesh.py:17-155)."""

    def __init__(self, data: DeviceMeshData, sites: np.ndarray, elements: np.ndarray,
                 build_operators: bool = True):
        self._data = data
        self.sites = sites
        self.elements = elements
        self.edge_mesh = EdgeMesh(data)
        self.operators: Optional[MeshOperators] = MeshOperators(data) if build_operators else None

    @staticmethod
    def from_triangulation(sites, elements, build_operators: bool = True,
                           weight_method: str = "half_cotangent", device: Optional[str] = None) -> "Mesh":
        sites = np.asarray(sites).squeeze()
        elements = np.asarray(elements).squeeze()
        if sites.ndim != 2 or sites.shape[1] != 2:
            raise ValueError(f"The site coordinates must have shape (n, 2), got {sites.shape!r}")
        if elements.ndim != 2 or elements.shape[1] != 3:
            raise ValueError(f"The elements must have shape (m, 3), got {elements.shape!r}.")
        sites = np.ascontiguousarray(sites, dtype=np.float64)
        elements = np.ascontiguousarray(elements, dtype=np.int64)
        if elements.min() < 0 or elements.max() >= len(sites):
            raise ValueError("elements reference vertices outside of sites")
        with _lib.nvtx_range("scb.mesh_operators"):
            data = DeviceMeshData(sites, elements, weight_method=weight_method, device=device)
        return Mesh(data, sites, elements, build_operators=build_operators)

    # lazily downloaded arrays (same names as the reference attributes)
    boundary_indices = property(lambda self: self._data.host("boundary_indices"))
    vertex_areas = property(lambda self: self._data.host("vertex_areas"))
    triangle_areas = property(lambda self: self._data.host("triangle_areas"))
    triangle_centroids = property(lambda self: self._data.host("centroids"))

    @staticmethod
    def find_boundary_indices(elements: np.ndarray) -> np.ndarray:
        n = int(np.max(elements)) + 1
        return Mesh.from_triangulation(np.zeros((n, 2)) + np.arange(n)[:, None], elements).boundary_indices

    def adjacency_matrix(self):
        import scipy.sparse as sp

        d = self._data
        idx = d.host("adj_indices")
        return sp.csr_array((np.ones(len(idx), dtype=int), idx, d.host("adj_indptr")), shape=(d.n, d.n))

    def directed_star(self):
        """(indptr, heads, tris): row form of fem.adj_directed_tri_indices (fem.py:101-121)."""
        d = self._data
        return d.host("star_indptr"), d.host("star_heads"), d.host("star_tris")

    def smooth(self, iterations: int, build_operators: bool = True) -> "Mesh":
        """Laplacian smoothing: every interior vertex moves to the arithmetic mean of its
        neighbours, ``iterations`` times (reference device/mesh.py:172-211).  The sweeps run on the
        device (``scb_mesh_smooth``) with the reference's summation order; the topology is
        unchanged, so only the final mesh is rebuilt."""
        if iterations <= 0:
            return self
        torch = _torch()
        L = _lib.lib()
        d = self._data
        with torch.cuda.device(d.device):
            cur = d.sites
            nb = int(d.t["boundary_indices"].numel())
            for _ in range(int(iterations)):
                new = torch.empty_like(cur)
                _lib.check(L.scb_mesh_smooth(d.n, _lib.ptr(cur), _lib.ptr(d.t["adj_indptr"]),
                                             _lib.ptr(d.t["adj_indices"]), nb, _lib.ptr(d.t["boundary_indices"]),
                                             _lib.ptr(new), _lib.stream_ptr()))
                cur = new
            new_sites = cur.cpu().numpy()
        return Mesh.from_triangulation(new_sites, self.elements, build_operators=build_operators,
                                       weight_method=d.weight_method, device=str(d.device))

    def stats(self) -> Dict[str, Union[int, float]]:
        el = self.edge_mesh.edge_lengths
        va = self.vertex_areas
        return dict(num_sites=len(self.sites), num_elements=len(self.elements),
                    min_edge_length=el.min(), max_edge_length=el.max(),
                    min_vertex_area=va.min(), max_vertex_area=va.max())

    def closest_site(self, xy) -> int:
        return int(np.argmin(np.linalg.norm(self.sites - np.atleast_2d(xy), axis=1)))

    def to_hdf5(self, h5group, compress: bool = True) -> None:
        """reference device/mesh.py:250-275 (``sites``, ``elements``; the derived arrays too if not ``compress``)"""
        from . import io as _io

        _io.mesh_to_hdf5(self, h5group, compress=compress)

    @staticmethod
    def from_hdf5(h5group) -> "Mesh":
        """reference device/mesh.py:277-293; the device-resident operators are rebuilt from the triangulation."""
        from . import io as _io

        return _io.mesh_from_hdf5(h5group)

    def copy(self) -> "Mesh":
        return self
