"""reference superscreen/distance.py: q_matrix on the device."""
from __future__ import annotations

import numpy as np


def q_matrix(points: np.ndarray) -> np.ndarray:
    """q_ij = 1 / (4 pi |r_i - r_j|^3), zero diagonal (reference distance.py:87-115).  Dense
    (n, n) output; the solve path never calls this (it is matrix-free)."""
    import torch

    from . import _lib

    L = _lib.lib()
    points = np.ascontiguousarray(points, dtype=np.float64)
    assert points.ndim == 2 and points.shape[1] == 2
    n = len(points)
    n_pad = -(-n // 128) * 128
    dev = torch.device(f"cuda:{torch.cuda.current_device()}")
    with torch.cuda.device(dev):
        sites = torch.as_tensor(points).to(dev)
        ones = torch.ones(n, dtype=torch.float64, device=dev)
        zeros = torch.zeros(n, dtype=torch.float64, device=dev)
        ix = torch.arange(n, dtype=torch.int64, device=dev)
        pos = torch.empty(n, dtype=torch.int32, device=dev)
        indptr = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        M = torch.empty(n_pad, n_pad, dtype=torch.float64, device=dev)
        # weights = 1, qdw = 0, empty sparse pattern: M[r, c] = q_rc (r != c), M[r, r] = 0
        _lib.check(L.scb_system_assemble(n, _lib.ptr(sites), _lib.ptr(ones), _lib.ptr(zeros), None, _lib.ptr(zeros),
                                         _lib.ptr(indptr), _lib.ptr(indptr), _lib.ptr(zeros), None, n, _lib.ptr(ix),
                                         _lib.ptr(pos), n_pad, _lib.ptr(M), None, None, _lib.stream_ptr()))
        return M[:n, :n].cpu().numpy()


def cdist(XA: np.ndarray, XB: np.ndarray, metric: str = "euclidean") -> np.ndarray:
    """Pointwise distance between observations in 2D or 3D (reference distance.py:57-84)."""
    import torch

    from . import _lib

    metrics = ("euclidean", "sqeuclidean")
    if metric not in metrics:
        raise ValueError(f"Metric must be one of {metrics!r}, got {metric!r}.")
    XA = np.ascontiguousarray(XA, dtype=np.float64)
    XB = np.ascontiguousarray(XB, dtype=np.float64)
    if XA.shape[1] != XB.shape[1]:
        raise ValueError(f"XA.shape[1] ({XA.shape[1]}) must be equal to XB.shape[1] ({XB.shape[1]}).")
    if XA.shape[1] not in (2, 3):
        raise ValueError(f"Excpected shape (n, 2) arrays, got {XA.shape} and {XB.shape}.")
    L = _lib.lib()
    dev = torch.device(f"cuda:{torch.cuda.current_device()}")
    with torch.cuda.device(dev):
        a, b = torch.as_tensor(XA).to(dev), torch.as_tensor(XB).to(dev)
        out = torch.empty(len(XA), len(XB), dtype=torch.float64, device=dev)
        _lib.check(L.scb_cdist(XA.shape[1], 1 if metric == "sqeuclidean" else 0, len(XA), _lib.ptr(a), len(XB),
                               _lib.ptr(b), _lib.ptr(out), _lib.stream_ptr()))
        return out.cpu().numpy()
