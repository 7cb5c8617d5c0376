"""Fluxoid helpers (reference superscreen/fluxoid.py)."""
from __future__ import annotations

from typing import Dict, List, Optional, Union

import numpy as np

from .device import Device
from .geometry import orient_ccw
from .solution import Solution
from .solver import FactorizedModel, solve


def make_fluxoid_polygons(device: Device, holes: Optional[Union[List[str], str]] = None,
                          interp_points: Optional[int] = None, **_) -> Dict[str, np.ndarray]:
    """reference fluxoid.py:12-52.  The reference buffers each hole with shapely by half the
    distance to the nearest other polygon; here the hole polygon is scaled about its centroid so
    that its mean radius grows by half the (vertex-sampled) distance to the nearest other
    polygon in the layer.  Identical for circles; pass explicit polygons for parity runs."""
    polys = {**device.films, **device.holes}
    if holes is None:
        holes = list(device.holes)
    if isinstance(holes, str):
        holes = [holes]
    out = {}
    for name in holes:
        hole = device.holes[name]
        pts = hole.points[:-1]
        dmin = np.inf
        for other in polys.values():
            if other.layer != hole.layer or other.name == name:
                continue
            d = np.linalg.norm(pts[:, None, :] - other.points[None, :-1, :], axis=2).min()
            dmin = min(dmin, d)
        c = pts.mean(axis=0)
        r = np.linalg.norm(pts - c, axis=1).mean()
        new = c + (pts - c) * (1.0 + 0.5 * dmin / r)
        if interp_points:
            t = np.linspace(0, len(new), interp_points, endpoint=False)
            closed = np.concatenate([new, new[:1]])
            new = np.stack([np.interp(t, np.arange(len(closed)), closed[:, k]) for k in range(2)], axis=1)
        out[name] = orient_ccw(new)
    return out


def find_fluxoid_solution(model: FactorizedModel, fluxoids: Optional[Dict[str, float]] = None,
                          hole_polygon_mapping: Optional[Dict[str, np.ndarray]] = None, **solve_kwargs) -> Solution:
    """reference fluxoid.py:55-119"""
    device = model.device
    fluxoids = fluxoids or {}
    hole_names = list(device.holes)
    current_units = model.current_units
    solve_kwargs = solve_kwargs.copy()
    applied_field = solve_kwargs.pop("applied_field", None)
    target = np.array([fluxoids.get(name, 0) for name in hole_names])
    if hole_polygon_mapping is None:
        hole_polygon_mapping = make_fluxoid_polygons(device)
    orig = model.circulating_currents
    try:
        model.set_circulating_currents({name: 0 for name in hole_names})
        solution_no_circ = solve(model=model, applied_field=applied_field, **solve_kwargs)[-1]
        if not hole_names:
            if np.any(target):
                raise ValueError("Cannot calculate nonzero fluxoid solution for a device with no holes.")
            return solution_no_circ
        current = np.array([
            sum(solution_no_circ.hole_fluxoid(name, points=hole_polygon_mapping[name], with_units=False))
            for name in hole_names
        ])
        M = device.mutual_inductance_matrix(hole_polygon_mapping, units=f"Phi_0 / ({current_units})", **solve_kwargs)
        I_circ = np.linalg.solve(M, target - current)
        model.set_circulating_currents(dict(zip(hole_names, I_circ)))
        solution = solve(model=model, applied_field=applied_field, **solve_kwargs)[-1]
    finally:
        model.set_circulating_currents(orig)
    return solution
