"""FilmInfo / LambdaInfo / unit helpers (reference solver/utils.py).

Index sets are computed on the host with the package's point-in-polygon test and handed to the
C ABI as inputs.  The dense ``kernel`` / ``laplacian`` / ``gradient`` arrays the
reference stores eagerly (solver/utils.py:290-297) are lazy properties here: the solve path
never densifies them.
"""
from __future__ import annotations

import logging
import numbers
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from .. import units as _u
from ..device import Device, Polygon
from ..geometry import path_vectors  # noqa: F401  (names the reference module exposes)
from ..mesh import Mesh
from ..solution import Vortex  # noqa: F401

logger = logging.getLogger("solve")


class LambdaInfo:
    """reference solver/utils.py:19-58"""

    lambda_str = "λ"
    Lambda_str = "Λ"

    def __init__(self, *, film: str, Lambda: np.ndarray, london_lambda: Optional[np.ndarray] = None,
                 thickness: Optional[float] = None, _constants: Optional[Tuple[float, Optional[float]]] = None):
        self.film = film
        self.Lambda = Lambda
        self.london_lambda = london_lambda
        self.thickness = thickness
        if _constants is not None:
            # Lambda (and london_lambda) are filled arrays of these scalars: the checks below on two numbers
            # instead of several passes over (n, 1) arrays
            lam, lon = _constants
            self.inhomogeneous = False
            if lon is not None:
                assert thickness is not None
                assert np.allclose(lam, lon**2 / thickness)
            if lam < 0:
                raise ValueError(f"Negative Lambda in film {film!r}.")
            return
        self.inhomogeneous = bool(
            np.ptp(self.Lambda) / max(np.min(np.abs(self.Lambda)), np.finfo(float).eps) > 1e-6
        )
        if self.inhomogeneous:
            logger.info(
                f"Inhomogeneous {LambdaInfo.Lambda_str} in film {self.film!r}, "
                f"which violates the assumptions of the London model. "
                f"Results may not be reliable."
            )
        if self.london_lambda is not None:
            assert self.thickness is not None
            assert np.allclose(self.Lambda, self.london_lambda**2 / self.thickness)
        if np.any(self.Lambda < 0):
            raise ValueError(f"Negative Lambda in film {film!r}.")


@dataclass
class FilmInfo:
    """reference solver/utils.py:96-132; dense members are lazy."""

    name: str
    layer: str
    lambda_info: LambdaInfo
    vortices: Tuple
    interior_indices: np.ndarray
    boundary_indices: np.ndarray
    hole_indices: Dict[str, np.ndarray]
    in_hole: np.ndarray
    circulating_currents: Dict[str, float]
    mesh: Mesh = field(repr=False, default=None)
    terminal_currents: Optional[Dict[str, float]] = None
    # device-resident state (torch tensors), filled by make_film_info
    dev: Dict[str, object] = field(repr=False, default_factory=dict)

    def to_hdf5(self, h5group) -> None:
        """reference solver/utils.py:134-164"""
        from .. import io as _io

        _io.film_info_to_hdf5(self, h5group)

    @property
    def weights(self) -> np.ndarray:
        return self.mesh.operators.weights

    @property
    def kernel(self) -> np.ndarray:
        """Dense Q (n, n); materialised on demand only."""
        return self.mesh.operators.Q

    @property
    def laplacian(self) -> np.ndarray:
        return self.mesh.operators.laplacian.toarray()

    @property
    def gradient(self) -> Optional[np.ndarray]:
        if not self.lambda_info.inhomogeneous:
            return None
        ops = self.mesh.operators
        return np.array([ops.gradient_x.toarray(), ops.gradient_y.toarray()])


def get_holes_and_vortices_by_film(device: Device, vortices: Sequence):
    """reference solver/utils.py:214-231"""
    from ..solution import Vortex

    vortices_by_film = {film_name: [] for film_name in device.films}
    holes_by_film = device.holes_by_film()
    for vortex in vortices:
        if not isinstance(vortex, Vortex):
            raise TypeError(f"Expected a Vortex, but got {type(vortex)}.")
        if not device.films[vortex.film].contains_points((vortex.x, vortex.y)).all():
            raise ValueError(f"Vortex {vortex!r} is not located in film {vortex.film!r}.")
        for hole in holes_by_film[vortex.film]:
            if hole.contains_points((vortex.x, vortex.y)).all():
                raise ValueError(f"Vortex {vortex} is located in hole {hole.name!r}.")
        vortices_by_film[vortex.film].append(vortex)
    return holes_by_film, vortices_by_film


def _evaluate(param, x, y):
    if isinstance(param, numbers.Real):
        return float(param) * np.ones_like(x)
    return np.asarray(param(x, y), dtype=np.float64) * np.ones_like(x)


def make_film_info(*, device: Device, vortices: Sequence, circulating_currents: Dict[str, float],
                   terminal_currents: Dict[str, Dict[str, float]]) -> Dict[str, FilmInfo]:
    """reference solver/utils.py:234-324"""
    import torch

    holes_by_film, vortices_by_film = get_holes_and_vortices_by_film(device, vortices)
    film_info = {}
    for name, film in device.films.items():
        mesh = device.meshes[name]
        layer = device.layers[film.layer]
        london_lambda = layer.london_lambda
        d = layer.thickness
        x, y = mesh.sites[:, 0], mesh.sites[:, 1]
        layer_Lambda = layer.Lambda
        constants = None
        if isinstance(layer_Lambda, numbers.Real) and (london_lambda is None or isinstance(london_lambda, numbers.Real)):
            constants = (float(layer_Lambda), None if london_lambda is None else float(london_lambda))
            Lambda = np.full((len(x), 1), constants[0], dtype=np.float64)
        else:
            Lambda = _evaluate(layer_Lambda, x, y).astype(np.float64)[:, np.newaxis]
        if london_lambda is not None:
            if isinstance(london_lambda, numbers.Real) and london_lambda <= d:
                logger.info(
                    f"Layer {name!r}: The film thickness, d = {d:.4f}, is greater than or equal to the "
                    f"London penetration depth; the thin-film assumption may not be valid."
                )
            if constants is not None:
                london_lambda = np.full((len(x), 1), constants[1], dtype=np.float64)
            else:
                london_lambda = _evaluate(london_lambda, x, y)[:, np.newaxis]
        # index sets (reference solver/utils.py:271-304) depend only on the mesh and the polygons:
        # memoised per device, so that repeated factorizations (Lambda sweeps, one model per
        # mutual-inductance call) do not redo the point-in-polygon tests
        geo_key = (id(mesh), film.points.tobytes(), tuple(h.points.tobytes() for h in holes_by_film[name]),
                   name in device.terminals)
        cache = device.__dict__.setdefault("_film_geometry_cache", {})
        hit = cache.get(name)
        if hit is not None and hit[0] == geo_key and hit[1] is mesh:
            hole_indices, in_hole, boundary_indices, interior_indices = hit[2]
            derived = hit[3]
        else:
            hole_indices = {
                hole.name: hole.contains_points(mesh.sites, index=True).astype(np.int64)
                for hole in holes_by_film[name]
            }
            in_hole = np.zeros(len(mesh.sites), dtype=bool)
            if hole_indices:
                in_hole[np.concatenate(list(hole_indices.values()))] = True
            if name in device.terminals:
                boundary_indices = device.boundary_vertices(name)  # ordered counter-clockwise
            else:
                boundary_indices = mesh.boundary_indices
            keep = film.contains_points(mesh.sites).copy()
            keep[boundary_indices] = False
            interior_indices = np.where(keep)[0].astype(np.int64)  # == setdiff1d(in film, boundary), ascending
            derived = {}  # host / device arrays derived from the index sets (factorize_linear_systems)
            cache[name] = (geo_key, mesh, (hole_indices, in_hole, boundary_indices, interior_indices), derived)
        circ = {h: c for h, c in circulating_currents.items() if h in hole_indices}
        lambda_info = LambdaInfo(film=name, Lambda=Lambda, london_lambda=london_lambda, thickness=layer.thickness,
                                 _constants=constants)
        info = FilmInfo(
            name=name, layer=layer.name, lambda_info=lambda_info, vortices=tuple(vortices_by_film[name]),
            interior_indices=interior_indices, boundary_indices=boundary_indices, hole_indices=hole_indices,
            in_hole=in_hole, circulating_currents=circ, mesh=mesh,
            terminal_currents=terminal_currents.get(name),
        )
        dev = mesh._data.device
        lam0 = float(Lambda[0, 0]) if len(Lambda) else 0.0
        if len(Lambda) and (constants is not None or float(Lambda.min()) == lam0 == float(Lambda.max())):
            # constant Lambda: filled on the device (no pageable host-to-device copy per factorization)
            info.dev["Lambda"] = torch.full((len(Lambda),), lam0, dtype=torch.float64, device=dev)
            info.dev["Lambda_constant"] = True
        else:
            info.dev["Lambda"] = torch.as_tensor(np.ascontiguousarray(Lambda[:, 0])).to(dev)
            info.dev["Lambda_constant"] = False
        info.dev["_geo"] = derived
        film_info[name] = info
    return film_info


def current_to_float(value, ureg, current_units: str) -> float:
    """reference solver/utils.py:327-335"""
    if isinstance(value, (str, _u.Quantity)):
        return float(_u.to_quantity(value, current_units).to(current_units).magnitude)
    return value


def currents_to_floats(currents: Dict[str, Union[float, str]], ureg, current_units: str) -> Dict[str, float]:
    return {k: current_to_float(v, ureg, current_units) for k, v in currents.items()}


def convert_field(value, new_units: str, old_units: Optional[str] = None, ureg=None, with_units: bool = True):
    """reference solver/utils.py:350-404: converts between H ([current]/[length]) and B = mu0 H."""
    if isinstance(value, str):
        value = _u.to_quantity(value, old_units or "dimensionless")
    if isinstance(value, _u.Quantity):
        old_units = value.units
        value = value.magnitude
    if old_units is None:
        raise ValueError("Old units must be specified if value is not a string or pint.Quantity.")
    so, do = _u.parse(old_units)
    sn, dn = _u.parse(new_units)
    if np.allclose(do, dn):
        factor = so / sn
    elif do[0] != 0:  # old is H (has [length]); want B = mu0 * H
        s_mu, d_mu = _u.parse("mu_0")
        if not np.allclose(do + d_mu, dn):
            raise _u.DimensionalityError(f"Cannot convert {old_units!r} to {new_units!r}")
        factor = so * s_mu / sn
    else:
        s_mu, d_mu = _u.parse("mu_0")
        if not np.allclose(do - d_mu, dn):
            raise _u.DimensionalityError(f"Cannot convert {old_units!r} to {new_units!r}")
        factor = so / s_mu / sn
    out = value * factor
    if with_units:
        return _u.Quantity(out, new_units)
    return out


def field_conversion_factor(field_units: str, current_units: str, length_units: str = "m", ureg=None) -> _u.Quantity:
    """reference solver/utils.py:407-437"""
    target = f"({current_units}) / ({length_units})"
    s, d = _u.parse(field_units)
    st, dt = _u.parse(target)
    if np.allclose(d, dt):
        mag = s / st
    else:
        s_mu, d_mu = _u.parse("mu_0")
        if not np.allclose(d - d_mu, dt):
            raise _u.DimensionalityError(f"Cannot convert {field_units!r} to {target!r}")
        mag = s / s_mu / st
    return _u.Quantity(mag, f"({target}) / ({field_units})")


def stream_from_current_density(points: np.ndarray, J: np.ndarray) -> np.ndarray:
    """reference solver/utils.py:440-463"""
    from scipy import integrate

    zhat_cross_J = J[:, [1, 0]]
    zhat_cross_J[:, 0] *= -1
    dl = np.diff(points, axis=0)
    integrand = np.sum(zhat_cross_J * dl, axis=1)
    return integrate.cumulative_trapezoid(integrand, initial=0)


def stream_from_terminal_current(points: np.ndarray, current: float) -> np.ndarray:
    """reference solver/utils.py:466-488: the current is spread uniformly along the terminal."""
    from ..geometry import path_vectors

    edge_lengths, unit_normals = path_vectors(points)
    J = current * unit_normals / np.sum(edge_lengths)
    g = stream_from_current_density(points, J)
    return g * current / g[-1]
