"""solve / factorize_model / FactorizedModel (reference solver/solve.py) on the B200.

Same signatures, return types and error behaviour as the reference; the arithmetic runs in
libsc_b200 (include/scb.h).  All per-film state stays on the device across the film-to-film
iterations (reference solve.py:491-547); only the O(n) result vectors of each ``Solution`` are
copied back to the host.
"""
from __future__ import annotations

import copy
import itertools
import logging
import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Union

import numpy as np

from .. import _lib
from .. import units as _u
from ..device import Device
from ..solution import FilmSolution, Solution, Vortex
from ..sources import ConstantField
from .solve_film import LinearSystem, factorize_linear_systems, solve_film_device
from .utils import FilmInfo, currents_to_floats, field_conversion_factor, make_film_info

logger = logging.getLogger("solve")


def _torch():
    import torch

    return torch


def biot_savart_film_to_film(*, film1_sites, film1_z0: float, film1_areas, film1_J, film2_sites,
                             film2_z0: float) -> np.ndarray:
    """reference solver/solve.py:28-73 (host arrays in/out; the pair sum runs on the device)."""
    torch = _torch()
    _lib.lib()
    dev = torch.device(f"cuda:{torch.cuda.current_device()}")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    out = film_to_film_device(t(film1_sites), float(film1_z0), t(film1_areas), t(film1_J), t(film2_sites),
                              float(film2_z0))
    return out.cpu().numpy()


def film_to_film_device(src_sites, src_z0: float, src_areas, src_J, tgt_sites, tgt_z0: float):
    torch = _torch()
    L = _lib.lib()
    m, n = int(tgt_sites.shape[0]), int(src_sites.shape[0])
    with torch.cuda.device(tgt_sites.device):
        out = torch.empty(m, dtype=torch.float64, device=tgt_sites.device)
        _lib.check(L.scb_biot_savart(0, m, _lib.ptr(tgt_sites), n, _lib.ptr(src_sites), _lib.ptr(src_areas),
                                     _lib.ptr(src_J.contiguous()), float(tgt_z0) - float(src_z0),
                                     1.0 / (4.0 * np.pi), 1, _lib.ptr(out), _lib.stream_ptr()))
    return out


@dataclass
class FactorizedModel:
    """reference solver/solve.py:76-220 (HDF5 persistence is out of scope)."""

    device: Device
    film_info: Dict[str, FilmInfo]
    film_systems: Dict[str, LinearSystem]
    hole_systems: Dict[str, Dict[str, LinearSystem]]
    terminal_systems: Dict[str, object]
    terminal_currents: Dict[str, Dict[str, float]]
    circulating_currents: Dict[str, float]
    vortices: Union[Sequence[Vortex], Dict[str, Sequence[Vortex]]]
    current_units: str

    def set_circulating_currents(self, circulating_currents: Dict[str, float]) -> None:
        diff = set(circulating_currents) - set(self.device.holes)
        if diff:
            raise KeyError(f"circulating_currents contains keys not in self.device.holes: {list(diff)!r}")
        self.circulating_currents = circulating_currents.copy()
        holes_by_film = self.device.holes_by_film()
        for film_name, film_info in self.film_info.items():
            holes = [hole.name for hole in holes_by_film[film_name]]
            film_info.circulating_currents = {}
            for hole, current in self.circulating_currents.items():
                if hole in holes:
                    film_info.circulating_currents[hole] = current

    def set_vortices(self, vortices: Sequence[Vortex]) -> None:
        for film in self.film_info.values():
            film.vortices = []
        for vortex in vortices:
            self.film_info[vortex.film].vortices.append(vortex)
        self.vortices = {}
        for name, film in self.film_info.items():
            film.vortices = tuple(film.vortices)
            self.vortices[name] = film.vortices

    def copy(self) -> "FactorizedModel":
        return copy.copy(self)


def factorize_model(*, device: Device, current_units: str, terminal_currents=None, circulating_currents=None,
                    vortices: Optional[Sequence[Vortex]] = None) -> FactorizedModel:
    """reference solver/solve.py:223-287"""
    ureg = device.ureg
    circulating_currents = currents_to_floats(circulating_currents or {}, ureg, current_units)
    terminal_currents = {
        film_name: currents_to_floats(currents, ureg, current_units)
        for film_name, currents in (terminal_currents or {}).items()
    }
    for film_name, currents in terminal_currents.items():
        if sum(currents.values()):
            raise ValueError(f"Terminal currents in film {film_name!r} are not conserved.")
    vortices = vortices or []
    if not device.meshes:
        raise ValueError("The device does not have a mesh. Call device.make_mesh() to generate it.")
    film_info = make_film_info(device=device, vortices=vortices, circulating_currents=circulating_currents,
                               terminal_currents=terminal_currents)
    film_systems, hole_systems, terminal_systems = factorize_linear_systems(device, film_info)
    return FactorizedModel(device, film_info, film_systems, hole_systems, terminal_systems, terminal_currents,
                           circulating_currents, vortices, current_units)


def _to_solution(device, film_info, results, applied_fields, others, field_conversion, solution_kwargs) -> Solution:
    film_solutions = {}
    for name, (g, J, self_field) in results.items():
        other = None if others is None else (others[name] / field_conversion).cpu().numpy()
        film_solutions[name] = FilmSolution(
            stream=g.cpu().numpy(),
            current_density=J.cpu().numpy(),
            applied_field=(applied_fields[name] / field_conversion).cpu().numpy(),
            self_field=(self_field / field_conversion).cpu().numpy(),
            field_from_other_films=other,
        )
    return Solution(device=device, film_solutions=film_solutions, **solution_kwargs)


def solve(device: Optional[Device] = None, *, model: Optional[FactorizedModel] = None,
          applied_field: Optional[Callable] = None, terminal_currents=None, circulating_currents=None,
          vortices: Optional[Sequence[Vortex]] = None, field_units: str = "mT", current_units: str = "uA",
          check_inversion: bool = False, iterations: int = 0, return_solutions: bool = True,
          save_path: Optional[os.PathLike] = None, log_level: Optional[int] = None, progress_bar: bool = True,
          _solver: str = "superscreen_b200.solve") -> List[Solution]:
    """reference solver/solve.py:290-549"""
    torch = _torch()
    if log_level is not None:
        logging.basicConfig(level=log_level)
    if model is None:
        if device is None:
            raise ValueError("Either a model or a device must be provided.")
        logger.info("Factorizing model.")
        model = factorize_model(device=device, current_units=current_units, terminal_currents=terminal_currents,
                                circulating_currents=circulating_currents, vortices=vortices)
    else:
        if (device is not None or terminal_currents is not None or circulating_currents is not None
                or vortices is not None):
            raise ValueError(
                "If model argument is provided, device, terminal_currents,"
                " circulating_currents, and vortices must be None."
            )
    if not isinstance(model, FactorizedModel):
        raise TypeError(f"model must be an instance of FactorizedModel (got {type(model)}).")
    if save_path is not None:
        raise NotImplementedError("HDF5 persistence is a 'next' row of the hot-path scope (SURVEY.md 8f.3).")

    device = model.device
    film_info = model.film_info
    film_systems = model.film_systems
    hole_systems = model.hole_systems
    circulating_currents = model.circulating_currents
    terminal_currents = model.terminal_currents
    vortices = model.vortices
    current_units = model.current_units
    if not device.meshes:
        raise ValueError("The device does not have a mesh. Call device.make_mesh() to generate it.")

    length_units = device.length_units
    meshes = device.meshes
    applied_field = applied_field or ConstantField(0)
    field_conversion = field_conversion_factor(field_units, current_units, length_units=length_units).magnitude

    applied_fields = {}
    for film, mesh in meshes.items():
        layer = device.layers[film_info[film].layer]
        z0 = layer.z0 * np.ones(len(mesh.sites))
        Hz_applied = np.squeeze(applied_field(mesh.sites[:, 0], mesh.sites[:, 1], z0) * field_conversion)
        Hz_applied = np.asarray(Hz_applied, dtype=np.float64)
        if Hz_applied.ndim != 1:
            raise ValueError(
                "Expected applied_field to return a 1D vector,"
                f" got a {Hz_applied.shape[1]}D vector."
            )
        applied_fields[film] = torch.as_tensor(np.ascontiguousarray(Hz_applied)).to(mesh._data.device)

    # Phi_0 / mu_0 in [current_units * length_units]  (reference solve.py:441)
    vortex_flux = _u.PHI_0 / _u.MU_0 * _u.conversion_factor("A * m", f"({current_units}) * ({length_units})")

    solution_kwargs = dict(applied_field_func=applied_field, field_units=field_units, current_units=current_units,
                           circulating_currents=circulating_currents, terminal_currents=terminal_currents,
                           vortices=vortices, solver=_solver)

    def run(others):
        return {
            name: solve_film_device(
                film_info=film_info[name], film_system=film_systems[name], hole_systems=hole_systems[name],
                applied_field=applied_fields[name], vortex_flux=vortex_flux,
                field_from_other_films=None if others is None else others[name], check_inversion=check_inversion)
            for name in device.films
        }

    solutions: List[Solution] = []
    results = run(None)
    if return_solutions:
        solutions.append(_to_solution(device, film_info, results, applied_fields, None, field_conversion,
                                      solution_kwargs))
    if len(device.films) < 2 or iterations < 1:
        return solutions if return_solutions else None

    z0s = {name: float(device.layers[film_info[name].layer].z0) for name in device.films}
    for i in range(iterations):
        # Jacobi step: all film-to-film fields from the previous iterate, then all re-solves
        others = {name: torch.zeros_like(applied_fields[name]) for name in device.films}
        for source_film, film in itertools.product(device.films, repeat=2):
            if film == source_film:
                continue
            src, dst = meshes[source_film]._data, meshes[film]._data
            J = results[source_film][1]
            if J.device != dst.device:  # multi-GPU placement: bring the source film over NVLink
                J = J.to(dst.device)
            others[film] += film_to_film_device(
                src.sites.to(dst.device), z0s[source_film], src.t["vertex_areas"].to(dst.device), J,
                dst.sites, z0s[film])
        results = run(others)
        if return_solutions:
            solutions.append(_to_solution(device, film_info, results, applied_fields, others, field_conversion,
                                          solution_kwargs))
    return solutions if return_solutions else None
