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
from .solve_film import (  # noqa: F401  (TerminalSystems, solve_film: names the reference module exposes)
    LinearSystem, TerminalSystems, factorize_linear_systems, solve_film, solve_film_device)
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
    """Field of the source film's sheet current at the target sites, solver units.  ``src_J`` is
    ``(n, 2)`` -> ``(m,)`` or a batch ``(B, n, 2)`` -> ``(B, m)``."""
    torch = _torch()
    L = _lib.lib()
    m, n = int(tgt_sites.shape[0]), int(src_sites.shape[0])
    nsets = int(src_J.shape[0]) if src_J.dim() == 3 else 1
    with torch.cuda.device(tgt_sites.device):
        J = src_J.contiguous()
        out = torch.empty((nsets, m) if src_J.dim() == 3 else (m,), dtype=torch.float64, device=tgt_sites.device)
        _lib.check(L.scb_biot_savart(0, m, _lib.ptr(tgt_sites), n, _lib.ptr(src_sites), _lib.ptr(src_areas),
                                     _lib.ptr(J), float(tgt_z0) - float(src_z0),
                                     1.0 / (4.0 * np.pi), nsets, _lib.ptr(out), _lib.stream_ptr()))
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
    comm: object = None  # parallel.Comm: film -> owner rank (None = single process)

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
                    vortices: Optional[Sequence[Vortex]] = None, comm=None) -> FactorizedModel:
    """reference solver/solve.py:223-287.  With a multi-rank ``comm`` (``parallel.DistComm``) each
    rank assembles and factorizes only the films it owns (one film factorization per GPU)."""
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
    from ..parallel import Comm, film_owners

    comm = comm or Comm()
    owners = film_owners(list(device.films), comm)
    owned = {f for f, r in owners.items() if r == comm.rank}
    film_systems, hole_systems, terminal_systems = factorize_linear_systems(device, film_info, owned=owned)
    return FactorizedModel(device, film_info, film_systems, hole_systems, terminal_systems, terminal_currents,
                           circulating_currents, vortices, current_units, comm)


_PINNED = {}  # device index -> pinned staging buffer (float64), grown on demand


def _to_host(t) -> np.ndarray:
    """Device tensor -> host numpy array.  Large arrays go through a cached pinned staging buffer
    (a pageable cudaMemcpy runs at a fraction of the PCIe/C2C rate)."""
    import torch

    t = t.contiguous()
    if t.dtype != torch.float64 or t.numel() < (1 << 17):
        return t.cpu().numpy()
    key = t.device.index
    buf = _PINNED.get(key)
    if buf is None or buf.numel() < t.numel():
        buf = torch.empty(max(t.numel(), 1 << 22), dtype=torch.float64, pin_memory=True)
        _PINNED[key] = buf
    stage = buf[: t.numel()].view(t.shape)
    stage.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return stage.numpy().copy()


def _to_solutions(device, film_names, results, applied_fields, others, field_conversion, kwargs_list):
    """Device tensors (solver units) -> host FilmSolutions (reference solve_film.py:566-573).  One
    device->host copy per array per film, shared by all batch entries; ``kwargs_list`` holds the
    Solution keyword arguments of every batch entry (length 1 when not batched)."""
    batched = next(iter(results.values()))[0].dim() == 2
    host = {}
    for name in film_names:
        g, J, self_field = results[name]
        applied = applied_fields[name] / field_conversion
        self_field = self_field / field_conversion
        other = None if others is None else others[name] / field_conversion
        if batched:
            # batch-major on the device, so that every batch entry is a contiguous host row (no per-entry copies)
            g, applied, self_field = g.t(), applied.t(), self_field.t()
            other = None if other is None else other.t()
        host[name] = (_to_host(g), _to_host(J), _to_host(applied), _to_host(self_field),
                      None if other is None else _to_host(other))
    out = []
    for b, kwargs in enumerate(kwargs_list):
        film_solutions = {}
        for name in film_names:
            g, J, applied, self_field, other = host[name]
            if batched:
                g, J, applied, self_field = g[b], J[b], applied[b], self_field[b]
                other = None if other is None else other[b]
            film_solutions[name] = FilmSolution(
                stream=np.ascontiguousarray(g), current_density=np.ascontiguousarray(J),
                applied_field=np.ascontiguousarray(applied), self_field=np.ascontiguousarray(self_field),
                field_from_other_films=None if other is None else np.ascontiguousarray(other))
        out.append(Solution(device=device, film_solutions=film_solutions, **kwargs))
    return out


def _evaluate_applied_field(applied_field, device, film_info, meshes, field_conversion):
    """reference solver/solve.py:422-436 -> {film: host float64 (n,) in solver units}"""
    out = {}
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
        out[film] = np.ascontiguousarray(Hz_applied)
    return out


def _run(model: "FactorizedModel", applied_fields, circ_by_film, vortex_flux, iterations, check_inversion):
    """Runs the solve + film-to-film iterations on the device for single or batched right-hand
    sides; returns the per-iteration results replicated on every rank."""
    torch = _torch()
    from ..parallel import Comm, film_owners, gather_film_results, run_film_iterations

    device = model.device
    comm = model.comm or Comm()
    film_names = list(device.films)
    owners = film_owners(film_names, comm)
    meshes = device.meshes
    film_info = model.film_info
    z0s = {name: float(device.layers[film_info[name].layer].z0) for name in film_names}
    some = next(iter(applied_fields.values()))
    batch = some.shape[1] if some.dim() == 2 else None

    def solve_fn(name, other):
        return solve_film_device(
            film_info=film_info[name], film_system=model.film_systems[name],
            hole_systems=model.hole_systems[name], applied_field=applied_fields[name], vortex_flux=vortex_flux,
            field_from_other_films=other, check_inversion=check_inversion,
            circulating_currents=None if circ_by_film is None else circ_by_film[name],
            terminal_systems=model.terminal_systems.get(name), device=device)

    def coupling_fn(src_name, J_src, dst_name):
        src, dst = meshes[src_name]._data, meshes[dst_name]._data
        out = film_to_film_device(src.sites.to(dst.device), z0s[src_name], src.t["vertex_areas"].to(dst.device),
                                  J_src.to(dst.device), dst.sites, z0s[dst_name])
        return out.t().contiguous() if batch is not None else out

    def zeros_fn(name):
        return torch.zeros_like(applied_fields[name])

    def j_shape(name):
        n = len(meshes[name].sites)
        return (batch, n, 2) if batch is not None else (n, 2)

    def v_shape(name):
        n = len(meshes[name].sites)
        return (n, batch) if batch is not None else (n,)

    # the films of one Jacobi step are independent and each is latency-bound (getrs sweeps, small
    # N-body launches): one side stream per owned film, joined after every step
    import contextlib

    mine = [f for f in film_names if owners[f] == comm.rank]
    side = {}
    if len(mine) > 1 and int(os.environ.get("SCB_FILM_STREAMS", "8")) > 0:
        pool = [torch.cuda.Stream(device=some.device) for _ in range(min(len(mine), int(os.environ.get("SCB_FILM_STREAMS", "8"))))]
        side = {f: pool[k % len(pool)] for k, f in enumerate(mine)}

    @contextlib.contextmanager
    def film_scope(name):
        st = side.get(name)
        if st is None:
            yield
            return
        st.wait_stream(torch.cuda.current_stream(some.device))
        with torch.cuda.stream(st):
            yield

    def join():
        main = torch.cuda.current_stream(some.device)
        for st in set(side.values()):
            main.wait_stream(st)

    per_iter = run_film_iterations(film_names, owners, comm, solve_fn, coupling_fn, zeros_fn, j_shape, iterations,
                                   film_scope=film_scope, join=join)
    return gather_film_results(per_iter, film_names, owners, comm,
                               {"g": v_shape, "J": j_shape, "self": v_shape, "other": v_shape}, some)


def _check_model_args(device, model, terminal_currents, circulating_currents, vortices, current_units):
    """Argument validation shared by solve / solve_batch (reference solver/solve.py:357-389)."""
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
    if not model.device.meshes:
        raise ValueError("The device does not have a mesh. Call device.make_mesh() to generate it.")
    return model


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
    model = _check_model_args(device, model, terminal_currents, circulating_currents, vortices, current_units)
    if save_path is not None:
        raise NotImplementedError("HDF5 persistence is a 'next' row of the hot-path scope (SURVEY.md 8f.3).")
    device = model.device
    current_units = model.current_units
    length_units = device.length_units
    applied_field = applied_field or ConstantField(0)
    field_conversion = field_conversion_factor(field_units, current_units, length_units=length_units).magnitude
    host_fields = _evaluate_applied_field(applied_field, device, model.film_info, device.meshes, field_conversion)
    applied_fields = {f: torch.as_tensor(h).to(device.meshes[f]._data.device) for f, h in host_fields.items()}
    # Phi_0 / mu_0 in [current_units * length_units]  (reference solve.py:441)
    vortex_flux = _u.PHI_0 / _u.MU_0 * _u.conversion_factor("A * m", f"({current_units}) * ({length_units})")
    solution_kwargs = dict(applied_field_func=applied_field, field_units=field_units, current_units=current_units,
                           circulating_currents=model.circulating_currents,
                           terminal_currents=model.terminal_currents, vortices=model.vortices, solver=_solver)
    per_iter = _run(model, applied_fields, None, vortex_flux, iterations, check_inversion)
    if not return_solutions:
        return None
    film_names = list(device.films)
    return [_to_solutions(device, film_names, results, applied_fields, others, field_conversion,
                          [solution_kwargs])[0] for results, others in per_iter]


def solve_batch(*, model: FactorizedModel, applied_fields: Sequence[Optional[Callable]],
                circulating_currents: Optional[Sequence[Dict[str, float]]] = None, field_units: str = "mT",
                iterations: int = 0, check_inversion: bool = False, last_only: bool = False,
                _solver: str = "superscreen_b200.solve_batch") -> List[List[Solution]]:
    """Solves B models that share one factorization in a single batched pass (multi-RHS getrs,
    multi-RHS matrix-free operator, one film-to-film exchange per iteration for the whole batch).

    This is an extension with no counterpart in the reference, which would call ``solve`` B times
    (e.g. once per driven hole in ``Device.mutual_inductance_matrix``, device/device.py:610-639, or
    once per applied field in a sweep).  ``out[b]`` equals
    ``solve(model=model_b, applied_field=applied_fields[b], iterations=iterations)`` where
    ``model_b`` has ``circulating_currents[b]`` (floats in ``model.current_units``).  With
    ``last_only`` only the final iterate is brought to the host (``out[b]`` has one element).
    """
    torch = _torch()
    model = _check_model_args(None, model, None, None, None, None)
    device = model.device
    B = len(applied_fields)
    if B == 0:
        return []
    if circulating_currents is None:
        circulating_currents = [model.circulating_currents] * B
    if len(circulating_currents) != B:
        raise ValueError("applied_fields and circulating_currents must have the same length.")
    for cc in circulating_currents:
        diff = set(cc) - set(device.holes)
        if diff:
            raise KeyError(f"circulating_currents contains keys not in self.device.holes: {list(diff)!r}")
    current_units = model.current_units
    length_units = device.length_units
    field_conversion = field_conversion_factor(field_units, current_units, length_units=length_units).magnitude
    funcs = [f or ConstantField(0) for f in applied_fields]
    per_b = [_evaluate_applied_field(f, device, model.film_info, device.meshes, field_conversion) for f in funcs]
    film_names = list(device.films)
    dev_fields = {}
    circ_by_film = {}
    for name in film_names:
        dev = device.meshes[name]._data.device
        # rows are contiguous on the host (fast stack + one H2D copy); the (n, nrhs) layout is made on the device
        dev_fields[name] = torch.as_tensor(np.stack([h[name] for h in per_b], axis=0)).to(dev).t().contiguous()
        circ_by_film[name] = {
            hole: torch.tensor([float(cc.get(hole, 0.0)) for cc in circulating_currents], dtype=torch.float64,
                               device=dev)
            for hole in model.film_info[name].hole_indices
        }
    vortex_flux = _u.PHI_0 / _u.MU_0 * _u.conversion_factor("A * m", f"({current_units}) * ({length_units})")
    per_iter = _run(model, dev_fields, circ_by_film, vortex_flux, iterations, check_inversion)
    kwargs_list = [dict(applied_field_func=funcs[b], field_units=field_units, current_units=current_units,
                        circulating_currents=dict(circulating_currents[b]),
                        terminal_currents=model.terminal_currents, vortices=model.vortices, solver=_solver)
                   for b in range(B)]
    if last_only:
        per_iter = per_iter[-1:]
    per_iter_solutions = [_to_solutions(device, film_names, results, dev_fields, others, field_conversion,
                                        kwargs_list) for results, others in per_iter]
    return [[it[b] for it in per_iter_solutions] for b in range(B)]
