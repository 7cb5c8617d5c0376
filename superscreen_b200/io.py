"""HDF5 wire layout of the solve path (reference: superscreen/io.py, solver/solve.py:102-180,
solver/solve_film.py:37-77,102-148, solver/utils.py:60-95,134-211, solution.py:132-164,936-1087,
device/device.py:936-1016, device/mesh.py:250-300, device/layer.py:108-138, device/polygon.py:621-634).

Everything here is written against the small *group protocol* that ``h5py.Group`` implements --
``create_group``, ``group[name] = array``, ``group[name]``, ``name in group``, ``group.items()``,
``group.attrs`` -- so the group / dataset / attribute names and dtypes are exactly the reference's.
``h5py`` itself is an optional dependency (it is not installed in the build image): file paths need it,
open groups do not, and :class:`MemoryGroup` is an in-memory implementation of the same protocol that
the tests use to pin the layout.

What is stored for a factorized model is the reference's content: per film the dense ``A``, ``indices``,
``lu`` / ``piv`` in ``scipy.linalg.lu_factor`` layout (identity ``piv`` for the unpivoted device
factorization, the real interchanges for ``scb_getrf_piv``), ``grad_Lambda_term``; per ``FilmInfo`` the
index sets, ``weights``, dense ``kernel`` and ``laplacian``.  Loading rebuilds the device-resident
operators and factors from the stored triangulations (a factorization is cheaper than an upload of the
factors would be to validate), so a file written by the reference loads as well.
"""
from __future__ import annotations

import contextlib
import datetime as dt
import os
import pickle
from typing import Any, Dict, List, Optional, Sequence

import numpy as np


# ----------------------------------------------------------------------------------------------
# group protocol
# ----------------------------------------------------------------------------------------------
class SoftLink:
    """Stand-in for ``h5py.SoftLink`` inside a :class:`MemoryGroup`."""

    def __init__(self, path: str):
        self.path = path


class MemoryGroup:
    """In-memory implementation of the subset of ``h5py.Group`` used by the wire layout."""

    def __init__(self, name: str = "/", parent: Optional["MemoryGroup"] = None):
        self.name = name
        self.attrs: Dict[str, Any] = {}
        self._items: Dict[str, Any] = {}
        self._parent = parent

    def _root(self) -> "MemoryGroup":
        g = self
        while g._parent is not None:
            g = g._parent
        return g

    def create_group(self, name: str) -> "MemoryGroup":
        if name in self._items:
            raise ValueError(f"Unable to create group (name already exists): {name}")
        child = MemoryGroup(self.name.rstrip("/") + "/" + name, self)
        self._items[name] = child
        return child

    def __setitem__(self, name: str, value) -> None:
        if name in self._items:
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        self._items[name] = value if isinstance(value, (SoftLink, MemoryGroup)) else np.array(value)

    def _resolve(self, value):
        if isinstance(value, SoftLink):
            return self._root()[value.path.lstrip("/")]
        return value

    def __getitem__(self, name: str):
        node = self
        for part in [p for p in str(name).split("/") if p]:
            node = node._resolve(node._items[part])
        return node

    def get(self, name: str, default=None):
        return self[name] if name in self else default

    def __contains__(self, name) -> bool:
        return str(name) in self._items

    def __iter__(self):
        return iter(self._items)

    def __len__(self) -> int:
        return len(self._items)

    def keys(self):
        return self._items.keys()

    def values(self):
        return [self._resolve(v) for v in self._items.values()]

    def items(self):
        return [(k, self._resolve(v)) for k, v in self._items.items()]


def _is_group(obj) -> bool:
    return hasattr(obj, "create_group") and hasattr(obj, "attrs")


@contextlib.contextmanager
def open_group(path_or_group, mode: str):
    """An open group, or an HDF5 file opened with ``mode`` (needs h5py)."""
    if _is_group(path_or_group):
        yield path_or_group
        return
    try:
        import h5py
    except ImportError as exc:  # pragma: no cover - depends on the environment
        raise ImportError(
            "Reading / writing HDF5 files needs the optional dependency h5py; pass an open group "
            "(h5py.Group or superscreen_b200.io.MemoryGroup) or install h5py."
        ) from exc
    with h5py.File(os.fspath(path_or_group), mode) as f:
        yield f


def _soft_link(group, name: str, path: str) -> None:
    if isinstance(group, MemoryGroup):
        group[name] = SoftLink(path)
        return
    import h5py

    group[name] = h5py.SoftLink(path)


def serialize_obj(group, obj: Any, name: str, attr: bool = False) -> None:
    """reference io.py:8-24 (dill when available, pickle otherwise)"""
    try:
        import dill as pickler
    except ImportError:  # pragma: no cover
        pickler = pickle
    if attr:
        try:
            if callable(obj):
                raise TypeError
            group.attrs[name] = obj
        except TypeError:
            group.attrs[f"{name}.pickle"] = np.void(pickler.dumps(obj))
    else:
        group[f"{name}.pickle"] = np.void(pickler.dumps(obj))


def deserialize_obj(group, name: str, attr: bool = False) -> Any:
    """reference io.py:27-44"""
    try:
        import dill as pickler
    except ImportError:  # pragma: no cover
        pickler = pickle
    if attr:
        if name in group.attrs:
            return group.attrs[name]
        if f"{name}.pickle" in group.attrs:
            return pickler.loads(np.void(group.attrs[f"{name}.pickle"]).tobytes())
    elif f"{name}.pickle" in group:
        return pickler.loads(np.void(np.asarray(group[f"{name}.pickle"])).tobytes())
    raise IOError(f"Unable to load {name}.")


def _attr_str(value) -> str:
    return value.decode() if isinstance(value, bytes) else str(value)


# ----------------------------------------------------------------------------------------------
# geometry: Layer / Polygon / Mesh / Device
# ----------------------------------------------------------------------------------------------
def layer_to_hdf5(layer, g) -> None:
    g.attrs["name"] = layer.name
    g.attrs["z0"] = layer.z0
    if layer.thickness is not None:
        g.attrs["thickness"] = layer.thickness
    if layer.london_lambda is not None:
        serialize_obj(g, layer.london_lambda, "london_lambda", attr=True)
    else:
        serialize_obj(g, layer.Lambda, "Lambda", attr=True)


def layer_from_hdf5(g):
    from .device import Layer

    Lambda = london_lambda = None
    thickness = g.attrs.get("thickness", None)
    if "london_lambda" in g.attrs or "london_lambda.pickle" in g.attrs:
        london_lambda = deserialize_obj(g, "london_lambda", attr=True)
    else:
        Lambda = deserialize_obj(g, "Lambda", attr=True)
    return Layer(_attr_str(g.attrs["name"]), Lambda=Lambda, london_lambda=london_lambda, thickness=thickness,
                 z0=g.attrs["z0"])


def polygon_to_hdf5(polygon, g) -> None:
    from .device import CompositePolygon

    if polygon.name:
        g.attrs["name"] = polygon.name
    if polygon.layer:
        g.attrs["layer"] = polygon.layer
    g["points"] = polygon.points
    if isinstance(polygon, CompositePolygon):  # (extension: the reference stores the shapely outline)
        rg = g.create_group("rings")
        for k, ring in enumerate(polygon.rings):
            rg[f"add_{k}"] = ring
        for k, ring in enumerate(polygon.cut_rings):
            rg[f"sub_{k}"] = ring


def polygon_from_hdf5(g):
    from .device import CompositePolygon, Polygon

    name, layer = g.attrs.get("name", None), g.attrs.get("layer", None)
    name = None if name is None else _attr_str(name)
    layer = None if layer is None else _attr_str(layer)
    if "rings" in g:
        rg = g["rings"]
        add = [np.array(rg[k]) for k in sorted(rg.keys()) if k.startswith("add_")]
        sub = [np.array(rg[k]) for k in sorted(rg.keys()) if k.startswith("sub_")]
        return CompositePolygon(name, layer=layer, add=add, sub=sub)
    return Polygon(name, layer=layer, points=np.asarray(g["points"]))


def mesh_to_hdf5(mesh, g, compress: bool = True) -> None:
    g["sites"] = mesh.sites
    g["elements"] = mesh.elements
    if not compress:
        g["triangle_centroids"] = mesh.triangle_centroids
        g["boundary_indices"] = mesh.boundary_indices
        g["vertex_areas"] = mesh.vertex_areas
        g["triangle_areas"] = mesh.triangle_areas
        e = g.create_group("edge_mesh")
        em = mesh.edge_mesh
        e["centers"], e["edges"], e["boundary_edge_indices"] = em.centers, em.edges, em.boundary_edge_indices
        e["directions"], e["edge_lengths"] = em.directions, em.edge_lengths


def mesh_from_hdf5(g):
    """The device-resident operators are rebuilt from the triangulation (the derived arrays of an
    uncompressed file are redundant with it)."""
    from .mesh import Mesh

    if not ("sites" in g and "elements" in g):
        raise IOError("Could not load mesh due to missing data.")
    return Mesh.from_triangulation(np.array(g["sites"]), np.array(g["elements"], dtype=np.int64))


def device_to_hdf5(device, path_or_group, save_mesh: bool = True, compress: bool = True) -> None:
    with open_group(path_or_group, "x") as g:
        g.attrs["name"] = device.name
        g.attrs["length_units"] = device.length_units
        g.attrs["solve_dtype"] = str(device.solve_dtype)
        groups = {k: g.create_group(k) for k in ("layers", "films", "holes", "terminals", "abstract_regions")}
        for name, layer in device.layers.items():
            layer_to_hdf5(layer, groups["layers"].create_group(name))
        for key, polygons in (("films", device.films), ("holes", device.holes),
                              ("abstract_regions", device.abstract_regions)):
            for name, polygon in polygons.items():
                polygon_to_hdf5(polygon, groups[key].create_group(name))
        for film_name, terminals in device.terminals.items():
            tg = groups["terminals"].create_group(film_name)
            for i, terminal in enumerate(terminals):
                polygon_to_hdf5(terminal, tg.create_group(str(i)))
        if save_mesh and device.meshes:
            mg = g.create_group("mesh")
            for name, mesh in device.meshes.items():
                mesh_to_hdf5(mesh, mg.create_group(name), compress=compress)


def device_from_hdf5(path_or_group):
    from .device import Device

    with open_group(path_or_group, "r") as g:
        terminals = {}
        for film, tg in g["terminals"].items():
            terminals[film] = [polygon_from_hdf5(tg[str(i)]) for i in range(len(tg))]
        device = Device(
            _attr_str(g.attrs["name"]),
            layers=[layer_from_hdf5(v) for v in g["layers"].values()],
            films=[polygon_from_hdf5(v) for v in g["films"].values()],
            holes=[polygon_from_hdf5(v) for v in g["holes"].values()],
            terminals=terminals or None,
            abstract_regions=[polygon_from_hdf5(v) for v in g["abstract_regions"].values()],
            length_units=_attr_str(g.attrs["length_units"]),
            solve_dtype=_attr_str(g.attrs["solve_dtype"]),
        )
        if "mesh" in g:
            device.meshes = {name: mesh_from_hdf5(mg) for name, mg in g["mesh"].items()}
        return device


# ----------------------------------------------------------------------------------------------
# solutions
# ----------------------------------------------------------------------------------------------
def vortex_to_hdf5(vortex, g) -> None:
    g.attrs["x"], g.attrs["y"], g.attrs["film"], g.attrs["nPhi0"] = vortex.x, vortex.y, vortex.film, vortex.nPhi0


def vortex_from_hdf5(g):
    from .solution import Vortex

    return Vortex(x=g.attrs["x"], y=g.attrs["y"], film=_attr_str(g.attrs["film"]), nPhi0=g.attrs["nPhi0"])


def film_solution_to_hdf5(fs, g) -> None:
    g["stream"] = fs.stream
    g["current_density"] = fs.current_density
    g["applied_field"] = fs.applied_field
    g["self_field"] = fs.self_field
    if fs.field_from_other_films is not None:
        g["field_from_other_films"] = fs.field_from_other_films


def film_solution_from_hdf5(g):
    from .solution import FilmSolution

    other = g.get("field_from_other_films", None)
    return FilmSolution(stream=np.array(g["stream"]), current_density=np.array(g["current_density"]),
                        applied_field=np.array(g["applied_field"]), self_field=np.array(g["self_field"]),
                        field_from_other_films=None if other is None else np.array(other))


def version_info() -> Dict[str, str]:
    """What every stored solution is stamped with (reference about.py:47-61; here: package, library ABI
    and CUDA device instead of the BLAS vendor)."""
    import platform

    import numpy

    info = {"superscreen_b200": _package_version(), "Numpy": numpy.__version__, "Python": platform.python_version()}
    try:
        import torch

        info["torch"] = torch.__version__
        if torch.cuda.is_available():
            info["GPU"] = torch.cuda.get_device_name(torch.cuda.current_device())
    except Exception:  # noqa: BLE001
        pass
    return info


def _package_version() -> str:
    from . import __version__

    return __version__


def solution_to_hdf5(solution, path_or_group, device_path: Optional[str] = None, compress: bool = True) -> None:
    with open_group(path_or_group, "x") as g:
        g.attrs["time_created"] = solution.time_created.isoformat()
        g.attrs["field_units"] = solution.field_units
        g.attrs["current_units"] = solution.current_units
        g.attrs["solver"] = solution.solver
        g.create_group("version_info").attrs.update(version_info())
        if device_path is None:
            device_to_hdf5(solution.device, g.create_group("device"), save_mesh=True, compress=compress)
        else:
            _soft_link(g, "device", device_path)
        fg = g.create_group("film_solutions")
        for name, fs in solution.film_solutions.items():
            film_solution_to_hdf5(fs, fg.create_group(name))
        vg = g.create_group("vortices")
        vortices = solution.vortices
        if isinstance(vortices, dict):
            vortices = [v for vs in vortices.values() for v in vs]
        for i, vortex in enumerate(vortices):
            vortex_to_hdf5(vortex, vg.create_group(str(i)))
        serialize_obj(g, solution.applied_field_func, "applied_field_func")
        g.create_group("circulating_currents").attrs.update(solution.circulating_currents)
        tg = g.create_group("terminal_currents")
        for film_name, currents in solution.terminal_currents.items():
            tg.create_group(film_name).attrs.update(currents)


def solution_from_hdf5(path_or_group):
    from .solution import Solution

    with open_group(path_or_group, "r") as g:
        device = device_from_hdf5(g["device"])
        film_solutions = {name: film_solution_from_hdf5(fg) for name, fg in g["film_solutions"].items()}
        vortices = [vortex_from_hdf5(g["vortices"][i]) for i in sorted(g["vortices"].keys(), key=int)]
        terminal_currents = {film: dict(tg.attrs) for film, tg in g["terminal_currents"].items()}
        solution = Solution(
            device=device, film_solutions=film_solutions, applied_field_func=deserialize_obj(g, "applied_field_func"),
            vortices=vortices, circulating_currents=dict(g["circulating_currents"].attrs),
            terminal_currents=terminal_currents, current_units=_attr_str(g.attrs["current_units"]),
            field_units=_attr_str(g.attrs["field_units"]), solver=_attr_str(g.attrs["solver"]))
        solution._time_created = dt.datetime.fromisoformat(_attr_str(g.attrs["time_created"]))
        solution._version_info = dict(g["version_info"].attrs)
    return solution


def save_solutions(solutions: Sequence, path_or_group, compress: bool = True) -> None:
    """reference solution.py:1031-1063: the device once, every solution soft-linked to it."""
    if not solutions:
        return
    with open_group(path_or_group, "x") as g:
        dg = g.create_group("device")
        device_to_hdf5(solutions[0].device, dg, compress=compress)
        for i, solution in enumerate(solutions):
            solution_to_hdf5(solution, g.create_group(str(i)), device_path=dg.name, compress=compress)


def load_solutions(path_or_group) -> List:
    with open_group(path_or_group, "r") as g:
        keys = sorted((k for k in g.keys() if str(k).isdigit()), key=int)
        return [solution_from_hdf5(g[k]) for k in keys]


# ----------------------------------------------------------------------------------------------
# factorized models
# ----------------------------------------------------------------------------------------------
def lambda_info_to_hdf5(info, g) -> None:
    g.attrs["film"] = info.film
    if info.london_lambda is not None:
        g["london_lambda"] = info.london_lambda
    if info.thickness is not None:
        g.attrs["thickness"] = info.thickness
    g["Lambda"] = info.Lambda


def film_info_to_hdf5(info, g) -> None:
    """reference solver/utils.py:134-164, dense ``kernel`` / ``laplacian`` (/ ``gradient``) included."""
    g.attrs["name"] = info.name
    g.attrs["layer"] = info.layer
    lambda_info_to_hdf5(info.lambda_info, g.create_group("lambda_info"))
    vg = g.create_group("vortices")
    for i, vortex in enumerate(info.vortices):
        vortex_to_hdf5(vortex, vg.create_group(str(i)))
    g["interior_indices"] = info.interior_indices
    g["boundary_indices"] = info.boundary_indices
    hg = g.create_group("hole_indices")
    for hole, indices in info.hole_indices.items():
        hg[hole] = indices
    g["in_hole"] = info.in_hole
    cg = g.create_group("circulating_currents")
    for hole, current in info.circulating_currents.items():
        cg.attrs[hole] = current
    g["weights"] = info.weights
    g["kernel"] = info.kernel
    g["laplacian"] = info.laplacian
    if info.gradient is not None:
        g["gradient"] = info.gradient
    if info.terminal_currents is not None:
        tg = g.create_group("terminal_currents")
        for name, current in info.terminal_currents.items():
            tg.attrs[name] = current


def linear_system_to_hdf5(system, g) -> None:
    """reference solver/solve_film.py:37-51: ``A``, ``indices``, ``lu`` + ``piv`` (scipy layout),
    ``grad_Lambda_term`` (dense array, or the scalar 0 as an attribute)."""
    g["A"] = system.A
    g["indices"] = system.indices
    if system.lu is not None:
        lu, piv = system.lu_piv
        g["lu"], g["piv"] = lu, piv
    T = system.grad_Lambda_term
    if isinstance(T, (int, float)):
        g.attrs["grad_Lambda_term"] = float(T)
    else:
        g["grad_Lambda_term"] = _dense_operator(system.film_info, T)


def _dense_operator(film_info, data) -> np.ndarray:
    """Dense (n, n) array of a device operator stored on the adjacency + I pattern."""
    import scipy.sparse as sp

    d = film_info.mesh._data
    return sp.csr_array((data.cpu().numpy(), d.host("op_indices"), d.host("op_indptr")), shape=(d.n, d.n)).toarray()


def terminal_systems_to_hdf5(ts, g) -> None:
    """reference solver/solve_film.py:102-120"""
    g.attrs["film"] = ts.film
    linear_system_to_hdf5(ts.boundary, g.create_group("boundary"))
    hg = g.create_group("holes")
    for hole, system in ts.holes.items():
        linear_system_to_hdf5(system, hg.create_group(hole))
    linear_system_to_hdf5(ts.film_without_boundary, g.create_group("film_without_boundary"))
    if ts.film_without_boundary_or_holes is not None:
        linear_system_to_hdf5(ts.film_without_boundary_or_holes, g.create_group("film_without_boundary_or_holes"))


def model_to_hdf5(model, path_or_group) -> None:
    """reference solver/solve.py:102-132"""
    with open_group(path_or_group, "x") as g:
        g.attrs["current_units"] = model.current_units
        device_to_hdf5(model.device, g.create_group("device"))
        ig = g.create_group("film_info")
        for film, info in model.film_info.items():
            film_info_to_hdf5(info, ig.create_group(film))
        sg = g.create_group("film_systems")
        for film, system in model.film_systems.items():
            linear_system_to_hdf5(system, sg.create_group(film))
        hg = g.create_group("hole_systems")
        for film, holes in model.hole_systems.items():
            fg = hg.create_group(film)
            for hole, system in holes.items():
                linear_system_to_hdf5(system, fg.create_group(hole))
        tg = g.create_group("terminal_systems")
        for film, systems in model.terminal_systems.items():
            terminal_systems_to_hdf5(systems, tg.create_group(film))
        cg = g.create_group("terminal_currents")
        for film, terminals in model.terminal_currents.items():
            cg.create_group(film).attrs.update(terminals)
        g.create_group("circulating_currents").attrs.update(model.circulating_currents)
        vg = g.create_group("vortices")
        vortices = model.vortices
        if isinstance(vortices, dict):
            vortices = [v for vs in vortices.values() for v in vs]
        for i, vortex in enumerate(vortices):
            vortex_to_hdf5(vortex, vg.create_group(str(i)))


def model_from_hdf5(path_or_group, comm=None):
    """reference solver/solve.py:134-180.  The stored device (triangulations included), currents and
    vortices define the model; operators and factors are rebuilt on the GPU."""
    from .solver.solve import factorize_model

    with open_group(path_or_group, "r") as g:
        device = device_from_hdf5(g["device"])
        if not device.meshes:
            raise IOError("The stored model has no meshes.")
        current_units = _attr_str(g.attrs["current_units"])
        terminal_currents = {film: dict(tg.attrs) for film, tg in g["terminal_currents"].items()}
        circulating_currents = dict(g["circulating_currents"].attrs)
        vortices = [vortex_from_hdf5(g["vortices"][i]) for i in sorted(g["vortices"].keys(), key=int)]
    return factorize_model(device=device, current_units=current_units, terminal_currents=terminal_currents or None,
                           circulating_currents=circulating_currents, vortices=vortices, comm=comm)
