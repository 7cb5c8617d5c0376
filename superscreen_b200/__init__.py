"""superscreen_b200: a B200-native (sm_100a) implementation of SuperScreen's solve hot path.

Public names mirror ``superscreen/__init__.py:1-20`` for the path in scope:
``solve``, ``factorize_model``, ``FactorizedModel``, ``Solution``, ``FilmSolution``, ``Fluxoid``,
``Vortex``, ``Device``, ``Layer``, ``Polygon``, ``Mesh``, ``fem``, ``distance``, ``sources``.
All arithmetic runs in ``libsc_b200.so`` (include/scb.h); there is no CPU fallback.
"""
from . import distance, fem, geometry, io, meshgen, parallel, sources, units
from .device import Device, Layer, Polygon
from .fluxoid import find_fluxoid_solution, make_fluxoid_polygons
from .mesh import Mesh, MeshOperators
from .meshgen import generate_mesh
from .solution import FilmSolution, Fluxoid, Solution, Vortex
from .solver import (
    FactorizedModel,
    FilmInfo,
    LambdaInfo,
    LinearSystem,
    convert_field,
    factorize_model,
    field_conversion_factor,
    solve,
    solve_batch,
)
from .sources import Constant, ConstantField, Parameter

__version__ = "0.1.0"
