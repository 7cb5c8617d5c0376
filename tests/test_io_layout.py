"""HDF5 wire layout (superscreen_b200/io.py) pinned through the in-memory group protocol: group, dataset
and attribute names must be the reference's (solution.py:132-164,936-1087, device/device.py:936-1016,
device/layer.py:108-138, device/polygon.py:621-634).  No GPU and no h5py needed; with h5py installed the
same code writes real files (the last test then runs too)."""
import numpy as np
import pytest

import superscreen_b200 as sc
from superscreen_b200 import io as scio
from superscreen_b200.geometry import box, circle


def _device():
    layers = [sc.Layer("base", london_lambda=0.5, thickness=0.05, z0=0.5), sc.Layer("top", Lambda=0.3, z0=1.0)]
    films = [sc.Polygon("ring", layer="base", points=circle(4.0, 40)),
             sc.Polygon("plate", layer="top", points=circle(1.0, 24)).union(box(1.0, 0.5, center=(1.2, 0.0)))]
    holes = [sc.Polygon("hole", layer="base", points=circle(2.0, 30))]
    return sc.Device("dev", layers=layers, films=films, holes=holes)


def _solution(device, it, rng):
    fs = {}
    for name, n in (("ring", 37), ("plate", 21)):
        fs[name] = sc.FilmSolution(stream=rng.normal(size=n), current_density=rng.normal(size=(n, 2)),
                                   applied_field=rng.normal(size=n), self_field=rng.normal(size=n),
                                   field_from_other_films=rng.normal(size=n) if it else None)
    return sc.Solution(device=device, film_solutions=fs, applied_field_func=sc.ConstantField(0.25), field_units="mT",
                       current_units="uA", circulating_currents={"hole": 1000.0}, terminal_currents={},
                       vortices=[sc.Vortex(x=3.0, y=0.1, film="ring")])


def test_device_layout_roundtrip():
    device = _device()
    g = scio.MemoryGroup()
    device.to_hdf5(g)
    assert set(g.keys()) == {"layers", "films", "holes", "terminals", "abstract_regions"}  # (no mesh attached)
    assert g.attrs["name"] == "dev" and g.attrs["length_units"] == "um" and g.attrs["solve_dtype"] == "float64"
    assert g["layers"]["base"].attrs["london_lambda"] == 0.5 and g["layers"]["base"].attrs["thickness"] == 0.05
    assert g["layers"]["top"].attrs["Lambda"] == 0.3 and "thickness" not in g["layers"]["top"].attrs
    assert np.array_equal(g["films"]["ring"]["points"], device.films["ring"].points)
    assert g["films"]["ring"].attrs["layer"] == "base"
    back = sc.Device.from_hdf5(g)
    assert list(back.films) == ["ring", "plate"] and list(back.holes) == ["hole"] and list(back.layers) == ["base", "top"]
    assert back.layers["base"].Lambda == pytest.approx(0.5**2 / 0.05) and back.layers["top"].Lambda == 0.3
    pts = np.array([[0.0, 0.0], [1.6, 0.1], [1.6, 0.4], [3.0, 3.0]])
    assert np.array_equal(back.films["plate"].contains_points(pts), device.films["plate"].contains_points(pts))
    with pytest.raises(ValueError):
        device.to_hdf5(g)  # groups already exist ("x" semantics)


def test_solutions_layout_roundtrip():
    device = _device()
    rng = np.random.default_rng(0)
    solutions = [_solution(device, it, rng) for it in range(3)]
    g = scio.MemoryGroup()
    sc.Solution.save_solutions(solutions, g)
    # reference solution.py:1031-1063: "device" + one group per solution, each soft-linking the device
    assert set(g.keys()) == {"device", "0", "1", "2"}
    s1 = g["1"]
    assert set(s1.keys()) == {"version_info", "device", "film_solutions", "vortices", "applied_field_func.pickle",
                              "circulating_currents", "terminal_currents"}
    assert {"time_created", "field_units", "current_units", "solver"} <= set(s1.attrs)
    assert set(s1["film_solutions"]["ring"].keys()) == {"stream", "current_density", "applied_field", "self_field",
                                                         "field_from_other_films"}
    assert "field_from_other_films" not in g["0"]["film_solutions"]["ring"]
    assert s1["circulating_currents"].attrs["hole"] == 1000.0
    assert s1["vortices"]["0"].attrs["film"] == "ring"
    loaded = sc.Solution.load_solutions(g)
    assert len(loaded) == 3
    for a, b in zip(solutions, loaded):
        assert b.field_units == "mT" and b.current_units == "uA" and b.time_created == a.time_created
        assert b.circulating_currents == {"hole": 1000.0} and b.vortices == a.vortices
        for name in ("ring", "plate"):
            fa, fb = a.film_solutions[name], b.film_solutions[name]
            assert np.array_equal(fa.stream, fb.stream) and np.array_equal(fa.current_density, fb.current_density)
            assert np.array_equal(fa.total_field, fb.total_field)
        assert b.applied_field_func(np.zeros(3), np.zeros(3), np.zeros(3)) == pytest.approx(0.25)
    # a single solution carries its own device when no link is given
    g1 = scio.MemoryGroup()
    solutions[0].to_hdf5(g1)
    assert "films" in g1["device"].keys()
    assert sc.Solution.from_hdf5(g1).film_solutions["plate"].stream.shape == (21,)


def test_hdf5_file_roundtrip_when_h5py_is_installed(tmp_path):
    h5py = pytest.importorskip("h5py")
    device = _device()
    rng = np.random.default_rng(1)
    solutions = [_solution(device, it, rng) for it in range(2)]
    path = tmp_path / "solutions.h5"
    sc.Solution.save_solutions(solutions, path)
    with h5py.File(path, "r") as f:
        assert set(f.keys()) == {"device", "0", "1"}
        assert isinstance(f["1"].get("device", getlink=True), h5py.SoftLink)
    loaded = sc.Solution.load_solutions(path)
    assert np.array_equal(loaded[1].film_solutions["ring"].stream, solutions[1].film_solutions["ring"].stream)
