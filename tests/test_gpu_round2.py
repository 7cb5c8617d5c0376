"""GPU parity tests added in round 2: the many-RHS tensor-core paths and the packed film-to-film
coupling against scipy / the CPU oracle (not against other kernels of this repo), the BASELINE
configurations closer to their full sizes, and the full-size C2 film against the oracle together
with bit-reproducibility of the symmetric factorization at 20k vertices."""
import os

import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-8


@pytest.fixture(scope="module")
def sc():
    import torch

    assert torch.cuda.is_available()
    import superscreen_b200 as sc

    return sc


def _square_device(sc, n_vertices, seed=3, Lambda=0.1):
    from superscreen_b200 import configs

    return configs.c2_square(n_vertices, seed=seed, Lambda=Lambda)


# ----------------------------------------------------------------------------------------
# many right-hand sides against scipy / the dense oracle operator
# ----------------------------------------------------------------------------------------
def test_many_rhs_against_scipy_and_dense_kernel(sc):
    """64 right-hand sides at > 6k vertices: scb_getrs_nopiv (blocked DMMA substitution) against
    scipy.linalg.lu_solve(lu_factor(-A), H) -- the reference's own call (solve_film.py:279,530) -- and the
    64-column matrix-free operator (kernel_gemm_kernel) against the oracle's dense Q @ (w * G)."""
    import scipy.linalg as la
    import torch

    from oracle import port
    from superscreen_b200.solver.solve_film import apply_operator, lu_solve

    device = _square_device(sc, 6400)
    mesh = device.meshes["film"]
    model = sc.factorize_model(device=device, current_units="uA")
    system, info = model.film_systems["film"], model.film_info["film"]
    n_int, n = len(system.indices), len(mesh.sites)
    assert n_int > 6000
    rng = np.random.default_rng(7)
    for nrhs in (64, 9):
        H = rng.normal(size=(n_int, nrhs))
        x = lu_solve(system, torch.as_tensor(H).cuda()).cpu().numpy()
        A = system.A  # dense -A re-assembled by the device; checked against the oracle below
        ref = la.lu_solve(la.lu_factor(-A), H)
        assert rel_l2(x, ref) <= 1e-10, (nrhs, rel_l2(x, ref))
    # the dense system matrix itself, against the oracle's build_system_2d
    omesh = port.build_mesh(mesh.sites, mesh.elements)
    Aref = port.build_system_2d(omesh.Q, omesh.vertex_areas, np.full(n, 0.1), omesh.laplacian.toarray(), 0,
                                system.indices)
    assert rel_l2(A, Aref) <= 1e-12
    # matrix-free Q @ (w * G) with 64 columns (tensor-core kernel GEMM) and 5 columns (CUDA cores)
    for ncol in (64, 5):
        G = rng.normal(size=(n, ncol))
        out = apply_operator(info, torch.as_tensor(G).cuda(), with_sparse=False).cpu().numpy()
        ref = omesh.Q @ (omesh.vertex_areas[:, None] * G)
        assert rel_l2(out, ref) <= 1e-11, (ncol, rel_l2(out, ref))


def test_film_coupling_kernel_against_oracle(sc):
    """scb_film_coupling (one launch per target film over the packed sources of all films, own segment
    skipped, zero-area padding rows) against the sum of the oracle's biot_savart_film_to_film."""
    import torch

    from oracle import port
    from superscreen_b200 import _lib

    L = _lib.lib()
    rng = np.random.default_rng(2)
    sizes = [700, 333, 1201]
    zs = [0.0, 0.5, 0.5]
    films = [(rng.uniform(-3, 3, (n, 2)) + 7.0 * k, rng.uniform(0.01, 0.02, n)) for k, n in enumerate(sizes)]
    pad = 57  # padding rows after every film (what a rank-major layout with unequal chunks produces)
    total = sum(sizes) + pad * len(sizes)
    for nsets in (1, 3, 8, 13):
        src = np.full((total, 3), 1e30)
        area = np.zeros(total)
        J = np.zeros((total, nsets, 2))
        rows = []
        o = 0
        for (xy, w), z in zip(films, zs):
            n = len(xy)
            src[o:o + n, :2], src[o:o + n, 2], area[o:o + n] = xy, z, w
            J[o:o + n] = rng.normal(size=(n, nsets, 2))
            rows.append((o, o + n))
            o += n + pad
        d = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
        src_d, area_d, J_d = d(src), d(area), d(J if nsets > 1 else J[:, 0, :])
        for k, ((xy, _), z) in enumerate(zip(films, zs)):
            lo, hi = rows[k]
            tgt_d = d(xy)
            out = torch.empty((len(xy), nsets) if nsets > 1 else (len(xy),), dtype=torch.float64, device="cuda")
            _lib.check(L.scb_film_coupling(len(xy), _lib.ptr(tgt_d), z, total, _lib.ptr(src_d), _lib.ptr(area_d),
                                           _lib.ptr(J_d), lo, hi, 1.0 / (4 * np.pi), nsets, _lib.ptr(out),
                                           _lib.stream_ptr()))
            got = out.cpu().numpy().reshape(len(xy), nsets)
            for s in range(nsets):
                ref = np.zeros(len(xy))
                for j, ((sxy, sw), sz) in enumerate(zip(films, zs)):
                    if j == k:
                        continue
                    a, b = rows[j]
                    ref += port.biot_savart_film_to_film(sxy, sz, sw, np.ascontiguousarray(J[a:b, s, :]), xy, z)
                assert rel_l2(got[:, s], ref) <= 1e-12, (nsets, k, s)
    # a film alone in the layout sees no other film
    out = torch.ones(sizes[0], dtype=torch.float64, device="cuda")
    _lib.check(L.scb_film_coupling(sizes[0], _lib.ptr(d(films[0][0])), 0.0, sizes[0], _lib.ptr(d(src[:sizes[0]])),
                                   _lib.ptr(d(area[:sizes[0]])), _lib.ptr(d(J[:sizes[0], 0, :])), 0, sizes[0],
                                   1.0, 1, _lib.ptr(out), _lib.stream_ptr()))
    assert float(out.abs().max()) == 0.0


# ----------------------------------------------------------------------------------------
# BASELINE configurations closer to full size
# ----------------------------------------------------------------------------------------
def _oracle_films(sc, device, circulating=None):
    from test_gpu_configs import oracle_films

    return oracle_films(sc, device, circulating)


def test_c3_susceptometer_3k_per_film(sc):
    from oracle import port
    from superscreen_b200 import configs
    from superscreen_b200.geometry import close_curve, points_in_polygon
    from test_gpu_configs import compare

    device, polygons = configs.c3_susceptometer(n_vertices=3000)
    assert min(len(m.sites) for m in device.meshes.values()) >= 2900
    films = _oracle_films(sc, device)
    model = sc.factorize_model(device=device, current_units="uA", circulating_currents={"fc_center": "1 mA"})
    sols = sc.solve(model=model, iterations=5)
    osols = port.solve(films, lambda x, y, z: 0 * x, circulating_currents={"fc_center": 1000.0}, iterations=5)
    for it in range(6):
        compare(sols[it], osols[it], list(device.films))
    by_name = {f.name: f for f in films}
    fl = sols[-1].hole_fluxoid("pl_center", points=polygons["pl_center"], with_units=False)
    of, osup = port.polygon_fluxoid(by_name["pl"], osols[-1]["pl"], close_curve(polygons["pl_center"]),
                                    lambda p, q: points_in_polygon(p, q))
    assert abs(sum(fl) - (of + osup)) <= TOL * abs(of + osup)


def test_c4_eight_rings(sc):
    """All 8 rings of C4 (2 x 4 grid, pitch 12 um), 8 x 8 mutual-inductance matrix with iterations=3,
    against one oracle solve per driven hole."""
    from oracle import port
    from superscreen_b200 import configs
    from superscreen_b200.geometry import close_curve, points_in_polygon

    device, polygons = configs.c4_ring_array(n_rings=8, n_vertices=1500)
    M = np.array(device.mutual_inductance_matrix(polygons, units="pH", iterations=3))
    assert M.shape == (8, 8)
    films = _oracle_films(sc, device)
    by_name = {f.name: f for f in films}
    holes = list(device.holes)
    conv_mA = 1e-3 / port.MU_0 * 1e-3
    Mref = np.zeros_like(M)
    for j, hole in enumerate(holes):
        osol = port.solve(films, lambda x, y, z: 0 * x, circulating_currents={hole: 1.0}, iterations=3,
                          field_conversion=conv_mA)[-1]
        for i, name in enumerate(holes):
            film = by_name[f"ring{i}"]
            f, s = port.polygon_fluxoid(film, osol[film.name], close_curve(polygons[name]),
                                        lambda p, q: points_in_polygon(p, q), current_to_A=1e-3)
            Mref[i, j] = (f + s) * port.PHI_0 / 1e-3 * 1e12
    assert np.abs(M - Mref).max() <= TOL * np.abs(Mref).max()
    assert np.abs(M - M.T).max() <= 0.10 * np.abs(M[0, 1])


def test_lambda_sweep_case(sc):
    """One point of the C5 Lambda sweep (configs.with_lambda: shared device-resident mesh, new
    factorization) with a 16-field batch, against the oracle."""
    from oracle import port
    from superscreen_b200 import configs

    device, fields = configs.c5_large(n_vertices=4000)
    mesh = device.meshes["film"]
    omesh = port.build_mesh(mesh.sites, mesh.elements)
    interior = np.setdiff1d(np.arange(len(mesh.sites)), omesh.boundary_indices)
    conv = port.field_conversion_mT_to_uA_per_um()
    for lam in (0.05, 0.4):
        model = sc.factorize_model(device=configs.with_lambda(device, lam), current_units="uA")
        assert np.array_equal(interior, model.film_systems["film"].indices)
        batch = sc.solve_batch(model=model, applied_fields=[sc.ConstantField(float(f)) for f in fields[:16]])
        film = port.OracleFilm(name="film", mesh=omesh, z0=0.0, Lambda=np.full(len(mesh.sites), lam),
                               interior_indices=interior, hole_indices={})
        port.factorize_film(film, keep_A=False)
        for b in (0, 15):
            ref = port.solve_film(film, np.full(len(mesh.sites), fields[b] * conv), {}, conv)
            fs = batch[b][0].film_solutions["film"]
            assert rel_l2(fs.stream, ref.stream) <= TOL
            assert rel_l2(fs.current_density, ref.current_density) <= TOL
            assert rel_l2(fs.total_field, ref.total_field) <= TOL


# ----------------------------------------------------------------------------------------
# full-size C2 against the oracle, and bit-reproducibility at 20k
# ----------------------------------------------------------------------------------------
def test_c2_full_size_against_oracle_and_reproducible(sc):
    import torch

    from oracle import port
    from superscreen_b200 import _lib
    from superscreen_b200.solver.solve_film import assemble_negA

    device = _square_device(sc, 20164, seed=0)
    mesh = device.meshes["film"]
    n = len(mesh.sites)
    assert n > 20000
    model = sc.factorize_model(device=device, current_units="uA")
    fs = sc.solve(model=model, applied_field=sc.ConstantField(1.0))[0].film_solutions["film"]
    # ---- oracle: the reference CPU path on the identical mesh (about 12 s on the GPU box's host) ----
    omesh = port.build_mesh(mesh.sites, mesh.elements)
    interior = np.setdiff1d(np.arange(n), omesh.boundary_indices)
    assert np.array_equal(interior, model.film_systems["film"].indices)
    film = port.OracleFilm(name="film", mesh=omesh, z0=0.0, Lambda=np.full(n, 0.1), interior_indices=interior,
                           hole_indices={})
    port.factorize_film(film, keep_A=False)
    conv = port.field_conversion_mT_to_uA_per_um()
    ref = port.solve_film(film, np.full(n, conv), {}, conv)
    assert rel_l2(fs.stream, ref.stream) <= TOL
    assert rel_l2(fs.current_density, ref.current_density) <= TOL
    assert rel_l2(fs.self_field, ref.self_field) <= TOL
    del film, omesh
    # ---- bit-reproducibility of the symmetric factorization at the full size, 3 repetitions ----
    L = _lib.lib()
    system, info = model.film_systems["film"], model.film_info["film"]
    assert system.sym_scale is not None, "constant Lambda takes the symmetric factorization"
    d = mesh._data
    sym_full = torch.sqrt(d.t["vertex_areas"])
    first = None
    for rep in range(3):
        M, _ = assemble_negA(info, system.indices_dev, len(system.indices), system.n_pad, None,
                             sym_scale_full=sym_full)
        dinv = torch.empty_like(system.dinv)
        flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(L.scb_getrf_sym_nopiv(system.n_pad, _lib.ptr(M), _lib.ptr(dinv), _lib.ptr(flag), _lib.stream_ptr()))
        torch.cuda.synchronize()
        assert int(flag.item()) == 0
        if first is None:
            first = M
            assert torch.equal(M, system.lu), "re-factorization differs from the model's factors"
        else:
            assert torch.equal(M, first), f"symmetric LU not bit-reproducible at 20k (repetition {rep})"
        del dinv


@pytest.mark.gpu
@pytest.mark.parametrize("symmetric", [True, False])
def test_lu_bit_reproducible_on_identical_input(sc, symmetric):
    """Four factorizations of bit-identical copies of one 20k system must agree bit for bit, and L U must
    reproduce the matrix.  (Regression test of the generic->async proxy fence in the trailing-update
    kernel: without it single 32-byte sectors of the TMA-staged operands were occasionally read wrong,
    which showed up as run-to-run different factors with errors of 1e-8 .. 1e-5.)"""
    import torch

    from superscreen_b200 import _lib
    from superscreen_b200.solver.solve_film import assemble_negA
    from superscreen_b200.solver.utils import make_film_info

    device = _square_device(sc, 20164, seed=0)
    info = make_film_info(device=device, vortices=[], circulating_currents={}, terminal_currents={})["film"]
    info.dev["T"] = None
    L = _lib.lib()
    ix = torch.as_tensor(info.interior_indices).cuda()
    n_int = len(info.interior_indices)
    n_pad = -(-n_int // 128) * 128
    sym_full = torch.sqrt(info.mesh._data.t["vertex_areas"]) if symmetric else None
    S, _ = assemble_negA(info, ix, n_int, n_pad, None, sym_scale_full=sym_full)
    getrf = L.scb_getrf_sym_nopiv if symmetric else L.scb_getrf_nopiv
    first = None
    for rep in range(4):
        M = S.clone()
        dinv = torch.zeros(int(L.scb_getrf_dinv_bytes(n_pad)) // 8, dtype=torch.float64, device="cuda")
        flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(getrf(n_pad, _lib.ptr(M), _lib.ptr(dinv), _lib.ptr(flag), _lib.stream_ptr()))
        torch.cuda.synchronize()
        assert int(flag.item()) == 0
        if first is None:
            first = M
        else:
            assert torch.equal(M, first), f"LU not bit-reproducible on identical input (repetition {rep})"
        del dinv
    rows = torch.arange(0, n_pad, 61, device="cuda")
    Lf = torch.tril(first, -1)[rows]
    Lf[torch.arange(len(rows), device="cuda"), rows] = 1.0
    err = float((Lf @ torch.triu(first) - S[rows]).abs().max() / S.abs().max())
    assert err <= 1e-12, err
    # the flag-driven substitution sweeps (1 / 8 / 64 right-hand sides) must reproduce themselves as well
    dinv = torch.zeros(int(L.scb_getrf_dinv_bytes(n_pad)) // 8, dtype=torch.float64, device="cuda")
    M = S.clone()
    _lib.check(getrf(n_pad, _lib.ptr(M), _lib.ptr(dinv), _lib.ptr(flag), _lib.stream_ptr()))
    g = torch.Generator(device="cuda").manual_seed(1)
    for nrhs in (1, 8, 64):
        B0 = torch.randn(n_pad, nrhs, dtype=torch.float64, device="cuda", generator=g)
        outs = []
        for rep in range(3):
            B = B0.clone()
            _lib.check(L.scb_getrs_nopiv(n_pad, _lib.ptr(M), _lib.ptr(dinv), nrhs, _lib.ptr(B), _lib.stream_ptr()))
            torch.cuda.synchronize()
            outs.append(B)
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), f"getrs nrhs={nrhs} not reproducible"
        # residual of the triangular solves against the factors: (L U) x = b
        x = outs[0]
        r = torch.tril(M, -1) @ (torch.triu(M) @ x) + torch.triu(M) @ x - B0
        assert float(r.abs().max() / B0.abs().max()) <= 1e-9


# ----------------------------------------------------------------------------------------
# partial pivoting (scb_getrf_piv)
# ----------------------------------------------------------------------------------------
def _vanishing_pivot_device(sc, n_vertices=900):
    """A square film whose Lambda(x, y) = s (eps + ramp) is tuned so that the LEADING diagonal entry of the
    system matrix cancels: A_00 = Q_00 w_0 - Lambda_0 lap_00 - (grad Lambda . grad)_00 = 0.  The matrix is
    well conditioned (cond ~ 2e4) and LAPACK's pivoted LU solves it to 1e-13; an unpivoted elimination
    divides by a pivot of rounding size."""
    from superscreen_b200.geometry import box
    from superscreen_b200.synthetic import square_mesh

    sites, elements = square_mesh(10.0, n_vertices, seed=4)
    mesh = sc.Mesh.from_triangulation(sites, elements)
    ops = mesh.operators
    interior = np.setdiff1d(np.arange(len(sites)), mesh.boundary_indices)
    i0 = int(interior[0])
    gx, gy, lap = ops.gradient_x.tocsr(), ops.gradient_y.tocsr(), ops.laplacian.tocsr()
    d = np.array([gx[i0, i0], gy[i0, i0]])
    d /= np.linalg.norm(d)
    x0, y0 = sites[i0]
    eps = 1e-3
    shape = lambda x, y: eps + np.clip((x - x0) * d[0] + (y - y0) * d[1], 0.0, None)
    Lh = shape(sites[:, 0], sites[:, 1])
    T00 = (gx @ Lh)[i0] * gx[i0, i0] + (gy @ Lh)[i0] * gy[i0, i0]
    c0 = Lh[i0] * lap[i0, i0] + T00
    assert c0 > 0
    qdw0 = ops.Q_diagonal[i0] * mesh.vertex_areas[i0]
    s = qdw0 / c0
    device = sc.Device("vanishing_pivot", layers=[sc.Layer("layer", Lambda=lambda x, y: s * shape(x, y), z0=0.0)],
                       films=[sc.Polygon("film", layer="layer", points=box(10.0, points=4))])
    device.set_meshes({"film": mesh})
    return device, s * Lh, interior


def _oracle_solution(device, Lambda, interior, field_mT=1.0):
    from oracle import port

    mesh = device.meshes["film"]
    film = port.OracleFilm(name="film", mesh=port.build_mesh(mesh.sites, mesh.elements), z0=0.0, Lambda=Lambda,
                           interior_indices=interior, hole_indices={})
    port.factorize_film(film)
    conv = port.field_conversion_mT_to_uA_per_um()
    return film, port.solve_film(film, np.full(len(mesh.sites), field_mT * conv), {}, conv)


def test_pivoted_factorization_on_vanishing_leading_pivot(sc, monkeypatch):
    import scipy.linalg as la

    device, Lambda, interior = _vanishing_pivot_device(sc)
    film, ref = _oracle_solution(device, Lambda, interior)
    assert abs(film.A[0, 0]) <= 1e-10 * np.abs(film.A[0]).max(), "the construction should cancel the leading pivot"
    # without pivoting the path must fail loudly (zero pivot, or refinement that cannot converge)
    monkeypatch.setenv("SCB_PIVOT", "0")
    with pytest.raises(np.linalg.LinAlgError):
        model = sc.factorize_model(device=device, current_units="uA")
        sc.solve(model=model, applied_field=sc.ConstantField(1.0))
    # default (auto): the non-dominant general system is factored with partial pivoting
    monkeypatch.delenv("SCB_PIVOT")
    model = sc.factorize_model(device=device, current_units="uA")
    system = model.film_systems["film"]
    assert system.piv is not None and not system.refine
    fs = sc.solve(model=model, applied_field=sc.ConstantField(1.0))[0].film_solutions["film"]
    assert rel_l2(fs.stream, ref.stream) <= TOL
    assert rel_l2(fs.current_density, ref.current_density) <= TOL
    assert rel_l2(fs.self_field, ref.self_field) <= TOL
    # lu_piv is a scipy-compatible (lu, piv) pair of -A: scipy's own lu_solve accepts it
    lu, piv = system.lu_piv
    assert piv.dtype == np.int32 and (piv != np.arange(len(piv))).any()
    rng = np.random.default_rng(0)
    h = rng.normal(size=len(piv))
    x_scipy = la.lu_solve((lu, piv), h)
    x_ref = la.lu_solve(film.lu_piv, h)
    assert rel_l2(x_scipy, x_ref) <= 1e-9
    # every pivot is the column maximum below the diagonal: |L| <= 1, as LAPACK guarantees
    assert np.abs(np.tril(lu, -1)).max() <= 1.0 + 1e-12
    # batched right-hand sides go through the same row permutation
    batch = sc.solve_batch(model=model, applied_fields=[sc.ConstantField(f) for f in (1.0, -2.0, 0.5)])
    assert rel_l2(batch[1][0].film_solutions["film"].stream, -2.0 * ref.stream) <= TOL


def test_forced_pivoting_matches_unpivoted_and_oracle(sc, monkeypatch):
    """SCB_PIVOT=1 on a well-behaved inhomogeneous film (golden square_inhomogeneous): same answer as the
    unpivoted general LU and as the live-reference golden solution."""
    from oracle import port
    from superscreen_b200 import configs

    device = configs.c2_square(2500, seed=9, Lambda=0.1)
    mesh = device.meshes["film"]
    lam = lambda x, y: 0.1 * (1.0 + 0.5 * np.sin(x) * np.cos(0.7 * y))
    device.layers["layer"].Lambda = lam
    interior = np.setdiff1d(np.arange(len(mesh.sites)), mesh.boundary_indices)
    _, ref = _oracle_solution(device, lam(mesh.sites[:, 0], mesh.sites[:, 1]), interior, field_mT=0.3)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SCB_PIVOT", mode)
        model = sc.factorize_model(device=device, current_units="uA")
        assert (model.film_systems["film"].piv is not None) == (mode == "1")
        out[mode] = sc.solve(model=model, applied_field=sc.ConstantField(0.3))[0].film_solutions["film"]
        assert rel_l2(out[mode].stream, ref.stream) <= TOL
        assert rel_l2(out[mode].total_field, ref.total_field) <= TOL
    assert rel_l2(out["1"].stream, out["0"].stream) <= 1e-11


# ----------------------------------------------------------------------------------------
# physical validation: the only measured numbers the reference publishes for this path
# ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("size,measured,sigma", [("small", 69.0, 7.0), ("medium", 166.0, 4.0)])
def test_ibm_susceptometer_mutual_inductance(sc, size, measured, sigma):
    """Field-coil <-> pickup-loop mutual inductance of the IBM scanning-SQUID susceptometers (reference
    docs/notebooks/scanning-squid.ipynb cell 3: 69 +- 7 and 166 +- 4 Phi_0/A measured; geometry from
    docs/notebooks/squids/ibm/{small,medium,layers}.py, closed field coil as in squids/mutuals.py:52-55,
    iterations=5).  Five films on three layers, composite (union) polygons, fluxoid = flux part +
    supercurrent part on a contour inside the pickup loop: an end-to-end check of the solve, the
    film-to-film iteration and `polygon_fluxoid` against an experiment (2 sigma window; the model lands
    within 1 sigma: 75.7 and 162.1 Phi_0/A at this mesh size, profiles/r02_ibm_susceptometer_validation.json)."""
    from superscreen_b200 import configs

    device, rings = configs.ibm_susceptometer(size, 6000)
    model = sc.factorize_model(device=device, current_units="uA", circulating_currents={"fc_center": "1 mA"})
    sols = sc.solve(model=model, iterations=5)
    assert len(sols) == 6
    M = [sum(s.hole_fluxoid("pl_center", points=rings["pl_center"], with_units=False)) / 1e-3 for s in sols]
    assert abs(M[0]) < 1e-6, "without film-to-film coupling no flux reaches the pickup loop"
    assert abs(M[-1] - measured) <= 2.0 * sigma, (size, M)
    assert abs(M[-1] - M[-2]) <= 0.02 * abs(M[-1]), "the Jacobi iteration has settled"
    # the fluxoid does not depend on the contour (to discretisation accuracy): a slightly wider ring
    if size == "small":
        from superscreen_b200.geometry import box

        M2 = sum(sols[-1].hole_fluxoid("pl_center", points=box(0.44, 2.68, points=240, center=(0.0, -1.125)),
                                       with_units=False)) / 1e-3
        assert abs(M2 - M[-1]) <= 0.05 * abs(M[-1]), (M2, M[-1])


# ----------------------------------------------------------------------------------------
# persistence: the reference's HDF5 layout for a factorized model (io.py), through the group protocol
# ----------------------------------------------------------------------------------------
def test_factorized_model_hdf5_layout_roundtrip(sc):
    import scipy.linalg as la

    from superscreen_b200 import configs
    from superscreen_b200 import io as scio

    device, _ = configs.c1_ring(900)
    model = sc.factorize_model(device=device, current_units="uA", circulating_currents={"ring_hole": "1 mA"},
                               vortices=[sc.Vortex(x=3.0, y=0.2, film="ring")])
    g = scio.MemoryGroup()
    model.to_hdf5(g)
    # reference solver/solve.py:102-132
    assert set(g.keys()) == {"device", "film_info", "film_systems", "hole_systems", "terminal_systems",
                             "terminal_currents", "circulating_currents", "vortices"}
    assert g.attrs["current_units"] == "uA" and g["circulating_currents"].attrs["ring_hole"] == pytest.approx(1000.0)
    fsys = g["film_systems"]["ring"]
    assert set(fsys.keys()) == {"A", "indices", "lu", "piv"} and fsys.attrs["grad_Lambda_term"] == 0.0
    A, lu, piv = np.array(fsys["A"]), np.array(fsys["lu"]), np.array(fsys["piv"])
    n_int = len(model.film_systems["ring"].indices)
    assert A.shape == (n_int, n_int) and lu.shape == (n_int, n_int) and piv.dtype == np.int32
    # the stored factors are scipy-compatible factors of -A (solver/solve_film.py:279)
    h = np.random.default_rng(0).normal(size=n_int)
    assert rel_l2(la.lu_solve((lu, piv), h), np.linalg.solve(-A, h)) <= 1e-10
    hsys = g["hole_systems"]["ring"]["ring_hole"]
    assert set(hsys.keys()) == {"A", "indices"} and np.array(hsys["A"]).shape == (len(device.meshes["ring"].sites),
                                                                                  len(np.array(hsys["indices"])))
    info = g["film_info"]["ring"]
    assert {"lambda_info", "vortices", "interior_indices", "boundary_indices", "hole_indices", "in_hole",
            "circulating_currents", "weights", "kernel", "laplacian"} <= set(info.keys())
    n = len(device.meshes["ring"].sites)
    assert np.array(info["kernel"]).shape == (n, n) and np.array(info["laplacian"]).shape == (n, n)
    assert set(g["device"]["mesh"]["ring"].keys()) == {"sites", "elements"}
    # load: operators and factors are rebuilt on the GPU; the solutions agree bit for bit
    back = sc.FactorizedModel.from_hdf5(g)
    assert back.circulating_currents == model.circulating_currents and len(back.vortices) == 1
    a = sc.solve(model=model, applied_field=sc.ConstantField(0.4))[0].film_solutions["ring"]
    store = scio.MemoryGroup()
    b = sc.solve(model=back, applied_field=sc.ConstantField(0.4), save_path=store)[0].film_solutions["ring"]
    assert np.array_equal(a.stream, b.stream) and np.array_equal(a.self_field, b.self_field)
    # solve(save_path=...) wrote the reference's file layout: /device + one group per iterate
    assert set(store.keys()) == {"device", "0"}
    c = sc.Solution.load_solutions(store)[0].film_solutions["ring"]
    assert np.array_equal(c.stream, a.stream)


@pytest.mark.gpu
def test_lu_latency_kernels_and_graph_replay_bit_identical():
    """The fine-tile kernels on the per-block chain (update_lat_kernel, trsm_sym_lat_kernel; SCB_LU_LAT) and
    the CUDA-graph replay of a repeated factorization (SCB_LU_GRAPH) only re-schedule the arithmetic of the
    full-size kernels: the factors and block inverses must agree bit for bit with both switched off."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def hashes(**env):
        e = dict(os.environ, **{k: str(v) for k, v in env.items()})
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "lu_factor_hash.py"), "1500", "5300"],
                           capture_output=True, text=True, env=e, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        return json.loads(r.stdout.strip().splitlines()[-1])

    base = hashes(SCB_LU_LAT=0, SCB_LU_GRAPH=0)
    for n, hs in base.items():
        assert len(set(hs)) == 1, f"n={n}: repeated factorizations differ"
    for env in (dict(SCB_LU_LAT=3, SCB_LU_GRAPH=0), dict(SCB_LU_LAT=1, SCB_LU_GRAPH=1), dict(SCB_LU_LAT=7, SCB_LU_GRAPH=0),
                dict(SCB_LU_LAT=7, SCB_LU_GRAPH=1)):
        got = hashes(**env)
        for n in base:
            assert set(got[n]) == set(base[n]), f"n={n} {env}: factors differ from the full-size kernels"


@pytest.mark.gpu
def test_deferred_factorization_check_raises_from_solve(sc):
    """`solve(device, ...)` / `mutual_inductance_matrix` enqueue the solve behind a factorization that is still
    running and read its zero-pivot flag only before the results are trusted: a broken system must still end
    in LinAlgError (as from `factorize_model`, which checks before it returns), never in a silent result."""
    from superscreen_b200.geometry import box
    from superscreen_b200.synthetic import square_mesh

    sites, elements = square_mesh(10.0, 900, seed=2)
    device = sc.Device("sq", layers=[sc.Layer("layer", Lambda=float("nan"), z0=0.0)],
                       films=[sc.Polygon("film", layer="layer", points=box(10.0, points=4))])
    device.set_meshes({"film": (sites, elements)})
    with pytest.raises(np.linalg.LinAlgError):
        sc.factorize_model(device=device, current_units="uA")
    with pytest.raises(np.linalg.LinAlgError):
        sc.solve(device, applied_field=sc.ConstantField(1.0))
    # and a healthy device right after it is unaffected
    good = sc.Device("sq", layers=[sc.Layer("layer", Lambda=0.5, z0=0.0)],
                     films=[sc.Polygon("film", layer="layer", points=box(10.0, points=4))])
    good.set_meshes({"film": (sites, elements)})
    sol = sc.solve(good, applied_field=sc.ConstantField(1.0))[0]
    assert np.isfinite(sol.film_solutions["film"].stream).all()


@pytest.mark.gpu
def test_solve_step_matches_the_separate_entry_points(sc):
    """`scb_solve_step` (one foreign call per film solve) is scb_solve_rhs -> scb_getrs_nopiv -> scb_solve_stream
    -> scb_current_density: bit-identical stream function and current density, batched right-hand sides."""
    import torch

    from superscreen_b200 import _lib, configs
    from superscreen_b200.solver.solve_film import hole_boundary_state

    L = _lib.lib()
    device, _ = configs.c1_ring(1500)
    model = sc.factorize_model(device=device, current_units="uA", circulating_currents={"ring_hole": "1 mA"})
    info, fsys = model.film_info["ring"], model.film_systems["ring"]
    d = info.mesh._data
    nrhs = 3
    rng = np.random.default_rng(0)
    applied = torch.as_tensor(rng.standard_normal((d.n, nrhs))).to(d.device).contiguous()
    other = torch.as_tensor(rng.standard_normal((d.n, nrhs))).to(d.device).contiguous()
    circ = {"ring_hole": torch.tensor([1.0, 0.0, -2.0], dtype=torch.float64, device=d.device)}
    g0, ha_eff = hole_boundary_state(info, model.hole_systems["ring"], circ, applied)
    n_int, n_pad = len(fsys.indices), fsys.n_pad
    s = _lib.stream_ptr()
    p = _lib.ptr
    # separate entry points
    B = torch.empty((n_pad, nrhs), dtype=torch.float64, device=d.device)
    _lib.check(L.scb_solve_rhs(n_int, n_pad, p(fsys.indices_dev), nrhs, p(applied), p(other), p(ha_eff),
                               p(fsys.sym_scale), p(B), s))
    _lib.check(L.scb_getrs_nopiv(n_pad, p(fsys.lu), p(fsys.dinv), nrhs, p(B), s))
    g_a = torch.empty_like(applied)
    _lib.check(L.scb_solve_stream(d.n, nrhs, p(fsys.pos), p(B), p(fsys.sym_scale), p(g0), p(g_a), s))
    J_a = torch.empty((d.n, nrhs, 2), dtype=torch.float64, device=d.device)
    _lib.check(L.scb_current_density(d.n, p(d.t["op_indptr"]), p(d.t["op_indices"]), p(d.t["gradient_x"]),
                                     p(d.t["gradient_y"]), nrhs, p(g_a), p(J_a), s))
    # one call
    B2 = torch.empty_like(B)
    g_b = torch.empty_like(applied)
    J_b = torch.empty_like(J_a)
    _lib.check(L.scb_solve_step(d.n, n_int, n_pad, nrhs, p(fsys.indices_dev), p(applied), p(other), p(ha_eff),
                                p(fsys.sym_scale), p(fsys.lu), p(fsys.dinv), p(B2), p(fsys.pos), p(g0), p(g_b),
                                p(d.t["op_indptr"]), p(d.t["op_indices"]), p(d.t["gradient_x"]),
                                p(d.t["gradient_y"]), p(J_b), s))
    torch.cuda.synchronize()
    assert torch.isfinite(g_a).all() and float(g_a.abs().max()) > 0
    assert torch.equal(g_a, g_b) and torch.equal(J_a, J_b)


@pytest.mark.gpu
def test_graph_replayed_mutual_inductance_and_untouched_user_arrays(sc):
    """Repeated `mutual_inductance_matrix` calls on one device go from direct launches to graph capture to
    graph replay of the factorizations (and through the per-device geometry caches): every call must return the
    same matrix bit for bit.  And `field_at_position` converts and sums its fields in place: never in an array
    that belongs to the user's applied-field function."""
    from superscreen_b200 import configs

    device, polys = configs.c4_ring_array(2, 1500)
    Ms = [np.array(device.mutual_inductance_matrix(polys, units="pH", iterations=2)) for _ in range(4)]
    for M in Ms[1:]:
        assert np.array_equal(M, Ms[0])
    assert np.isfinite(Ms[0]).all() and abs(Ms[0][0, 0]) > 10 * abs(Ms[0][0, 1]) > 0

    held = {}

    def applied(x, y, z):  # hands out the same array on every call
        if "a" not in held or held["a"].shape != np.shape(x):
            held["a"] = 0.25 + 0.01 * np.asarray(x, dtype=np.float64)
            held["copy"] = held["a"].copy()
        return held["a"]

    sol = sc.solve(device, applied_field=applied, field_units="mT", current_units="uA")[0]
    pos = np.column_stack([np.linspace(-3.0, 15.0, 257), np.linspace(-2.0, 2.0, 257), np.full(257, 1.5)])
    held.clear()
    f_mT = sol.field_at_position(pos, units="mT", with_units=False)
    f_uT = sol.field_at_position(pos, units="uT", with_units=False)
    assert np.array_equal(held["a"], held["copy"]), "the applied-field array of the user was modified"
    assert np.allclose(f_uT, 1e3 * f_mT, rtol=1e-13, atol=0)
    parts = sol.field_at_position(pos, units="mT", with_units=False, return_sum=False)
    assert np.allclose(sum(parts.values()), f_mT, rtol=1e-13, atol=1e-300)
