"""Runs one stage of the hot path in isolation (for ncu): python tools/run_stage.py --stage getrf --n 20164"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import _lib
from superscreen_b200.geometry import box
from superscreen_b200.synthetic import square_mesh

ap = argparse.ArgumentParser()
ap.add_argument("--stage", default="getrf")
ap.add_argument("--n", type=int, default=20164)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--nrhs", type=int, default=1)
ap.add_argument("--sym", type=int, default=1)
ap.add_argument("--piv", type=int, default=0, help="1: time scb_getrf_piv (partial pivoting) instead")
args = ap.parse_args()
L = _lib.lib()
sites, elements = square_mesh(10.0, args.n, seed=0)
device = sc.Device("c2", layers=[sc.Layer("layer", Lambda=0.1, z0=0.0)],
                   films=[sc.Polygon("film", layer="layer", points=box(10.0, points=4))])
device.set_meshes({"film": (sites, elements)})
torch.cuda.synchronize()
if args.stage == "nbody":
    # kernel-only rates of the N-body family (CUDA events, inputs resident)
    rng = np.random.default_rng(0)
    n = args.n
    src3 = torch.as_tensor(np.column_stack([sites, np.zeros(len(sites))])).cuda().contiguous()
    src2 = torch.as_tensor(sites).cuda()
    area = torch.rand(len(sites), dtype=torch.float64, device="cuda")
    J = torch.randn(len(sites), 2, dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for m in (len(sites), 1_000_000):
        tgt3 = torch.as_tensor(np.column_stack([rng.uniform(-7, 7, m), rng.uniform(-7, 7, m), np.full(m, 1.0)])).cuda()
        tgt2 = tgt3[:, :2].contiguous()
        for kind, name, tgt, src, oc, flop in ((0, "film_to_film", tgt2, src2, 1, 20), (1, "Bz", tgt3, src3, 1, 22), (2, "Bvec", tgt3, src3, 3, 28)):
            out = torch.empty(m * oc, dtype=torch.float64, device="cuda")
            for rep in range(3):
                e0.record()
                _lib.check(L.scb_biot_savart(kind, m, _lib.ptr(tgt), len(sites), _lib.ptr(src), _lib.ptr(area), _lib.ptr(J), 0.5, 1.0, 1, _lib.ptr(out), _lib.stream_ptr()))
                e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1); pairs = m * len(sites)
            print(f"nbody {name}: targets={m} sources={len(sites)} {ms:.3f} ms  {pairs/ms*1e-6:.1f} Gpair/s  ~{pairs*flop/ms*1e-9:.2f} TFLOP/s (at {flop} flop/pair)")
    sys.exit(0)
if args.stage == "mesh":
    for _ in range(args.reps):
        sc.Mesh.from_triangulation(sites, elements)
    torch.cuda.synchronize(); sys.exit(0)
info = sc.solver.utils.make_film_info(device=device, vortices=[], circulating_currents={}, terminal_currents={})["film"]
info.dev["T"] = None
from superscreen_b200.solver.solve_film import assemble_negA, lu_solve, LinearSystem
ix = torch.as_tensor(info.interior_indices).cuda()
n_int = len(info.interior_indices); n_pad = -(-n_int // 128) * 128
M = torch.empty(n_pad, n_pad, dtype=torch.float64, device="cuda")
dinv = torch.empty(int(L.scb_getrf_dinv_bytes(n_pad)) // 8, dtype=torch.float64, device="cuda")
flag = torch.zeros(1, dtype=torch.int32, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
sym_full = torch.sqrt(info.mesh._data.t["vertex_areas"]) if args.sym else None
for rep in range(args.reps):
    assemble_negA(info, ix, n_int, n_pad, None, out=M, sym_scale_full=sym_full)
    torch.cuda.synchronize()
    e0.record()
    if args.piv:
        piv = torch.empty(n_pad, dtype=torch.int32, device="cuda"); perm = torch.empty(n_pad, dtype=torch.int32, device="cuda")
        _lib.check(L.scb_getrf_piv(n_pad, _lib.ptr(M), _lib.ptr(dinv), _lib.ptr(piv), _lib.ptr(perm), _lib.ptr(flag), _lib.stream_ptr()))
    else:
        fn = L.scb_getrf_sym_nopiv if args.sym else L.scb_getrf_nopiv
        _lib.check(fn(n_pad, _lib.ptr(M), _lib.ptr(dinv), _lib.ptr(flag), _lib.stream_ptr()))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    fl = (1/3 if args.sym else 2/3) * n_int**3
    print(f"getrf sym={args.sym} piv={args.piv} n_int={n_int} n_pad={n_pad}: {ms:.2f} ms  {fl/ms*1e-9:.2f} TFLOP/s executed  ({(2/3)*n_int**3/ms*1e-9:.2f} getrf-equivalent)  info={int(flag.item())}")
if args.stage == "getrs":
    system = LinearSystem(indices=info.interior_indices, film_info=info, n_pad=n_pad, lu=M, dinv=dinv, indices_dev=ix,
                          sym_scale=None if sym_full is None else sym_full[ix].contiguous())
    h = torch.randn(n_int, args.nrhs, dtype=torch.float64, device="cuda")
    ts = []
    for rep in range(args.reps + 1):
        e0.record(); x = lu_solve(system, h); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        print(f"getrs nrhs={args.nrhs}: {ts[-1]:.3f} ms")
    if len(ts) >= 5:
        print(f"getrs nrhs={args.nrhs} n_pad={n_pad}: steady state (median of the last {len(ts) - 3} calls) {float(np.median(ts[3:])):.3f} ms")
    # consistency with the single-RHS flag-driven sweeps on a few columns
    cols = sorted(set([0, args.nrhs // 2, args.nrhs - 1]))
    ref = torch.stack([lu_solve(system, h[:, c].contiguous()) for c in cols], dim=1)
    err = float((x[:, cols] - ref).norm() / ref.norm())
    print(f"getrs nrhs={args.nrhs}: rel-L2 vs single-RHS solves on columns {cols}: {err:.3e}")
