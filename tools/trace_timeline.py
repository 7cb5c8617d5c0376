"""Kernel timelines (CUPTI through torch.profiler; nsys is not in the image) of the latency-bound parts of
the path: one small-film factorization, its 8-column substitution, the C4 ring array on one GPU and the 20k
factorization.  Writes gpurun_out/<tag>/<section>.csv (ts_us, dur_us, stream, grid, kernel) and a summary.

    python tools/trace_timeline.py --tag trace [--sections film5k,getrs5k,c4,getrf20k]
"""
import argparse, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import superscreen_b200 as sc
from superscreen_b200 import _lib, configs
from superscreen_b200.geometry import box
from superscreen_b200.synthetic import square_mesh

ap = argparse.ArgumentParser()
ap.add_argument("--tag", default="trace")
ap.add_argument("--sections", default="film5k,getrs5k,c4,getrf20k")
ap.add_argument("--warm", type=int, default=2)
ap.add_argument("--emulate-world", type=int, default=0, help="c4 section: rank 0 of WORLD ranks emulated on one GPU")
args = ap.parse_args()
out_dir = os.path.join("gpurun_out", args.tag)
os.makedirs(out_dir, exist_ok=True)
L = _lib.lib()


def short(name):
    name = name.replace("scb::", "")
    return name.split("(")[0][:60]


def trace(section, fn, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    path = os.path.join(tempfile.gettempdir(), f"{section}.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    if not ev:
        print(f"[{section}] no kernel events"); return
    t0 = ev[0]["ts"]
    rows = [(e["ts"] - t0, e["dur"], e["args"].get("stream", -1), "x".join(map(str, e["args"].get("grid", []))), short(e["name"]))
            for e in ev]
    with open(os.path.join(out_dir, f"{section}.csv"), "w") as f:
        f.write("ts_us,dur_us,stream,grid,kernel\n")
        for r in rows:
            f.write("%.3f,%.3f,%s,%s,%s\n" % r)
    span = max(r[0] + r[1] for r in rows)
    # union of busy intervals and per-kernel totals
    busy, end = 0.0, -1.0
    for ts, dur, *_ in rows:
        if ts > end:
            busy += dur; end = ts + dur
        elif ts + dur > end:
            busy += ts + dur - end; end = ts + dur
    tot = {}
    for ts, dur, st, grid, name in rows:
        c, d = tot.get(name, (0, 0.0)); tot[name] = (c + 1, d + dur)
    lines = [f"[{section}] span {span / 1e3:.3f} ms, busy (>=1 kernel running) {busy / 1e3:.3f} ms, "
             f"sum of kernel durations {sum(r[1] for r in rows) / 1e3:.3f} ms, {len(rows)} device activities"]
    for name, (c, d) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:14]:
        lines.append(f"    {name:<60s} {c:6d} x  {d / 1e3:9.3f} ms  (avg {d / c:8.2f} us)")
    print("\n".join(lines))
    with open(os.path.join(out_dir, "summary.txt"), "a") as f:
        f.write("\n".join(lines) + "\n")


def film_setup(n):
    sites, elements = square_mesh(10.0, n, seed=0)
    device = sc.Device("c2", layers=[sc.Layer("layer", Lambda=0.1, z0=0.0)],
                       films=[sc.Polygon("film", layer="layer", points=box(10.0, points=4))])
    device.set_meshes({"film": (sites, elements)})
    info = sc.solver.utils.make_film_info(device=device, vortices=[], circulating_currents={}, terminal_currents={})["film"]
    info.dev["T"] = None
    from superscreen_b200.solver.solve_film import assemble_negA, LinearSystem
    ix = torch.as_tensor(info.interior_indices).cuda()
    n_int = len(info.interior_indices); n_pad = -(-n_int // 128) * 128
    M = torch.empty(n_pad, n_pad, dtype=torch.float64, device="cuda")
    dinv = torch.empty(int(L.scb_getrf_dinv_bytes(n_pad)) // 8, dtype=torch.float64, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    sym_full = torch.sqrt(info.mesh._data.t["vertex_areas"])
    M0 = torch.empty_like(M)
    assemble_negA(info, ix, n_int, n_pad, None, out=M0, sym_scale_full=sym_full)

    def factor():
        M.copy_(M0)
        _lib.check(L.scb_getrf_sym_nopiv(n_pad, _lib.ptr(M), _lib.ptr(dinv), _lib.ptr(flag), _lib.stream_ptr()))
    system = LinearSystem(indices=info.interior_indices, film_info=info, n_pad=n_pad, lu=M, dinv=dinv, indices_dev=ix,
                          sym_scale=sym_full[ix].contiguous())
    return factor, system, n_int


sections = args.sections.split(",")
if "film5k" in sections or "getrs5k" in sections:
    factor, system, n_int = film_setup(5300)
    if "film5k" in sections:
        trace("film5k_getrf", factor)
    if "getrs5k" in sections:
        from superscreen_b200.solver.solve_film import lu_solve
        factor(); torch.cuda.synchronize()
        for nrhs in (1, 8):
            h = torch.randn(n_int, nrhs, dtype=torch.float64, device="cuda")
            trace(f"film5k_getrs{nrhs}", lambda: lu_solve(system, h))
    del factor, system
    torch.cuda.empty_cache()
if "c4" in sections:
    device, polys = configs.c4_ring_array(8, 5000)
    if args.emulate_world > 1:
        from superscreen_b200 import parallel

        class EmulatedRankComm(parallel.Comm):
            def __init__(self, world):
                self.world, self.rank = world, 0

            def owner(self, index):
                return index % self.world

            def all_gather_into(self, out, send):
                out.view(self.world, -1).copy_(send.reshape(1, -1).expand(self.world, -1))

            def all_gather_chunks(self, chunk, sizes):
                return torch.cat([chunk] * self.world, dim=0)

        comm = EmulatedRankComm(args.emulate_world)
        trace(f"c4_rank0of{args.emulate_world}",
              lambda: device.mutual_inductance_matrix(polys, units="pH", iterations=5, comm=comm), warm=args.warm)
    else:
        trace("c4_1gpu", lambda: device.mutual_inductance_matrix(polys, units="pH", iterations=5), warm=args.warm)
if "getrf20k" in sections:
    factor, system, n_int = film_setup(20164)
    trace("film20k_getrf", factor, warm=1)
