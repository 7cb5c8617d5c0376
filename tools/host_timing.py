"""Host-side section times (SCB_HOST_TIMING=1 ranges, enqueue cost without device synchronisation) and the
synchronised wall time of the C4 mutual-inductance call and the C3 solve on one GPU.
    SCB_HOST_TIMING=1 python tools/host_timing.py [--reps 8]"""
import argparse, os, sys, time
os.environ.setdefault("SCB_HOST_TIMING", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import _lib, configs, parallel

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--warm", type=int, default=8)
ap.add_argument("--emulate-world", type=int, default=0,
                help="time the work of rank 0 of a WORLD-rank run on one GPU (collectives replaced by local copies: "
                     "same kernels and host work per rank, wrong numbers)")
args = ap.parse_args()


class EmulatedRankComm(parallel.Comm):
    """Rank 0 of a `world`-rank job without the other ranks: every rank slot of an all-gather receives this
    rank's chunk.  For timing the per-rank critical path only."""

    def __init__(self, world):
        self.world, self.rank = world, 0

    def owner(self, index):
        return index % self.world

    def all_gather_into(self, out, send):
        out.view(self.world, -1).copy_(send.reshape(1, -1).expand(self.world, -1))

    def all_gather_chunks(self, chunk, sizes):
        return torch.cat([chunk] * self.world, dim=0)


def run(label, fn):
    for _ in range(args.warm):
        fn()
    torch.cuda.synchronize()
    _lib.host_times.clear()
    ts = []
    for _ in range(args.reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print(f"== {label}: median {1e3 * float(np.median(ts)):.3f} ms, min {1e3 * min(ts):.3f} ms over {args.reps} calls; "
          f"host sections summed over those calls:")
    print(_lib.host_timing_report())


device, polys = configs.c4_ring_array(8, 5000)
if args.emulate_world > 1:
    comm = EmulatedRankComm(args.emulate_world)
    run(f"C4 mutual_inductance_matrix(iterations=5), rank 0 of {args.emulate_world} emulated on one GPU",
        lambda: device.mutual_inductance_matrix(polys, units="pH", iterations=5, comm=comm))
    sys.exit(0)
run("C4 mutual_inductance_matrix(iterations=5), 8 rings x 5k, 1 GPU",
    lambda: device.mutual_inductance_matrix(polys, units="pH", iterations=5))
device3, polys3 = configs.c3_susceptometer(4000)
run("C3 factorize_model", lambda: sc.factorize_model(device=device3, current_units="uA",
                                                      circulating_currents={"fc_center": "1 mA"}))
model3 = sc.factorize_model(device=device3, current_units="uA", circulating_currents={"fc_center": "1 mA"})
run("C3 solve(iterations=5)", lambda: sc.solve(model=model3, iterations=5))
