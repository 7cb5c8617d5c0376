"""Host/device time of the sharded 1M-point field evaluation for rank 0 of WORLD emulated ranks on one GPU
(collectives replaced by local copies) -- where the non-kernel time of C5 field_at_position goes.
    python tools/field_timing.py --world 8 [--n 60000]"""
import argparse, cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superscreen_b200 as sc
from superscreen_b200 import configs, parallel

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--n", type=int, default=20000, help="vertices of the source film (the sources only scale the kernel time)")
ap.add_argument("--profile", type=int, default=1)
args = ap.parse_args()


class EmulatedRankComm(parallel.Comm):
    def __init__(self, world):
        self.world, self.rank = world, 0

    def owner(self, index):
        return index % self.world

    def all_gather_into(self, out, send):
        out.view(self.world, -1).copy_(send.reshape(1, -1).expand(self.world, -1))

    def all_gather_chunks(self, chunk, sizes):
        m = max(sizes)
        return torch.cat([chunk[:s] if len(chunk) >= s else torch.cat([chunk, chunk[: s - len(chunk)]]) for s in sizes], dim=0)


device5, fields = configs.c5_large(args.n)
model5 = sc.factorize_model(device=device5, current_units="uA")
sol5 = sc.solve(model=model5, applied_field=sc.ConstantField(1.0))[0]
grid = configs.evaluation_grid(1000)
for world in sorted({1, args.world}):
    comm = EmulatedRankComm(world) if world > 1 else None
    fn = lambda: parallel.field_at_position_sharded(sol5, grid, comm=comm, units="mT")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(7):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"world {world}: field_at_position_sharded on {len(grid)} points x {len(device5.meshes['film'].sites)} sources: "
          f"median {1e3 * float(np.median(ts)):.3f} ms (rank 0 only)")
    if args.profile:
        pr = cProfile.Profile(); pr.enable()
        for _ in range(5):
            fn()
        pr.disable()
        s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
