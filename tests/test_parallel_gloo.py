"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in superscreen_b200/parallel.py:
film ownership, the per-iteration J exchange of the film-to-film Jacobi loop, result gathering and
target sharding.  The arithmetic is injected from the CPU oracle, the communication pattern is the
production one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from superscreen_b200 import parallel

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, result_dict):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import port as oracle

        g = dict(np.load(os.path.join(GOLDEN, "two_rings.npz")))
        names = ["lower", "upper"]
        comm = parallel.DistComm()
        owners = parallel.film_owners(names, comm)
        assert owners == {"lower": 0, "upper": 1}
        films = {}
        for name in names:
            mesh = oracle.build_mesh(g[f"in_{name}_sites"], g[f"in_{name}_elements"],
                                     with_Q=(owners[name] == rank))
            film = oracle.OracleFilm(name=name, mesh=mesh, z0=float(g[f"in_{name}_z0"]), Lambda=g[f"in_{name}_Lambda"],
                                     interior_indices=g[f"in_{name}_interior_indices"],
                                     hole_indices={f"{name}_hole": g[f"in_{name}_hole_indices"]})
            if owners[name] == rank:  # one film factorization per rank
                oracle.factorize_film(film)
            films[name] = film
        conv = oracle.field_conversion_mT_to_uA_per_um()
        applied = {n: np.full(len(films[n].mesh.sites), float(g["in_applied_mT"]) * conv) for n in names}
        circ = {"lower_hole": 1000.0}

        def solve_fn(name, other):
            assert films[name].lu_piv is not None, "solve_fn called for a film this rank does not own"
            fs = oracle.solve_film(films[name], applied[name], circ, conv,
                                   field_from_other_films=None if other is None else other.numpy())
            return (torch.from_numpy(fs.stream), torch.from_numpy(np.ascontiguousarray(fs.current_density)),
                    torch.from_numpy(fs.self_field * conv))

        sizes = {n: len(films[n].mesh.sites) for n in names}
        layout = parallel.FilmLayout(names, sizes, owners, world)
        assert layout.chunk == max(sizes.values()) and layout.total == world * layout.chunk
        assert layout.rows("upper") == (layout.chunk, layout.chunk + sizes["upper"])
        assert layout.local_rows("upper") == (0, sizes["upper"])

        def coupling_fn(dst, J_all):
            # the packed J of ALL films arrives in one all-gather; sum the fields of the other films
            assert tuple(J_all.shape) == (layout.total, 2)
            acc = np.zeros(sizes[dst])
            for src in names:
                if src == dst:
                    continue
                lo, hi = layout.rows(src)
                acc += oracle.biot_savart_film_to_film(
                    films[src].mesh.sites, float(films[src].z0), films[src].mesh.vertex_areas,
                    np.ascontiguousarray(J_all[lo:hi].numpy()), films[dst].mesh.sites, float(films[dst].z0))
            return torch.from_numpy(acc)

        iterations = int(g["in_iterations"])
        # the per-film scope / join hooks (one CUDA stream per film in production) must wrap every
        # owned film exactly once per step and be joined after every step
        import contextlib

        log = []

        @contextlib.contextmanager
        def film_scope(name):
            log.append(("enter", name))
            yield
            log.append(("exit", name))

        ncoll = {"n": 0}
        orig = comm.all_gather_into

        def counting(out, send):
            ncoll["n"] += 1
            return orig(out, send)

        comm.all_gather_into = counting
        packer = parallel.ResultPacker(layout, comm, list(range(iterations + 1)), None,
                                       torch.zeros(1, dtype=torch.float64))
        per_iter = parallel.run_film_iterations(layout, comm, solve_fn, coupling_fn, iterations,
                                                film_scope=film_scope, join=lambda: log.append(("join",)),
                                                on_result=packer.put)
        assert ncoll["n"] == iterations, "exactly one collective per Jacobi step"
        assert log == [("enter", names[rank]), ("exit", names[rank]), ("join",)] * (iterations + 1)
        assert len(per_iter) == iterations + 1
        assert set(per_iter[0][0]) == {names[rank]}, "each rank solves only its own film"
        packer.scale_fields(1.0 / conv)
        local = packer.to_host(gather=False)
        assert local.films() == [names[rank]]
        full = packer.to_host(gather=True)
        assert ncoll["n"] == iterations + 1, "results are replicated by ONE all-gather after the last step"
        assert full.films() == names
        errs = []
        rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
        for it in range(iterations + 1):
            for n in names:
                gg, J, sf, other = full.film(it, 0, n)
                assert gg.flags.c_contiguous and J.flags.c_contiguous
                errs.append(rel(gg, g[f"out_it{it}_{n}_stream"]))
                errs.append(rel(J, g[f"out_it{it}_{n}_J"]))
                errs.append(rel(sf, g[f"out_it{it}_{n}_self_field"]))
                if it > 0:
                    errs.append(rel(other, g[f"out_it{it}_{n}_other"]))
                else:
                    assert other is None
        # target sharding + ragged all-gather
        m = 11
        lo, hi, sizes = parallel.sharded_targets(m, comm)
        chunk = torch.arange(lo, hi, dtype=torch.float64)[:, None] * torch.ones(1, 3, dtype=torch.float64)
        gathered = comm.all_gather_chunks(chunk, sizes)
        assert gathered.shape == (m, 3) and torch.equal(gathered[:, 0], torch.arange(m, dtype=torch.float64))
        result_dict[rank] = max(errs)
    finally:
        dist.destroy_process_group()


def test_two_rank_film_sharding_matches_reference_golden():
    world = 2
    port = _free_port()
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert set(results.keys()) == {0, 1}
    assert max(results.values()) < 1e-9, dict(results)


def test_split_and_single_process_comm():
    assert parallel.split_range(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert parallel.split_range(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    comm = parallel.Comm()
    assert parallel.film_owners(["a", "b", "c"], comm) == {"a": 0, "b": 0, "c": 0}
    t = torch.ones(3)
    assert parallel.all_gather_chunks_equal(t, comm) is t
    layout = parallel.FilmLayout(["a", "b", "c"], {"a": 5, "b": 7, "c": 2}, {"a": 0, "b": 1, "c": 0}, 2)
    assert layout.by_rank == [["a", "c"], ["b"]] and layout.chunk == 7 and layout.total == 14
    assert layout.rows("a") == (0, 5) and layout.rows("c") == (5, 7) and layout.rows("b") == (7, 14)
    assert layout.local_rows("c") == (5, 7) and layout.local_rows("b") == (0, 7)
    assert isinstance(parallel.default_comm(), parallel.Comm)
