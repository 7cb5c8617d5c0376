"""Sharding of the solve path over the GPUs of one node.

The path shards in three places, none of which needs a collective inside a kernel:

* independent films (and their factorizations) -> one owner rank per film (round robin);
* the film-to-film Jacobi iteration (reference solver/solve.py:491-547) -> one exchange step per
  iteration: every owner broadcasts its films' sheet current ``J`` (16*n bytes per film), then
  each rank evaluates the Biot-Savart sums for the films it owns;
* evaluation points of ``field_at_position`` -> contiguous chunks per rank, results all-gathered.

A single film's LU is never sharded ("replicas only").  Everything here is backend agnostic
(torch tensors on any device, ``torch.distributed`` with NCCL on GPUs or gloo on CPU); the
arithmetic is injected as callables, so the same driver runs the CUDA kernels in production and
the CPU oracle in the world_size-2 gloo tests.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple


def _dist():
    import torch.distributed as dist

    return dist


class Comm:
    """Single-process communicator: every film is owned by rank 0, exchanges are identities."""

    rank = 0
    world = 1

    def owner(self, index: int) -> int:
        return 0

    def broadcast(self, tensor, src: int):
        return tensor

    def all_gather_chunks(self, chunk, sizes: Sequence[int]):
        return chunk

    def barrier(self) -> None:
        return None


class DistComm(Comm):
    """torch.distributed communicator (NCCL over NVLink on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        dist = _dist()
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def owner(self, index: int) -> int:
        return index % self.world

    def broadcast(self, tensor, src: int):
        _dist().broadcast(tensor, src=src, group=self.group)
        return tensor

    def all_gather_chunks(self, chunk, sizes: Sequence[int]):
        """Concatenation over ranks of per-rank chunks with (possibly different) leading sizes."""
        import torch

        m = max(sizes)
        pad = torch.zeros((m,) + tuple(chunk.shape[1:]), dtype=chunk.dtype, device=chunk.device)
        pad[: chunk.shape[0]] = chunk
        out = [torch.empty_like(pad) for _ in range(self.world)]
        _dist().all_gather(out, pad, group=self.group)
        return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)

    def barrier(self) -> None:
        _dist().barrier(group=self.group)


def default_comm() -> Comm:
    """DistComm when torch.distributed is initialised with more than one rank, else Comm."""
    try:
        dist = _dist()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return DistComm()
    except Exception:
        pass
    return Comm()


def film_owners(film_names: Sequence[str], comm: Comm) -> Dict[str, int]:
    """Round-robin assignment film -> owner rank (deterministic on every rank)."""
    return {name: comm.owner(k) for k, name in enumerate(film_names)}


def split_range(m: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, near-equal [lo, hi) chunks of range(m); chunk r belongs to rank r."""
    base, extra = divmod(m, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def exchange_films(local: Dict[str, object], owners: Dict[str, int], shapes: Dict[str, Tuple[int, ...]],
                   comm: Comm, like) -> Dict[str, object]:
    """Every owner broadcasts its films' tensors; returns the full {film: tensor} on every rank.
    ``shapes`` gives the tensor shape of every film (needed by the non-owners to allocate)."""
    import torch

    if comm.world == 1:
        return dict(local)
    out = {}
    for name, owner in owners.items():
        if owner == comm.rank:
            t = local[name].contiguous()
        else:
            t = torch.empty(shapes[name], dtype=like.dtype, device=like.device)
        out[name] = comm.broadcast(t, owner)
    return out


def run_film_iterations(
    film_names: Sequence[str],
    owners: Dict[str, int],
    comm: Comm,
    solve_fn: Callable[[str, Optional[object]], Tuple[object, object, object]],
    coupling_fn: Callable[[str, object, str], object],
    zeros_fn: Callable[[str], object],
    j_shape_fn: Callable[[str], Tuple[int, ...]],
    iterations: int,
    film_scope: Optional[Callable[[str], object]] = None,
    join: Optional[Callable[[], None]] = None,
) -> List[Tuple[Dict[str, Tuple[object, object, object]], Optional[Dict[str, object]]]]:
    """The driver loop of reference solver/solve.py:454-547 for the films owned by this rank.

    Args:
        solve_fn(film, field_from_other_films | None) -> (g, J, self_field) for an owned film.
        coupling_fn(source_film, J_source, target_film) -> field of the source film's currents at
            the target film's sites (reference biot_savart_film_to_film).
        zeros_fn(film) -> zero field tensor for that film.
        j_shape_fn(film) -> shape of that film's J tensor.
        film_scope(film) -> context manager inside which all the work of one film and one Jacobi
            step is issued; ``join()`` is called after every step.  The films of a step are
            independent, so the CUDA backend runs them on one stream per film.

    Returns one ``(results, others)`` pair per solution (``iterations + 1`` of them when there are
    at least two films): ``results[film] = (g, J, self_field)`` and ``others[film]`` = field from the
    other films used for that solve, for the films this rank owns.
    """
    import contextlib

    scope = film_scope or (lambda f: contextlib.nullcontext())
    join = join or (lambda: None)
    mine = [f for f in film_names if owners[f] == comm.rank]
    results = {}
    for f in mine:
        with scope(f):
            results[f] = solve_fn(f, None)
    join()
    out = [(results, None)]
    if len(film_names) < 2 or iterations < 1:
        return out
    like = None
    for f in mine:
        like = results[f][1]
        break
    if like is None:  # a rank that owns nothing still takes part in the exchanges
        like = zeros_fn(film_names[0])
    shapes = {f: j_shape_fn(f) for f in film_names}
    for _ in range(iterations):
        # Jacobi step: all film-to-film fields from the previous iterate, then all re-solves
        J_all = exchange_films({f: results[f][1] for f in mine}, owners, shapes, comm, like)
        others, new = {}, {}
        for dst in mine:
            with scope(dst):
                acc = zeros_fn(dst)
                for src in film_names:  # (fixed summation order: results do not depend on the scopes)
                    if src != dst:
                        acc = acc + coupling_fn(src, J_all[src], dst)
                others[dst] = acc
                new[dst] = solve_fn(dst, acc)
        join()
        results = new
        out.append((results, others))
    return out


def gather_film_results(per_iteration, film_names: Sequence[str], owners: Dict[str, int], comm: Comm,
                        shape_fns: Dict[str, Callable[[str], Tuple[int, ...]]], like):
    """Replicates the owned results of ``run_film_iterations`` on every rank.  ``shape_fns`` maps the
    keys 'g', 'J', 'self', 'other' to film -> shape."""
    if comm.world == 1:
        return per_iteration
    full = []
    for results, others in per_iteration:
        g = exchange_films({f: r[0] for f, r in results.items()}, owners,
                           {f: shape_fns["g"](f) for f in film_names}, comm, like)
        J = exchange_films({f: r[1] for f, r in results.items()}, owners,
                           {f: shape_fns["J"](f) for f in film_names}, comm, like)
        sf = exchange_films({f: r[2] for f, r in results.items()}, owners,
                            {f: shape_fns["self"](f) for f in film_names}, comm, like)
        oth = None
        if others is not None:
            oth = exchange_films(others, owners, {f: shape_fns["other"](f) for f in film_names}, comm, like)
        full.append(({f: (g[f], J[f], sf[f]) for f in film_names}, oth))
    return full


def sharded_targets(m: int, comm: Comm) -> Tuple[int, int, List[int]]:
    """(lo, hi) of this rank's chunk of ``m`` evaluation points and the sizes of all chunks."""
    chunks = split_range(m, comm.world)
    lo, hi = chunks[comm.rank]
    return lo, hi, [b - a for a, b in chunks]


def field_at_position_sharded(solution, positions, *, zs=None, comm: Optional[Comm] = None, units=None,
                              vector: bool = False):
    """``Solution.field_at_position`` (or the vector screening field) with the evaluation points
    split over the ranks of ``comm`` (reference solution.py:611-831; sources replicated, targets
    sharded, one all-gather of the results).  Returns the full (m,) / (m, 3) array on every rank."""
    import numpy as np
    import torch

    comm = comm or default_comm()
    positions = np.atleast_2d(np.asarray(positions, dtype=np.float64))
    if positions.shape[1] == 3:
        zs_arr = positions[:, 2]
        positions = positions[:, :2]
    else:
        zs_arr = np.broadcast_to(np.asarray(zs, dtype=np.float64), (len(positions),)).copy()
    lo, hi, sizes = sharded_targets(len(positions), comm)
    if vector:
        local = solution.screening_field_at_position(positions[lo:hi], zs=zs_arr[lo:hi], vector=True, units=units,
                                                     with_units=False)
    else:
        local = solution.field_at_position(positions[lo:hi], zs=zs_arr[lo:hi], units=units, with_units=False)
    if comm.world == 1:
        return np.asarray(local)
    dev = torch.device(f"cuda:{torch.cuda.current_device()}") if torch.cuda.is_available() else torch.device("cpu")
    chunk = torch.as_tensor(np.ascontiguousarray(np.asarray(local, dtype=np.float64))).to(dev)
    return comm.all_gather_chunks(chunk, sizes).cpu().numpy()
