"""Sharding of the solve path over the GPUs of one node.

The path shards in three places, none of which needs a collective inside a kernel:

* independent films (and their factorizations) -> one owner rank per film (round robin);
* the film-to-film Jacobi iteration (reference solver/solve.py:491-547) -> one exchange step per
  iteration: ONE all-gather of the ranks' packed sheet currents ``J`` (16*n bytes per film and
  right-hand side), then each rank evaluates the Biot-Savart sums for the films it owns, one
  kernel launch per target film over the packed sources of all other films;
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

    def all_gather_into(self, out, send) -> None:
        out.copy_(send)

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
        """Concatenation over ranks of per-rank chunks with (possibly different) leading sizes: one
        flat all-gather of the chunks padded to the largest size."""
        import torch

        m = max(sizes)
        if chunk.shape[0] == m:
            pad = chunk.contiguous()
        else:
            pad = torch.zeros((m,) + tuple(chunk.shape[1:]), dtype=chunk.dtype, device=chunk.device)
            pad[: chunk.shape[0]] = chunk
        out = torch.empty((self.world * m,) + tuple(chunk.shape[1:]), dtype=chunk.dtype, device=chunk.device)
        self.all_gather_into(out, pad)
        if all(s_ == m for s_ in sizes):
            return out
        return torch.cat([out[r * m: r * m + s_] for r, s_ in enumerate(sizes)], dim=0)

    def all_gather_into(self, out, send) -> None:
        """out[(world * k, ...)] <- concatenation over ranks of send[(k, ...)] (one collective)."""
        dist = _dist()
        try:
            dist.all_gather_into_tensor(out, send, group=self.group)
        except (RuntimeError, NotImplementedError):  # backends without the flat variant
            dist.all_gather(list(out.chunk(self.world, dim=0)), send, group=self.group)

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


class FilmLayout:
    """Packed, rank-major arrangement of the vertices of all films.

    Rank ``r`` owns the rows ``[r * chunk, (r + 1) * chunk)``: its films back to back in device
    order, then zero padding up to ``chunk`` = the largest per-rank total.  With this layout the
    per-iteration exchange of the sheet currents is ONE equal-size all-gather of the ranks' chunks
    (SURVEY.md section 8e), and every rank sees the sources of all films as one contiguous array in
    which a film is the row range ``rows(film)``.
    """

    def __init__(self, film_names: Sequence[str], sizes: Dict[str, int], owners: Dict[str, int], world: int):
        self.films = list(film_names)
        self.sizes = {f: int(sizes[f]) for f in self.films}
        self.owners = dict(owners)
        self.world = int(world)
        self.by_rank = [[f for f in self.films if owners[f] == r] for r in range(self.world)]
        self.chunk = max(1, max(sum(self.sizes[f] for f in fs) for fs in self.by_rank))
        self.total = self.world * self.chunk
        self.offset: Dict[str, int] = {}
        for r, fs in enumerate(self.by_rank):
            o = r * self.chunk
            for f in fs:
                self.offset[f] = o
                o += self.sizes[f]

    def rows(self, film: str) -> Tuple[int, int]:
        """[lo, hi) of a film in the packed array."""
        lo = self.offset[film]
        return lo, lo + self.sizes[film]

    def local_rows(self, film: str) -> Tuple[int, int]:
        """[lo, hi) of a film inside its owner's chunk."""
        lo = self.offset[film] - self.owners[film] * self.chunk
        return lo, lo + self.sizes[film]


def all_gather_chunks_equal(send, comm: "Comm"):
    """One all-gather of equal-size per-rank chunks: (chunk, ...) on every rank -> (world * chunk, ...)."""
    if comm.world == 1:
        return send
    import torch

    out = torch.empty((comm.world * send.shape[0],) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    comm.all_gather_into(out, send.contiguous())
    return out


def run_film_iterations(
    layout: FilmLayout,
    comm: Comm,
    solve_fn: Callable[[str, Optional[object]], Tuple[object, object, object]],
    coupling_fn: Callable[[str, object], object],
    iterations: int,
    film_scope: Optional[Callable[[str], object]] = None,
    join: Optional[Callable[[], None]] = None,
    on_result: Optional[Callable[[int, str, Tuple[object, object, object], Optional[object]], None]] = None,
    j_like=None,
):
    """The driver loop of reference solver/solve.py:454-547 for the films owned by this rank.

    Args:
        solve_fn(film, field_from_other_films | None) -> (g, J, self_field) for an owned film;
            ``J`` has shape ``(n, 2)`` or, for batched right-hand sides, ``(n, B, 2)``.
        coupling_fn(target_film, J_all) -> field of ALL other films' currents at the target film's
            sites (the sum of reference biot_savart_film_to_film over the other films), where
            ``J_all`` is the packed ``(layout.total, ...)`` array of every film's ``J``.
        film_scope(film) -> context manager inside which all the work of one film and one Jacobi
            step is issued; ``join()`` is called after every step.  The films of a step are
            independent, so the CUDA backend runs them on one stream per film.
        on_result(iteration, film, (g, J, self_field), other) is called (inside the film's scope)
            for every owned film and iterate, e.g. to pack the result for a single download.
        j_like: a tensor with the dtype / device / trailing shape of a film's ``J`` (needed only
            by a rank that owns no film and still has to take part in the exchange).

    One collective per Jacobi step: the all-gather of the ranks' packed ``J`` chunks.  Jacobi
    ordering as in the reference: all film-to-film fields come from the previous iterate
    (solve.py:495-515) before any film is re-solved (:519-536).

    Returns one ``(results, others)`` pair per solution (``iterations + 1`` of them when there are
    at least two films) for the films this rank owns.
    """
    import contextlib

    import torch

    scope = film_scope or (lambda f: contextlib.nullcontext())
    join = join or (lambda: None)
    mine = layout.by_rank[comm.rank]
    results = {}
    for f in mine:
        with scope(f):
            results[f] = solve_fn(f, None)
            if on_result is not None:
                on_result(0, f, results[f], None)
    join()
    out = [(results, None)]
    if len(layout.films) < 2 or iterations < 1:
        return out
    send = None
    for it in range(iterations):
        # Jacobi step: pack my films' J, ONE all-gather, all film-to-film fields, all re-solves
        if send is None:
            like = results[mine[0]][1] if mine else j_like
            if like is None:
                raise ValueError("a rank that owns no film needs j_like")
            send = torch.zeros((layout.chunk,) + tuple(like.shape[1:]), dtype=like.dtype, device=like.device)
        for f in mine:
            lo, hi = layout.local_rows(f)
            send[lo:hi] = results[f][1]
        J_all = all_gather_chunks_equal(send, comm)
        others, new = {}, {}
        for dst in mine:
            with scope(dst):
                others[dst] = coupling_fn(dst, J_all)
                new[dst] = solve_fn(dst, others[dst])
                if on_result is not None:
                    on_result(it + 1, dst, new[dst], others[dst])
        join()
        results = new
        out.append((results, others))
    return out


class ResultPacker:
    """Keeps the results of every owned film and stored iterate in ONE flat buffer on the compute
    device, so that a whole solve needs a single (optional) all-gather and a single download.

    Layout of the flat buffer (``S`` stored iterates, ``B`` right-hand sides, ``R = layout.chunk``
    rows): ``A[S, 3B, R]`` = stream / self field / field from the other films, batch-major, followed
    by ``J[S, B, R, 2]``.  Every (iterate, right-hand side, film) slice is contiguous, so the host
    side hands out zero-copy views.
    """

    def __init__(self, layout: FilmLayout, comm: Comm, iterates: Sequence[int], batch: Optional[int], like):
        import torch

        self.layout, self.comm = layout, comm
        self.iterates = list(iterates)               # iterate numbers that are stored
        self.slot = {it: k for k, it in enumerate(self.iterates)}
        self.batched = batch is not None
        self.B = B = int(batch) if batch is not None else 1
        S, R = len(self.iterates), layout.chunk
        self.nA, self.nJ = S * 3 * B * R, S * B * R * 2
        self.flat = torch.zeros(self.nA + self.nJ, dtype=like.dtype, device=like.device)
        self.A = self.flat[: self.nA].view(S, 3 * B, R)
        self.J = self.flat[self.nA:].view(S, B, R, 2)

    def put(self, it: int, film: str, result, other) -> None:
        k = self.slot.get(it)
        if k is None:
            return
        g, J, self_field = result
        B = self.B
        lo, hi = self.layout.local_rows(film)
        n = hi - lo
        self.A[k, 0:B, lo:hi] = g.reshape(n, B).t()
        self.A[k, B:2 * B, lo:hi] = self_field.reshape(n, B).t()
        if other is not None:
            self.A[k, 2 * B:3 * B, lo:hi] = other.reshape(n, B).t()
        self.J[k, :, lo:hi, :] = J.reshape(n, B, 2).permute(1, 0, 2)

    def scale_fields(self, factor: float) -> None:
        """Multiplies the self / other field sections (solver units -> field units)."""
        self.A[:, self.B:, :].mul_(factor)

    def to_host(self, gather: bool, to_numpy: Optional[Callable] = None):
        """-> HostResults.  ``gather``: replicate every rank's results on all ranks (one all-gather);
        otherwise only the films owned by this rank are available."""
        flat = self.flat
        if gather and self.comm.world > 1:
            flat = all_gather_chunks_equal(flat.view(1, -1), self.comm)  # (world, len)
            ranks = list(range(self.comm.world))
        else:
            flat = flat.view(1, -1)
            ranks = [self.comm.rank]
        host = to_numpy(flat) if to_numpy is not None else flat.cpu().numpy()
        return HostResults(self, host, ranks)


class HostResults:
    """Host views of a ResultPacker buffer: ``film(iterate, b, name)`` -> (stream, J, self_field,
    field_from_other_films | None) as contiguous numpy views."""

    def __init__(self, packer: ResultPacker, host, ranks: Sequence[int]):
        self.p = packer
        S, B, R = len(packer.iterates), packer.B, packer.layout.chunk
        self.A = {r: host[k, : packer.nA].reshape(S, 3 * B, R) for k, r in enumerate(ranks)}
        self.J = {r: host[k, packer.nA:].reshape(S, B, R, 2) for k, r in enumerate(ranks)}

    @property
    def iterates(self) -> List[int]:
        return self.p.iterates

    def films(self) -> List[str]:
        return [f for f in self.p.layout.films if self.p.layout.owners[f] in self.A]

    def film(self, it: int, b: int, name: str):
        p = self.p
        k, B = p.slot[it], p.B
        r = p.layout.owners[name]
        lo, hi = p.layout.local_rows(name)
        A, J = self.A[r], self.J[r]
        other = A[k, 2 * B + b, lo:hi] if it > 0 else None
        return A[k, b, lo:hi], J[k, b, lo:hi, :], A[k, B + b, lo:hi], other


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
    full = comm.all_gather_chunks(chunk, sizes)
    if full.is_cuda:
        from .solver.solve import _to_host

        return _to_host(full)  # pinned staging: the full result is 8 bytes x m on every rank
    return full.cpu().numpy()
