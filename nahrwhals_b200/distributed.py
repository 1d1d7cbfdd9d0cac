"""Multi-GPU drivers.

* `shard_regions` / `solve_regions`: regionfile and whole-genome modes -- regions are independent units, each rank
  solves its share with its own context, there is NO collective on the data path (the reference runs one solver
  process per region from PSOCK workers, R/nahrwhals.R:171-215).
* `init_sharded` / `solve_sharded`: ONE deep search sharded over the ranks of a launcher (torchrun): every rank
  runs the same deterministic expansion / dedup / selection, the banded DP of each level is divided between the
  ranks, and each BFS level ends with one NCCL all-gather of the score arrays issued by the library itself on the
  context's stream (include/nwsolve.h, nw_dist_* / nw_solve_sharded).  torch.distributed is only used once, to hand
  rank 0's communicator id to the other ranks.  Results are bit-identical to the single-GPU search.
* `solve_sharded_local`: the same search with all ranks inside one process (one context per rank, on one GPU or on
  several; the exchange is a set of peer copies).  Used by the GPU tests when fewer GPUs than ranks are available
  and by `nwsolve --devices`.
"""
import ctypes as C

import numpy as np

from . import _lib
from .solver import Solver, _p


# --------------------------------------------------------------------------------------------- region sharding
def shard_regions(costs, world):
    """Static longest-processing-time partition of regions over ranks.  costs[k] = estimated work of region k
    (e.g. root_pairs ** depth).  Returns a list of index lists, one per rank; deterministic."""
    order = sorted(range(len(costs)), key=lambda k: (-costs[k], k))
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for k in order:
        r = min(range(world), key=lambda q: (load[q], q))
        out[r].append(k)
        load[r] += float(costs[k])
    for r in range(world):
        out[r].sort()
    return out


def estimate_region_cost(aln, maxdepth, maxwidth=None):
    """Work estimate for LPT sharding: (#block pairs sharing an alignment) ** depth, capped by the beam width."""
    a = np.asarray(aln) != 0
    pairs = int(np.triu((a.astype(np.int32) @ a.T.astype(np.int32)) > 0, 1).sum())
    per_level = max(pairs, 1)
    cost, frontier = 0.0, 1.0
    for _ in range(maxdepth):
        gen = frontier * per_level
        cost += gen * a.shape[0] * a.shape[1]
        frontier = gen if maxwidth is None else min(gen, float(maxwidth))
    return cost


def solve_regions(solver, regions, rank=0, world=1, **opts):
    """regions: list of (aln, len_x, len_y).  Solves this rank's share; returns {region index: SolveResult}."""
    costs = [estimate_region_cost(r[0], opts.get("maxdepth", 3), opts.get("maxwidth")) for r in regions]
    mine = shard_regions(costs, world)[rank]
    return {k: solver.solve(*regions[k], **opts) for k in mine}


# --------------------------------------------------------------------------------------------- sharded single search
def own_chunk(n, world):
    """Items per rank of the per-level score exchange (host mirror of own_chunk in csrc/nw_kernels.cuh)."""
    c = ((n + world - 1) // world + 3) & ~3
    return max(c, 4)


def own_range(n, world, rank):
    """[lo, hi) of the fresh candidates (by discovery rank) whose DP rank `rank` runs."""
    c = own_chunk(n, world)
    return min(rank * c, n), min((rank + 1) * c, n)


def broadcast_id(dist, id_bytes, src=0, group=None, device=None):
    """Hands rank `src`'s communicator id (bytes; ignored on the other ranks) to every rank of the group through one
    torch.distributed broadcast of a uint8 tensor (CPU tensor under gloo, device tensor under nccl)."""
    import torch
    n = _lib.DIST_ID_BYTES
    t = torch.zeros(n, dtype=torch.uint8, device=device if device is not None else "cpu")
    if dist.get_rank(group) == src:
        if id_bytes is None or len(id_bytes) != n:
            raise ValueError("the source rank needs a %d-byte id" % n)
        t.copy_(torch.frombuffer(bytearray(id_bytes), dtype=torch.uint8))
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().numpy().tobytes())


def init_sharded(solver, group=None):
    """Collective over the torch.distributed group: creates the NCCL communicator the library uses for
    nw_solve_sharded on `solver` (this rank's nahrwhals_b200.Solver, bound to this rank's GPU)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ident = None
    if rank == 0:
        buf = C.create_string_buffer(_lib.DIST_ID_BYTES)
        _lib.check(solver.L.nw_dist_unique_id(buf))
        ident = buf.raw
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else None
    ident = broadcast_id(dist, ident, 0, group, dev)
    _lib.check(solver.L.nw_dist_init(solver.h, ident, rank, world))
    return rank, world


def solve_sharded(solver, aln, len_x, len_y, **opts):
    """One deep search over all ranks that took part in init_sharded (collective: same problem and options on every
    rank).  Every rank returns the same SolveResult (stats: times / launches / bytes are this rank's)."""
    aln, lx, ly = Solver._problem(aln, len_x, len_y)
    o = solver._opts(**opts)
    r = C.c_void_p()
    _lib.check(solver.L.nw_solve_sharded(solver.h, _p(aln), aln.shape[0], aln.shape[1], _p(lx), _p(ly), C.byref(o),
                                         C.byref(r)))
    try:
        return solver._unpack(r, o)
    finally:
        solver.L.nw_result_free(r)


def solve_sharded_local(solvers, aln, len_x, len_y, **opts):
    """The sharded search with len(solvers) ranks inside this process (nw_solve_sharded_ctxs): one host thread per
    context, contexts on one GPU (emulated ranks) or on several.  Returns the list of per-rank SolveResults."""
    aln, lx, ly = Solver._problem(aln, len_x, len_y)
    s0 = solvers[0]
    o = s0._opts(**opts)
    world = len(solvers)
    hs = (C.c_void_p * world)(*[s.h for s in solvers])
    outs = (C.c_void_p * world)()
    _lib.check(s0.L.nw_solve_sharded_ctxs(hs, world, _p(aln), aln.shape[0], aln.shape[1], _p(lx), _p(ly), C.byref(o), outs))
    res = []
    try:
        for r in range(world):
            res.append(s0._unpack(C.c_void_p(outs[r]), o))
    finally:
        for r in range(world):
            if outs[r]:
                s0.L.nw_result_free(C.c_void_p(outs[r]))
    return res
