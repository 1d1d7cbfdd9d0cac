"""Multi-GPU drivers.

* `shard_regions` / `solve_regions`: regionfile and whole-genome modes -- regions are independent units, each rank
  solves its share with its own context, there is NO collective on the data path (the reference runs one solver
  process per region from PSOCK workers, R/nahrwhals.R:171-215).
* `solve_sharded`: ONE deep search sharded over the ranks of a torch.distributed (NCCL) group: the frontier is
  replicated, parents are dealt round-robin, and each BFS level ends with one small all-gather of 40-byte records
  followed by the same deterministic merge on every rank (include/nwsolve.h, nw_dist_*).  Results are bit-identical
  to the single-GPU search.
* `solve_sharded_emulated`: the same protocol with all ranks emulated inside one process on one GPU (several
  contexts, "all-gather" = concatenation).  Used by the GPU tests when fewer GPUs than ranks are available.
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


# --------------------------------------------------------------------------------------------- variable-length gather
def pad_gather_layout(counts):
    """Layout of a padded all-gather: every rank sends stride = max(counts) items.  Returns (stride, offsets)."""
    stride = int(max(counts)) if len(counts) else 0
    return stride, [r * stride for r in range(len(counts))]


def all_gather_varlen(dist, local, count, item_bytes, group=None):
    """All-gather of `count` items of `item_bytes` bytes held in the uint8 tensor `local` (device or CPU).
    Returns (all uint8 tensor of world*stride*item_bytes, counts list, stride)."""
    import torch
    world = dist.get_world_size(group)
    cnt = torch.tensor([count], dtype=torch.int64, device=local.device)
    allc = torch.zeros(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(allc, cnt, group=group)
    counts = [int(v) for v in allc.tolist()]
    stride, _ = pad_gather_layout(counts)
    send = torch.zeros(max(stride, 1) * item_bytes, dtype=torch.uint8, device=local.device)
    send[:count * item_bytes] = local[:count * item_bytes]
    out = torch.zeros(world * max(stride, 1) * item_bytes, dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, send, group=group)
    return out, counts, max(stride, 1)


# --------------------------------------------------------------------------------------------- sharded single search
class _Rank:
    """Per-rank wrapper of the nw_dist_* protocol (device buffers are torch tensors)."""

    def __init__(self, solver, rank, world):
        self.s, self.rank, self.world = solver, rank, world
        self.L = solver.L

    def begin(self, aln, lx, ly, opts):
        aln, lx, ly = Solver._problem(aln, lx, ly)
        self.o = self.s._opts(**opts)
        _lib.check(self.L.nw_dist_begin(self.s.h, _p(aln), aln.shape[0], aln.shape[1], _p(lx), _p(ly), C.byref(self.o),
                                        self.rank, self.world))

    def expand(self):
        n = C.c_int64()
        _lib.check(self.L.nw_dist_expand(self.s.h, C.byref(n)))
        return n.value

    def pack(self, torch, n, device):
        buf = torch.zeros(max(n, 1) * _lib.DIST_RECORD_BYTES, dtype=torch.uint8, device=device)
        torch.cuda.synchronize(device)
        _lib.check(self.L.nw_dist_pack(self.s.h, C.c_void_p(buf.data_ptr())))
        return buf

    def merge(self, torch, allbuf, counts, stride):
        cnt = np.ascontiguousarray(counts, dtype=np.int64)
        more = C.c_int32()
        torch.cuda.synchronize(allbuf.device)
        _lib.check(self.L.nw_dist_merge(self.s.h, C.c_void_p(allbuf.data_ptr()), _p(cnt), stride, C.byref(more)))
        return bool(more.value)

    def report_begin(self, torch, device):
        n = self.L.nw_dist_n_ids(self.s.h)
        needed = torch.zeros(n + 16, dtype=torch.uint8, device=device)
        torch.cuda.synchronize(device)
        _lib.check(self.L.nw_dist_report_begin(self.s.h, C.c_void_p(needed.data_ptr()), n + 16))
        return needed

    def mark(self, torch, needed, p):
        torch.cuda.synchronize(needed.device)
        _lib.check(self.L.nw_dist_report_mark(self.s.h, C.c_void_p(needed.data_ptr()), p))

    def edges(self, torch, needed):
        n = C.c_int64()
        torch.cuda.synchronize(needed.device)
        _lib.check(self.L.nw_dist_report_edges(self.s.h, C.c_void_p(needed.data_ptr()), C.byref(n)))
        buf = torch.zeros(max(n.value, 1) * _lib.DIST_EDGE_BYTES, dtype=torch.uint8, device=needed.device)
        torch.cuda.synchronize(needed.device)
        _lib.check(self.L.nw_dist_report_pack(self.s.h, C.c_void_p(buf.data_ptr())))
        return n.value, buf

    def finish(self, torch, needed, alledges, counts, stride):
        cnt = np.ascontiguousarray(counts, dtype=np.int64)
        r = C.c_void_p()
        torch.cuda.synchronize(needed.device)
        _lib.check(self.L.nw_dist_finish(self.s.h, C.c_void_p(needed.data_ptr()), C.c_void_p(alledges.data_ptr()),
                                         _p(cnt), stride, C.byref(r)))
        try:
            return self.s._unpack(r, self.o)
        finally:
            self.L.nw_result_free(r)


def solve_sharded(solver, aln, len_x, len_y, group=None, **opts):
    """One deep search over all ranks of a torch.distributed group (backend nccl).  Every rank returns the same
    SolveResult.  `solver` is this rank's nahrwhals_b200.Solver (bound to this rank's GPU)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    device = torch.device("cuda", torch.cuda.current_device())
    R = _Rank(solver, rank, world)
    R.begin(aln, len_x, len_y, opts)
    more = True
    while more:
        n = R.expand()
        send = R.pack(torch, n, device)
        allbuf, counts, stride = all_gather_varlen(dist, send, n, _lib.DIST_RECORD_BYTES, group)
        more = R.merge(torch, allbuf, counts, stride)
    needed = R.report_begin(torch, device)
    for p in range(1, R.o.maxdepth + 2):
        R.mark(torch, needed, p)
        dist.all_reduce(needed, op=dist.ReduceOp.MAX, group=group)
    ne, ebuf = R.edges(torch, needed)
    alle, ecounts, estride = all_gather_varlen(dist, ebuf, ne, _lib.DIST_EDGE_BYTES, group)
    return R.finish(torch, needed, alle, ecounts, estride)


def solve_sharded_emulated(solvers, aln, len_x, len_y, **opts):
    """The sharded protocol with len(solvers) ranks emulated in this process (all contexts on one GPU); the
    collectives become concatenations / element-wise maxima.  Returns the list of per-rank SolveResults."""
    import torch
    world = len(solvers)
    device = torch.device("cuda", torch.cuda.current_device())
    ranks = [_Rank(s, r, world) for r, s in enumerate(solvers)]
    for R in ranks:
        R.begin(aln, len_x, len_y, opts)
    more = True
    rb = _lib.DIST_RECORD_BYTES
    while more:
        ns = [R.expand() for R in ranks]
        sends = [R.pack(torch, n, device) for R, n in zip(ranks, ns)]
        stride = max(max(ns), 1)
        allbuf = torch.zeros(world * stride * rb, dtype=torch.uint8, device=device)
        for r, (b, n) in enumerate(zip(sends, ns)):
            allbuf[r * stride * rb: r * stride * rb + n * rb] = b[:n * rb]
        mores = [R.merge(torch, allbuf, ns, stride) for R in ranks]
        assert len(set(mores)) == 1, "ranks disagree on continuing the search"
        more = mores[0]
    needs = [R.report_begin(torch, device) for R in ranks]
    for p in range(1, ranks[0].o.maxdepth + 2):
        for R, nd in zip(ranks, needs):
            R.mark(torch, nd, p)
        mx = torch.stack(needs).max(dim=0).values
        for nd in needs:
            nd.copy_(mx)
    eb = _lib.DIST_EDGE_BYTES
    es = [R.edges(torch, nd) for R, nd in zip(ranks, needs)]
    ecounts = [e[0] for e in es]
    estride = max(max(ecounts), 1)
    alle = torch.zeros(world * estride * eb, dtype=torch.uint8, device=device)
    for r, (n, b) in enumerate(es):
        alle[r * estride * eb: r * estride * eb + n * eb] = b[:n * eb]
    return [R.finish(torch, nd, alle, ecounts, estride) for R, nd in zip(ranks, needs)]
