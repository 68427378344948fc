"""Host-side multi-GPU plumbing: one process per GPU, torch.distributed for rendezvous only.

Two ways the path shards:
  * independent problems (config C4): problem index blocks per rank, no data-path collective;
    results are gathered at the end (`shard_range`, `gather_objects`);
  * one large problem (config C5): residual rows are partitioned (`row_partition`), each rank's
    handle gets an NCCL communicator (`attach_comm`) and the library does the collectives
    (allreduce of the partial normal equations, distributed PCG).
Everything here is plain Python over torch.distributed (gloo on CPU, nccl on GPUs); the numeric
work stays in libsaddlepoint_b200.so.
"""
import numpy as np


def shard_range(nitems, world, rank):
    """Contiguous block [lo, hi) of `nitems` independent problems owned by `rank` (balanced: the
    first nitems % world ranks get one extra)."""
    base, extra = divmod(nitems, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def row_partition(m, world, rank, align=1):
    """Contiguous residual-row block [lo, hi) of a global m-row Jacobian owned by `rank`.
    `align` keeps block boundaries on multiples (e.g. 2 so Rosenbrock row pairs stay together)."""
    units = (m + align - 1) // align
    lo, hi = shard_range(units, world, rank)
    return min(m, lo * align), min(m, hi * align)


def attach_comm(handle, dist=None):
    """Create the handle's NCCL communicator across the default process group.
    Rank 0 makes the 128-byte id, torch.distributed broadcasts it."""
    if dist is None:
        import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    box = [handle.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    handle.comm_init(box[0], rank, world)
    return rank, world


def gather_objects(obj, dist=None):
    """All ranks' `obj` as a list on every rank (results of independently solved problems)."""
    if dist is None:
        import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def merge_problem_results(gathered):
    """Flatten per-rank lists of (problem_index, result) into one list ordered by problem index."""
    flat = [item for part in gathered for item in part]
    flat.sort(key=lambda t: t[0])
    idx = [t[0] for t in flat]
    if idx != list(range(len(idx))):
        raise ValueError("problem shards do not tile the batch: %r" % (idx[:10],))
    return [t[1] for t in flat]


def local_values(global_colptr, global_rowidx, global_vals, row_begin, row_end):
    """The owned entries of a global CSC value array (global CSC order restricted to the owned
    rows): what spb_assemble expects after spb_analyze_partitioned."""
    keep = (global_rowidx >= row_begin) & (global_rowidx < row_end)
    return np.ascontiguousarray(global_vals[keep])


def column_partition(n, world, rank, halo):
    """Contiguous block of unknowns of `rank` for the subdomain (halo) mode; every block holds at least
    2*halo unknowns (the exchange slabs of the two neighbours must not overlap)."""
    lo, hi = shard_range(n, world, rank)
    if hi - lo < 2 * halo:
        raise ValueError("too many ranks for this problem: a rank needs at least 2*halo = %d unknowns" % (2 * halo))
    return lo, hi


def ring_neighbours(world, rank):
    return (rank - 1) % world, (rank + 1) % world
