"""Multi-GPU host logic: genome sharding and hit-table merging.

The reference's unit of parallelism is one genome per worker (PEPPAN.py:922); here genome g goes
to rank g mod world, the exemplar query set is replicated, and the per-rank hit tables are
concatenated in rank order by pb_allgather_hits (NCCL).  `merge_hit_tables` is the host statement
of that concatenation (used by the CPU gloo tests and by callers that gather by other means)."""
import numpy as np


def shard_indices(n_items, rank, world):
    """indices of the items (genomes / query blocks) owned by `rank`: round-robin."""
    return list(range(rank, n_items, world))


def merge_hit_tables(tables):
    """tables: list over ranks of (hits structured array, cigar uint32 array).  Returns
    (hits, cigar, rank_offsets) exactly as pb_allgather_hits lays them out."""
    hits, cigs, roff, co = [], [], [0], 0
    for h, c in tables:
        h = h.copy()
        h['cigar_off'] = h['cigar_off'] + np.uint32(co)
        hits.append(h); cigs.append(c)
        co += len(c)
        roff.append(roff[-1] + len(h))
    if not hits:
        return np.zeros(0), np.zeros(0, np.uint32), np.array([0], np.int64)
    return np.concatenate(hits), np.concatenate(cigs).astype(np.uint32), np.array(roff, dtype=np.int64)


def init_context_from_env(backend_pg=None):
    """Create the per-process GPU context of a torchrun-style launch (RANK / LOCAL_RANK / WORLD_SIZE).
    The NCCL unique id is made by rank 0 and broadcast over the caller's (gloo) process group."""
    import os
    from ._lib import Context, nccl_unique_id
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    if world == 1:
        return Context(local)
    if backend_pg is None:
        raise ValueError('world > 1 needs a process group to distribute the NCCL unique id')
    # The exchanges of this path are small (a few MB per batch of genomes) and run BESIDE the search kernels of worker
    # contexts; a rank that arrives early spins in the collective until its peers arrive.  Two channels (= two thread blocks)
    # keep that wait off the SMs the searches need; the caller's own setting wins.
    os.environ.setdefault('NCCL_MAX_NCHANNELS', '2')
    os.environ.setdefault('NCCL_MIN_NCHANNELS', '1')
    obj = [nccl_unique_id() if rank == 0 else None]
    backend_pg.broadcast_object_list(obj, src=0)
    return Context(local, rank, world, obj[0])
