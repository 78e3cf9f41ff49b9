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
    obj = [nccl_unique_id() if rank == 0 else None]
    backend_pg.broadcast_object_list(obj, src=0)
    return Context(local, rank, world, obj[0])


_ABANDONED = []


def init_context_watchdog(pg, local, rank, world, timeout_s=120.0, make_uid=None, make_nccl_ctx=None, make_plain_ctx=None):
    """Context with an NCCL communicator, created under a watchdog.  The communicator set-up and a first (empty)
    hit-table allgather run in a helper thread; if ANY rank has not finished after `timeout_s` (or failed), EVERY rank
    falls back to a context without NCCL -- the decision is taken with a MIN all-reduce over `pg` (gloo) -- so a
    collective that cannot be established costs a bounded wait instead of hanging the job.
    Returns (ctx, None) or (plain ctx, note).  The make_* hooks exist for the CPU test of this control flow."""
    import threading
    import torch
    from ._lib import Context, nccl_unique_id

    def _nccl_ctx(uid):
        import ctypes as C
        from . import search as _s
        c = Context(local, rank, world, uid)
        _s.bind(c.lib)
        h = _s.Hits()
        c.check(c.lib.pb_allgather_hits(c.h, C.byref(h)), 'pb_allgather_hits')
        c.lib.pb_free_hits(C.byref(h))
        return c

    make_uid = make_uid or nccl_unique_id
    make_nccl_ctx = make_nccl_ctx or _nccl_ctx
    make_plain_ctx = make_plain_ctx or (lambda: Context(local))
    obj = [make_uid() if rank == 0 else None]
    pg.broadcast_object_list(obj, src=0)
    box = {}

    def work():
        try:
            box['ctx'] = make_nccl_ctx(obj[0])
        except Exception as e:          # reported in the note; the fallback context is made by the caller's thread
            box['err'] = repr(e)

    th = threading.Thread(target=work, daemon=True)
    th.start(); th.join(timeout_s)
    ok = torch.tensor([1.0 if 'ctx' in box else 0.0], dtype=torch.float64)
    pg.all_reduce(ok, op=pg.ReduceOp.MIN)
    if float(ok[0]) == 1.0:
        return box['ctx'], None
    _ABANDONED.append(box)              # keep a half-made communicator alive: destroying it could block as well
    why = box.get('err') or ('timeout' if 'ctx' not in box else 'another rank was not ready')
    return make_plain_ctx(), 'NCCL communicator / first allgather not ready on every rank after %.0f s (%s): hit tables not gathered' % (timeout_s, why)
