"""world_size-2 gloo test (CPU) of the multi-GPU host logic: sharding and hit-table merging."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _fake_table(rank, n):
    from peppan_b200.search import HIT_DTYPE
    rng = np.random.default_rng(100 + rank)
    h = np.zeros(n, dtype=HIT_DTYPE)
    h['q_id'] = rng.integers(0, 50, n); h['s_id'] = rank; h['raw_score'] = rng.integers(50, 500, n)
    lens = rng.integers(1, 5, n)
    h['cigar_n'] = lens; h['cigar_off'] = np.concatenate([[0], np.cumsum(lens)[:-1]])
    c = rng.integers(4, 4000, int(lens.sum())).astype(np.uint32)
    return h, c


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from peppan_b200 import dist as pbd
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = pbd.shard_indices(7, rank, world)
    h, c = _fake_table(rank, 5 + 3 * rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, (h, c, mine))
    hits, cig, roff = pbd.merge_hit_tables([(g[0], g[1]) for g in gathered])
    ok = roff.tolist() == [0, 5, 13] and sorted(sum((g[2] for g in gathered), [])) == list(range(7))
    for r in range(world):
        hr, cr = _fake_table(r, 5 + 3 * r)
        seg = hits[roff[r]:roff[r + 1]]
        ok = ok and np.array_equal(seg['raw_score'], hr['raw_score'])
        for i in range(len(seg)):
            ok = ok and np.array_equal(cig[seg['cigar_off'][i]:seg['cigar_off'][i] + seg['cigar_n'][i]],
                                       cr[hr['cigar_off'][i]:hr['cigar_off'][i] + hr['cigar_n'][i]])
    t = __import__('torch').tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        open(out, 'w').write('ok' if float(t[0]) == 1.0 else 'bad')
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_merge_and_sharding(tmp_path):
    out = os.path.join(tmp_path, 'res.txt')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == 'ok'


def test_shard_indices_partition():
    from peppan_b200 import dist as pbd
    for world in (1, 2, 4, 8):
        allidx = sorted(sum((pbd.shard_indices(37, r, world) for r in range(world)), []))
        assert allidx == list(range(37))


class _FakeCtx(object):
    def __init__(self, kind):
        self.kind = kind


def _watchdog_worker(rank, world, port, out, scenario):
    sys.path.insert(0, ROOT)
    import time
    import torch.distributed as dist
    from peppan_b200 import dist as pbd
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)

    def make_nccl(uid):
        assert uid == b'uid-from-rank-0'
        if scenario == 'hang' and rank == 1:
            time.sleep(30)                       # a communicator that never comes up on one rank
        if scenario == 'error' and rank == 0:
            raise RuntimeError('no NCCL here')
        return _FakeCtx('nccl')

    t0 = time.time()
    ctx, note = pbd.init_context_watchdog(dist, 0, rank, world, timeout_s=1.5, make_uid=lambda: b'uid-from-rank-0',
                                          make_nccl_ctx=make_nccl, make_plain_ctx=lambda: _FakeCtx('plain'))
    dt = time.time() - t0
    if scenario == 'fine':
        ok = ctx.kind == 'nccl' and note is None
    else:
        ok = ctx.kind == 'plain' and note is not None and dt < 20
        if scenario == 'error' and rank == 0:
            ok = ok and 'no NCCL here' in note
    t = __import__('torch').tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        open(out, 'w').write('ok' if float(t[0]) == 1.0 else 'bad')
    dist.barrier()
    if scenario == 'hang':
        os._exit(0)                              # as bench.py does: a helper thread is still asleep
    dist.destroy_process_group()


def test_nccl_watchdog_agrees_on_fallback(tmp_path):
    """every rank ends with the same decision: NCCL context when all ranks made one in time, plain context otherwise"""
    for scenario in ('fine', 'error', 'hang'):
        out = os.path.join(tmp_path, 'wd_%s.txt' % scenario)
        mp.spawn(_watchdog_worker, args=(2, _free_port(), out, scenario), nprocs=2, join=True)
        assert open(out).read() == 'ok', scenario
