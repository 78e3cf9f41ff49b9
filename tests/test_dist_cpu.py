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
