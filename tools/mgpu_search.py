"""Multi-GPU check (launch with torch.distributed.run): genomes sharded over ranks, per-rank
pb_search, one pb_allgather_hits; every rank must end with the same table, equal to the
concatenation (in rank order) of the single-GPU results."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch.distributed as dist
from peppan_b200 import dist as pbd, search, seqio, workloads
from peppan_b200._lib import Context

dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
ctx = pbd.init_context_from_env(dist)
pool = workloads.GenePool(40, 60, seed=workloads.SEED + 5)
genomes = [workloads.synth_genome(pool, g, n_acc_per_genome=20, seed=workloads.SEED + 5)[0] for g in range(2 * world + 1)]
qn, qb, qo = seqio.to_seqset(pool.fasta_items())


def run(c, idx, allgather):
    tn, tb, to = seqio.to_seqset([('g%d' % g, genomes[g]) for g in idx])
    return search.search(c, qb, qo, tb, to, search.MODE_PROT6, 0.4, 50, 0.25, allgather=allgather)


mine = pbd.shard_indices(len(genomes), rank, world)
hits, cig, st = run(ctx, mine, True)
ok = True
if rank == 0:
    solo = Context(int(os.environ.get('LOCAL_RANK', 0)))
    parts = []
    for r in range(world):
        h, c, _ = run(solo, pbd.shard_indices(len(genomes), r, world), False)
        parts.append((h, c))
    eh, ec, eoff = pbd.merge_hit_tables(parts)
    ok = len(eh) == len(hits) and all(np.array_equal(eh[k], hits[k]) for k in eh.dtype.names) and np.array_equal(ec, cig) \
        and np.array_equal(eoff, st['rank_offsets'])
    print('rank0: merged hits', len(hits), 'per rank', np.diff(st['rank_offsets']).tolist(), 'equal to concatenation:', ok)
# every rank holds the same table
import hashlib, torch
dig = hashlib.sha1(hits.tobytes() + cig.tobytes()).digest()[:8]
t = torch.tensor(list(dig), dtype=torch.int64)
ts = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(ts, t)
same = all(bool((x == ts[0]).all()) for x in ts)
# clustering split over the ranks (queries of both phases dealt round-robin, joined[] / edges exchanged over NCCL): every
# rank must end with the assignment a single GPU computes
from peppan_b200 import clust
rng = np.random.default_rng(11)
genes = []
for a in range(len(pool.genes)):
    for c in range(int(rng.integers(1, 5))):
        g = workloads._diverge(rng, pool.genes[a], float(rng.uniform(0.85, 1.0)))
        genes.append(workloads._NT[g].tobytes().decode())
genes.sort(key=lambda x: -len(x))
gn, gb, go = seqio.to_seqset([(str(i), x) for i, x in enumerate(genes)])
os.environ['PB_CLUSTER_BLOCK'] = '60000'          # several blocks, so both phases and the representative hand-over are exercised
rep, cst = clust.cluster(ctx, gb, go, 0.9, 0.8)
cl_ok = True
if rank == 0:
    rep1, _ = clust.cluster(solo, gb, go, 0.9, 0.8)
    cl_ok = bool(np.array_equal(rep, rep1)) and cst['n_blocks'] > 2
    print('rank0: clustering over', world, 'ranks:', int(cst['n_reps']), 'clusters in', int(cst['n_blocks']), 'blocks, equal to one GPU:', cl_ok)
dig2 = hashlib.sha1(rep.tobytes()).digest()[:8]
t2 = torch.tensor(list(dig2), dtype=torch.int64)
ts2 = [torch.zeros_like(t2) for _ in range(world)]
dist.all_gather(ts2, t2)
same = same and all(bool((x == ts2[0]).all()) for x in ts2)
flag = torch.tensor([1.0 if (ok and same and cl_ok) else 0.0]); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print('MGPU OK' if float(flag[0]) == 1.0 else 'MGPU FAILED', 'world', world)
dist.barrier()
