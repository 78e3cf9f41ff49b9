"""BASELINE.json configs[3] ("config 4" of SURVEY.md 8d): N synthetic genomes searched against the exemplar (representative)
gene set -- nucleotide + protein-vs-6-frame -- sharded over the ranks of a torchrun launch (genome g -> rank g mod world, query
set replicated, no data-path collective), every per-genome hit table merged by pb_allgather_hits (NCCL).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/run_config4.py 1000
Genome synthesis (numpy, host) is outside the timed region; the timed region is the C-ABI calls with host buffers."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from peppan_b200 import dist as pbd, search, seqio, workloads

ngen = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
if world > 1:
    dist.init_process_group('gloo')
ctx = pbd.init_context_from_env(dist if world > 1 else None)
pool = workloads.GenePool(3000, 12000)
qn, qb, qo = seqio.to_seqset(pool.fasta_items())
mine = pbd.shard_indices(ngen, rank, world)
modes = (('nt', search.MODE_NT), ('prot6', search.MODE_PROT6))
t_gen = t_search = 0.0
dev_ms = {k: 0.0 for k, _ in modes}
hits_total = {k: 0 for k, _ in modes}
found = 0
# warm-up (kernels, allocator pool, NCCL connections)
seq, annot = workloads.synth_genome(pool, mine[0] if mine else 0)
rn, rb, ro = seqio.to_seqset([('w', seq)])
for _, m in modes:
    search.search(ctx, qb, qo, rb, ro, m, 0.4, 50, 0.25, allgather=world > 1)
steps = (ngen + world - 1) // world
for i in range(steps):
    g = mine[i] if i < len(mine) else None
    t0 = time.perf_counter()
    if g is not None:
        seq, annot = workloads.synth_genome(pool, g)
        rn, rb, ro = seqio.to_seqset([('g%d' % g, seq)])
    else:                                   # ragged tail: an empty target keeps the collective aligned
        rb, ro, annot = np.zeros(0, np.uint8), np.zeros(2, np.int64), []
    t_gen += time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k, m in modes:
        h, c, st = search.search(ctx, qb, qo, rb, ro, m, 0.4, 50, 0.25, allgather=world > 1)
        dev_ms[k] += st['ms_total']; hits_total[k] = hits_total[k] + len(h)
        if k == 'prot6' and g is not None:
            mineh = h if world == 1 else h[st['rank_offsets'][rank]:st['rank_offsets'][rank + 1]]
            present = set(a[0] for a in annot)
            found += len(present & set(int(x) for x in mineh['q_id'][(mineh['q_end'] - mineh['q_start'] + 1) >= 0.8 * mineh['q_len']]))
    t_search += time.perf_counter() - t0
t = torch.tensor([t_search, t_gen, float(found)], dtype=torch.float64)
tmax = t.clone()
if world > 1:
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
if rank == 0:
    nq = len(qo) - 1
    print(json.dumps({'config': '%d synthetic genomes (4,500 genes, ~5 Mbp each) vs %d exemplar genes, nt + protein 6-frame, %d GPU(s)' % (ngen, nq, world),
                      'search_seconds_max_over_ranks': float(tmax[0]), 'genome_synthesis_seconds_max_over_ranks': float(tmax[1]),
                      'genomes_per_s': ngen / float(tmax[0]), 'query_genes_per_s': ngen * nq / float(tmax[0]),
                      'merged_hits_seen_by_rank0': hits_total, 'device_ms_rank0': dev_ms,
                      'planted_genes_found_ge80pct_span_protein': int(t[2]), 'planted_genes': ngen * 4500}))
if world > 1:
    dist.barrier()
