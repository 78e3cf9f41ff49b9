"""Per-genome search at realistic size (config-3-like: 15k exemplar genes vs one ~4.8 Mbp genome)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from peppan_b200 import workloads, seqio, search
from peppan_b200._lib import Context
ncore = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
nacc = int(sys.argv[2]) if len(sys.argv) > 2 else 12000
modes = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [1, 2]
ngen = int(sys.argv[4]) if len(sys.argv) > 4 else 1
t0 = time.time()
pool = workloads.GenePool(ncore, nacc)
genomes = [workloads.synth_genome(pool, g, n_acc_per_genome=nacc // 8)[0] for g in range(ngen)]
print('generated in %.1f s: genome %d bp, %d genomes, exemplars %d' % (time.time() - t0, len(genomes[0]), ngen, ncore + nacc))
ctx = Context(0)
qn, qb, qo = seqio.to_seqset(pool.fasta_items())
rn, rb, ro = seqio.to_seqset([('g%d' % i, s) for i, s in enumerate(genomes)])
for mode in modes:
    for rep in range(2):
        t0 = time.time()
        hits, cigar, st = search.search(ctx, qb, qo, rb, ro, mode, 0.4, 50, 0.25)
        dt = time.time() - t0
    gbs = st['algo_bytes_seed'] / (st['ms_seed'] * 1e-3) / 1e9 if st['ms_seed'] > 0 else 0
    print('mode', mode, 'hits', len(hits), 'wall %.1f ms' % (dt * 1e3), 'seed %.2f ms (%.1f GB/s algorithmic)' % (st['ms_seed'], gbs), st)
