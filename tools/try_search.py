import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from peppan_b200 import workloads, seqio, search
from peppan_b200._lib import Context
ncore = int(sys.argv[1]) if len(sys.argv) > 1 else 300
nacc = int(sys.argv[2]) if len(sys.argv) > 2 else 600
pool = workloads.GenePool(ncore, nacc)
seq, annot = workloads.synth_genome(pool, 0, n_acc_per_genome=nacc // 4)
print('genome', len(seq), 'genes', len(annot))
ctx = Context(0)
qn, qb, qo = seqio.to_seqset(pool.fasta_items())
rn, rb, ro = seqio.to_seqset([('ctg', seq)])
for mode in (1, 2):
    for rep in range(2):
        t0 = time.time()
        hits, cigar, st = search.search(ctx, qb, qo, rb, ro, mode, 0.4, 50, 0.25)
        dt = time.time() - t0
    print('mode', mode, 'hits', len(hits), 'wall %.1f ms' % (dt * 1e3), st)
    found = set(int(h['q_id']) for h in hits if h['q_end'] - h['q_start'] + 1 >= 0.8 * h['q_len'])
    present = set(a[0] for a in annot)
    print('   genes present', len(present), 'found (>=80% span)', len(found & present), 'false', len(found - present))
    print(hits[:3])
