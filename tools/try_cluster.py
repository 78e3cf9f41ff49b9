"""Clustering at config-3-like shape: the genes of N synthetic genomes in priority (length) order through pb_cluster."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from peppan_b200 import workloads, seqio, clust
from peppan_b200._lib import Context
ngen = int(sys.argv[1]) if len(sys.argv) > 1 else 5
iden = float(sys.argv[2]) if len(sys.argv) > 2 else 0.9
t0 = time.time()
pool = workloads.GenePool(3000, 12000)
comp = str.maketrans('ACGT', 'TGCA')
genes = []
for g in range(ngen):
    seq, annot = workloads.synth_genome(pool, g)
    for (gid, s, e, strand, idn) in annot:
        x = seq[s:e]
        genes.append(x if strand > 0 else x.translate(comp)[::-1])
genes.sort(key=lambda x: -len(x))
print('generated %d genes of %d genomes in %.1f s' % (len(genes), ngen, time.time() - t0))
names, buf, off = seqio.to_seqset([(str(i), s) for i, s in enumerate(genes)])
ctx = Context(0)
for rep in range(2):
    t0 = time.time()
    r, st = clust.cluster(ctx, buf, off, iden, 0.8)
    dt = time.time() - t0
print('clusters', int((r == np.arange(len(r))).sum()), 'wall %.2f s' % dt, '%.0f genes/s' % (len(genes) / dt), st)
