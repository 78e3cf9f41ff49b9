"""Profiling driver: one pb_cluster call on the genes of G synthetic genomes.  python tools/prof_cluster.py [genomes] [identity]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from peppan_b200 import workloads, clust
from peppan_b200._lib import Context
ngen = int(sys.argv[1]) if len(sys.argv) > 1 else 20
iden = float(sys.argv[2]) if len(sys.argv) > 2 else 0.9
made = workloads.synth_genomes_parallel(range(ngen), procs=min(16, os.cpu_count() or 1))
comp = bytes.maketrans(b'ACGT', b'TGCA')
genes = []
for _, seq, annot in made:
    sb = seq.tobytes()
    for gid, a, b, strand in annot.tolist():
        x = sb[a:b]
        genes.append(x if strand > 0 else x.translate(comp)[::-1])
genes.sort(key=lambda x: -len(x))
buf = np.frombuffer(b''.join(genes), dtype=np.uint8)
off = np.zeros(len(genes) + 1, np.int64); off[1:] = np.cumsum([len(g) for g in genes])
ctx = Context(0)
clust.cluster(ctx, buf[:off[2000]], off[:2001], iden, 0.8)
t0 = time.time()
rep, st = clust.cluster(ctx, buf, off, iden, 0.8)
dt = time.time() - t0
print('genes', len(genes), 'identity', iden, 'seconds %.2f' % dt, 'gcups %.0f' % (st['sw_cells'] / dt / 1e9), st)
