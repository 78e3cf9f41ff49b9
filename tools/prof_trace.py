"""Profiling driver: one grouped search of G synthetic genomes (nucleotide and / or protein mode) -- the batch shape of the
bench's config-4 leg.  python tools/prof_trace.py [genomes] [modes]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from peppan_b200 import workloads, seqio, search
from peppan_b200._lib import Context
ngen = int(sys.argv[1]) if len(sys.argv) > 1 else 16
modes = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
made = workloads.synth_genomes_parallel(range(ngen), procs=min(16, os.cpu_count() or 1))
pool = workloads.GenePool(3000, 12000)
ctx = Context(0)
qn, qb, qo = seqio.to_seqset(pool.fasta_items())
tb = np.concatenate([g[1] for g in made]); to = np.zeros(ngen + 1, np.int64); to[1:] = np.cumsum([len(g[1]) for g in made])
groups = np.arange(ngen, dtype=np.int32)
for mode in modes:
    for rep in range(reps):
        t0 = time.time()
        hits, cig, goff, st = search.search_grouped_raw(ctx, qb, qo, tb, to, groups, mode, min_id=0.4, min_cov=50, min_ratio=0.25)
        dt = time.time() - t0
    print('mode', mode, 'genomes', ngen, 'hits', len(hits), 'wall %.1f ms = %.2f ms / genome' % (dt * 1e3, dt * 1e3 / ngen),
          {k: (round(v / ngen, 3) if k.startswith('ms_') else v) for k, v in st.items()})
