"""Kernel-only timing of the config-2 micro-bench (device-resident job): python tools/bench_sw.py [pairs] [runs]
PB_SW_CFG=G,K,R,LONG picks one of the compiled forward / reverse kernel shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from peppan_b200 import seqcodec, sw, workloads
from peppan_b200._lib import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = Context(0)
q, qoff, t, toff = workloads.sw_microbench_pairs(n)
job = sw.SwJob(ctx, q, qoff, t, toff, seqcodec.protein_params(), coords=True)
for _ in range(3):
    job.run()
f = r = 0.0
for _ in range(runs):
    st = job.run(); f += st['ms_forward'] / runs; r += st['ms_reverse'] / runs
res = job.fetch()
cells = n * 300.0 * 300.0
print('cfg %s: forward %.2f ms (%.0f GCUPS), reverse %.2f ms, step %.0f GCUPS, checksum %d %d %d' %
      (os.environ.get('PB_SW_CFG', 'default'), f, cells / f / 1e6, r, cells / (f + r) / 1e6, int(res['score'].sum()), int(res['qs'].sum()), int(res['te'].sum())))
