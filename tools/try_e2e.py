import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from peppan_b200 import seqcodec, sw, workloads
from peppan_b200._lib import Context
ctx = Context(0)
n = 1000000
q0, qoff0, t0, toff0 = workloads.sw_microbench_pairs(n)
q = ctx.pinned_empty(q0.shape, np.uint8); q[:] = q0
t = ctx.pinned_empty(t0.shape, np.uint8); t[:] = t0
qoff = ctx.pinned_empty(qoff0.shape, np.int64); qoff[:] = qoff0
toff = ctx.pinned_empty(toff0.shape, np.int64); toff[:] = toff0
params = seqcodec.protein_params()
for _ in range(2): sw.sw_batch(ctx, q, qoff, t, toff, params)
ts = []
for _ in range(5):
    a = time.perf_counter(); out, st = sw.sw_batch(ctx, q, qoff, t, toff, params); ts.append(time.perf_counter() - a)
print('chunk', os.environ.get('PB_SW_CHUNK'), 'e2e ms %.1f' % (1e3 * np.median(ts)), 'device ms %.1f' % st['ms_total_device'], 'fwd %.1f rev %.1f' % (st['ms_forward'], st['ms_reverse']))
