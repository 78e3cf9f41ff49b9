import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from peppan_b200 import seqcodec, sw, workloads
from peppan_b200._lib import Context
ctx = Context(0)
npairs = int(sys.argv[1]); L = int(sys.argv[2])
rng = np.random.default_rng(1)
qs, ts = [], []
for p in range(npairs):
    m = int(rng.integers(L // 2, L))
    q = rng.integers(0, 4, m).astype(np.uint8)
    t = q.copy(); mask = rng.random(m) < 0.08; t[mask] = (t[mask] + 1) % 4
    t = np.concatenate([rng.integers(0, 4, 60).astype(np.uint8), t, rng.integers(0, 4, 60).astype(np.uint8)])
    qs.append(q); ts.append(t)
q, qoff = sw.concat(qs); t, toff = sw.concat(ts)
for coords in (False, True):
    job = sw.SwJob(ctx, q, qoff, t, toff, seqcodec.nt_params(), coords=coords)
    for _ in range(3):
        st = job.run()
    print('pairs', npairs, 'L', L, 'coords', coords, 'cells %.3g' % st['cells'], 'fwd %.2f ms rev %.2f ms' % (st['ms_forward'], st['ms_reverse']), 'GCUPS fwd %.0f' % (st['cells'] / st['ms_forward'] / 1e6))
