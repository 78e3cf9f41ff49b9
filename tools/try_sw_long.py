import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from peppan_b200 import seqcodec, sw
from peppan_b200._lib import Context
ctx = Context(0)
rng = np.random.default_rng(1)
def mk(npairs, lo, hi):
    qs, ts = [], []
    for p in range(npairs):
        m = int(rng.integers(lo, hi))
        q = rng.integers(0, 4, m).astype(np.uint8)
        t = q.copy(); mask = rng.random(m) < 0.08; t[mask] = (t[mask] + 1) % 4
        t = np.concatenate([rng.integers(0, 4, 40).astype(np.uint8), t, rng.integers(0, 4, 40).astype(np.uint8)])
        qs.append(q); ts.append(t)
    return qs, ts
for name, (n, lo, hi) in dict(long135=(135, 2500, 9500), one9k=(1, 9400, 9500), mid4000=(4000, 300, 2400), mix=(4500, 200, 2400)).items():
    qs, ts = mk(n, lo, hi)
    if name == 'mix':
        q2, t2 = mk(135, 2500, 9500); qs += q2; ts += t2
    q, qoff = sw.concat(qs); t, toff = sw.concat(ts)
    job = sw.SwJob(ctx, q, qoff, t, toff, seqcodec.nt_params(), coords=True)
    for _ in range(3):
        st = job.run()
    print(name, 'pairs', len(qs), 'cells %.3g' % st['cells'], 'fwd %.2f ms rev %.2f ms' % (st['ms_forward'], st['ms_reverse']))
