"""Single long pair: forward-pass time against the number of column blocks (wavefront hand-off cost)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from peppan_b200 import seqcodec, sw
from peppan_b200._lib import Context
ctx = Context(0)
rng = np.random.default_rng(1)
m = 9500
q = rng.integers(0, 4, m).astype(np.uint8)
for n in (2500, 9600):
    t = np.resize(q, n).copy(); mask = rng.random(n) < 0.08; t[mask] = (t[mask] + 1) % 4
    qq, qoff = sw.concat([q]); tt, toff = sw.concat([t])
    job = sw.SwJob(ctx, qq, qoff, tt, toff, seqcodec.nt_params(), coords=False)
    for _ in range(3):
        st = job.run()
    print('m', m, 'n', n, 'fwd %.3f ms' % st['ms_forward'], 'launches', st['kernel_launches'])
    job.close()
