"""Traceback timing on nucleotide boxes shaped like a per-genome search (incl. a few ~9 kb genes)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from peppan_b200 import seqcodec, sw
from peppan_b200._lib import Context
ctx = Context(0)
rng = np.random.default_rng(1)
def mk(npairs, lo, hi, iden=0.92):
    qs, ts = [], []
    for p in range(npairs):
        m = int(rng.integers(lo, hi))
        q = rng.integers(0, 4, m).astype(np.uint8)
        t = q.copy(); mask = rng.random(m) > iden; t[mask] = (t[mask] + 1) % 4
        t = np.concatenate([rng.integers(0, 4, 40).astype(np.uint8), t, rng.integers(0, 4, 40).astype(np.uint8)])
        qs.append(q); ts.append(t)
    return qs, ts
for name, (n, lo, hi) in dict(one9k=(1, 9400, 9500), mid4000=(4000, 300, 2400), mix=(4500, 200, 2400), bulk=(20000, 500, 1000), bulk2k=(6000, 1500, 2400)).items():
    qs, ts = mk(n, lo, hi)
    if name == 'mix':
        q2, t2 = mk(25, 2500, 9500); qs += q2; ts += t2
    q, qoff = sw.concat(qs); t, toff = sw.concat(ts)
    for _ in range(2):
        t0 = time.time()
        out, st = sw.sw_align_batch(ctx, q, qoff, t, toff, seqcodec.nt_params())
        wall = time.time() - t0
    box = float(((out['qe'] - out['qs'] + 1).astype(np.float64) * (out['te'] - out['ts'] + 1)).sum())
    print(name, 'pairs', len(qs), 'box cells %.3g' % box, 'fwd %.2f rev %.2f trace %.2f ms (%.0f GCUPS) wall %.1f ms' % (
        st['ms_forward'], st['ms_reverse'], st['ms_traceback'], box / max(st['ms_traceback'], 1e-9) / 1e6, wall * 1e3))
