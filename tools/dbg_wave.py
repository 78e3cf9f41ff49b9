import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import numpy as np
from peppan_b200 import seqcodec, sw, workloads
from peppan_b200._lib import Context
import pb_oracle
ctx = Context(0)
def run(qs, ts, params, mat, go, ge, tag):
    q, qoff = sw.concat(qs); t, toff = sw.concat(ts)
    ref, _ = pb_oracle.sw_batch(q, qoff, t, toff, mat, go, ge, with_cigar=False, nthreads=8)
    out, st = sw.sw_batch(ctx, q, qoff, t, toff, params, coords=False)
    for p in range(len(qs)):
        print(tag, 'fwd pair', p, 'm', len(qs[p]), 'n', len(ts[p]), 'gpu', out['score'][p], out['qe'][p], out['te'][p], 'ref', ref['score'][p], ref['qe'][p], ref['te'][p])
    try:
        out, st = sw.sw_batch(ctx, q, qoff, t, toff, params, coords=True)
        for p in range(len(qs)):
            print(tag, 'full pair', p, 'gpu', [out[k][p] for k in ('score','qs','qe','ts','te')], 'ref', [ref[k][p] for k in ('score','qs','qe','ts','te')])
    except Exception as e:
        print(tag, 'full failed', e)
qs, ts = workloads.random_pairs(3, seed=31, nsym_real=4, min_len=2600, max_len=4000, related=1.0)
run(qs, ts, seqcodec.nt_params(), seqcodec.nt_matrix().reshape(-1), 6, 2, 'nt16')
rng = np.random.default_rng(5)
big = rng.integers(0, 20, 5000).astype(np.uint8); big[::2] = 17
run([big, big[:4100]], [big.copy(), big.copy()], seqcodec.protein_params(), seqcodec.protein_matrix().reshape(-1), 11, 1, 'aa32')
