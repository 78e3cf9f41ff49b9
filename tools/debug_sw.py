import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import numpy as np
from peppan_b200 import seqcodec, sw, workloads
from peppan_b200._lib import Context
import pb_oracle

ctx = Context(0)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
maxlen = int(sys.argv[3]) if len(sys.argv) > 3 else 400
qs, ts = workloads.random_pairs(n, seed=seed, max_len=maxlen)
q, qoff = sw.concat(qs); t, toff = sw.concat(ts)
mat = seqcodec.protein_matrix().reshape(-1)
ref, _ = pb_oracle.sw_batch(q, qoff, t, toff, mat, 11, 1, with_cigar=False, nthreads=8)
out, st = sw.sw_batch(ctx, q, qoff, t, toff, seqcodec.protein_params(), coords=False)
for k in ('score', 'qe', 'te'):
    bad = np.nonzero(out[k] != ref[k])[0]
    print('fwd', k, 'mismatches', len(bad), 'of', n)
    for b in bad[:8]:
        print('   pair', b, 'm', len(qs[b]), 'n', len(ts[b]), 'gpu', out['score'][b], out['qe'][b], out['te'][b], 'ref', ref['score'][b], ref['qe'][b], ref['te'][b])
try:
    out, st = sw.sw_batch(ctx, q, qoff, t, toff, seqcodec.protein_params(), coords=True)
    for k in ('score', 'qe', 'te', 'qs', 'ts'):
        bad = np.nonzero(out[k] != ref[k])[0]
        print('full', k, 'mismatches', len(bad), 'of', n)
        for b in bad[:8]:
            print('   pair', b, 'm', len(qs[b]), 'n', len(ts[b]), 'gpu', [out[x][b] for x in ('score','qs','qe','ts','te')], 'ref', [ref[x][b] for x in ('score','qs','qe','ts','te')])
    print(st)
except Exception as e:
    print('full failed:', e)
