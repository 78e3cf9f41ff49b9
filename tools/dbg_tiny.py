import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import numpy as np
from peppan_b200 import seqcodec, sw
from peppan_b200._lib import Context
import pb_oracle
ctx = Context(0)
mat = seqcodec.protein_matrix().reshape(-1)
for qs, ts in [([np.array([17], np.uint8)], [np.array([17], np.uint8)]),
               ([np.array([17, 3], np.uint8)], [np.array([17, 3], np.uint8)]),
               ([np.array([0, 1, 2, 3, 4], np.uint8)], [np.array([9, 0, 1, 2, 3, 4, 9], np.uint8)]),
               ([np.arange(20, dtype=np.uint8)] * 2, [np.arange(20, dtype=np.uint8), np.arange(20, dtype=np.uint8)[::-1].copy()])]:
    q, qoff = sw.concat(qs); t, toff = sw.concat(ts)
    out, st = sw.sw_batch(ctx, q, qoff, t, toff, seqcodec.protein_params(), coords=False)
    ref, _ = pb_oracle.sw_batch(q, qoff, t, toff, mat, 11, 1, with_cigar=False)
    print('gpu', out['score'], out['qe'], out['te'], 'ref', ref['score'], ref['qe'], ref['te'])
