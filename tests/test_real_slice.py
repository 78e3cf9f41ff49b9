"""Real sequences: 164 CDS of one bundled E. coli genome against the syntenic 138 kb of another (tests/golden/real_slice.json.gz,
cut by tests/golden/make_real_slice.py).  The oracle half runs everywhere; the GPU half checks pb_search == oracle on these
real sequences, all three modes, with the oracle run live (the whole-genome version is tests/test_real_genomes.py)."""
import gzip
import json
import os

import numpy as np
import pytest

from peppan_b200 import seqcodec, seqio

HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    with gzip.open(os.path.join(HERE, 'golden', 'real_slice.json.gz'), 'rt') as f:
        d = json.load(f)
    qn, qb, qo = seqio.to_seqset([tuple(x) for x in d['queries']]); tn, tb, to = seqio.to_seqset([tuple(x) for x in d['target']])
    return qb, qo, tb, to


def test_oracle_finds_the_real_orthologs(oracle):
    qb, qo, tb, to = _load()
    n = len(qo) - 1
    for mode in (1, 2):
        hits, cig = oracle.search(qb, qo, tb, to, mode, seqcodec.BLOSUM62.reshape(-1), min_id=0.4, min_cov=50, min_ratio=0.25)
        span = (hits['q_end'] - hits['q_start'] + 1) / hits['q_len']
        full = set(hits['q_id'][(span >= 0.8) & (hits['identity'] >= 0.9)].tolist())
        assert len(full) >= 0.8 * n, (mode, len(full), n)      # the slice holds the syntenic region of ~85 % of the genes


@pytest.mark.gpu
@pytest.mark.parametrize('mode', [1, 2, 3])
def test_gpu_search_equals_oracle_on_real_sequences(ctx, oracle, mode):
    from peppan_b200 import search
    qb, qo, tb, to = _load()
    if mode == 3:
        tb, to = qb, qo
    hits, cigar, st = search.search(ctx, qb, qo, tb, to, mode, min_id=0.4, min_cov=50, min_ratio=0.25)
    ref, rcig = oracle.search(qb, qo, tb, to, mode, seqcodec.BLOSUM62.reshape(-1), min_id=0.4, min_cov=50, min_ratio=0.25)
    assert len(hits) == len(ref)
    for k in hits.dtype.names:
        if k in ('identity', 'evalue'):
            assert np.allclose(hits[k], ref[k], rtol=1e-6, atol=1e-30), k
        else:
            assert np.array_equal(hits[k], ref[k]), k
    assert np.array_equal(cigar, rcig)
