"""Columnar hit-table file: round trip, mapped views, object rows."""
import os

import numpy as np
import pytest

from peppan_b200 import hitio
from peppan_b200.search import HIT_DTYPE


def _table(n, seed=0):
    rng = np.random.default_rng(seed)
    h = np.zeros(n, dtype=HIT_DTYPE)
    for k in ('q_id', 's_id'):
        h[k] = rng.integers(0, 5, n)
    for k in ('q_start', 'q_end', 's_start', 's_end', 'aln_len', 'mismatch', 'gapopen', 'raw_score', 'q_len', 's_len', 'frame'):
        h[k] = rng.integers(0, 5000, n)
    h['identity'] = rng.random(n).astype(np.float32); h['evalue'] = 1e-30
    lens = rng.integers(1, 6, n)
    h['cigar_n'] = lens; h['cigar_off'] = np.concatenate([[0], np.cumsum(lens)[:-1]]) if n else []
    cig = ((rng.integers(1, 900, int(lens.sum())) << 2) | rng.integers(0, 3, int(lens.sum()))).astype(np.uint32)
    return h, cig


@pytest.mark.parametrize('n', [0, 1, 257])
def test_round_trip(tmp_path, n):
    h, c = _table(n)
    qn = ['%d' % (100 + i) for i in range(5)]; sn = ['contig_%d' % i for i in range(5)]
    p = os.path.join(tmp_path, 't.pbh')
    hitio.save_hits(p, h, c, qn, sn)
    assert os.path.getsize(p) == 56 + n * 68 + len(c) * 4 + len('\n'.join(qn)) + len('\n'.join(sn))
    for mm in (True, False):
        h2, c2, q2, s2 = hitio.load_hits(p, mmap=mm)
        assert np.array_equal(h2, h) and np.array_equal(c2, c) and q2 == qn and s2 == sn
    rows = hitio.to_object_rows(h, c, qn, sn)
    assert rows.shape == (n, 15)
    for i in range(n):
        assert rows[i, 0] == qn[h['q_id'][i]] and rows[i, 11] == int(h['raw_score'][i]) and isinstance(rows[i, 11], int)
        assert sum(k for k, _ in rows[i, 14]) == int((c[h['cigar_off'][i]:h['cigar_off'][i] + h['cigar_n'][i]] >> 2).sum())
        assert all(o in 'MID' for _, o in rows[i, 14])


def test_rejects_foreign_files(tmp_path):
    p = os.path.join(tmp_path, 'x.bin')
    open(p, 'wb').write(b'not a hit table at all, definitely' * 4)
    with pytest.raises(ValueError):
        hitio.load_hits(p)
    with pytest.raises(ValueError):
        hitio.save_hits(p, *_table(1), ['a\nb'], ['c'])
