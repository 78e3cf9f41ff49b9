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


def test_bsn_file_round_trips_the_pipeline_tables(tmp_path, oracle_as_search):
    """save_bsn / load_bsn (the flat replacement of `<prefix>.bsn.npz`, PEPPAN.py:866,924): the blastab uberBlast() returns for
    iter_map_bsn's flag set (names, ints, floats, CIGAR strings, merge-group lists) and its overlap table come back cell for
    cell with the same Python types."""
    import os
    from peppan_b200 import hitio, uberBlast as ub, workloads
    pool = workloads.GenePool(30, 30, seed=workloads.SEED + 9)
    seq, annot = workloads.synth_genome(pool, 0, n_acc_per_genome=15, seed=workloads.SEED + 9)
    ref, qry = os.path.join(tmp_path, 'g.fa'), os.path.join(tmp_path, 'q.fa')
    open(ref, 'w').write('>g0\n%s\n' % seq)
    open(qry, 'w').write(''.join('>%s\n%s\n' % x for x in pool.fasta_items()))
    tab, ovl = ub.uberBlast(['-r', ref, '-q', qry, '-f', '-m', '-O', '--blastn', '--diamond', '--min_id', '0.4', '--min_cov', '50', '--min_ratio', '0.25',
                             '-s', '1', '-e', '0,3'])
    assert tab.shape[0] > 30 and tab.shape[1] == 17
    path = os.path.join(tmp_path, 'x.bsn')
    hitio.save_bsn(path, tab, ovl)
    tab2, ovl2 = hitio.load_bsn(path)
    assert tab2.shape == tab.shape and np.array_equal(ovl2, ovl) and ovl2.dtype == np.int64
    for a, b in zip(tab.reshape(-1).tolist(), tab2.reshape(-1).tolist()):
        assert type(a) is type(b) and a == b
        if isinstance(a, list):
            assert [type(x) for x in a] == [type(x) for x in b]
    # raw (not rescored) tables carry integer scores next to float identities
    tab3 = ub.uberBlast(['-r', ref, '-q', qry, '--blastn', '--min_id', '0.4', '--min_cov', '50', '--min_ratio', '0.25'])
    hitio.save_bsn(path, tab3, np.zeros([0, 3], dtype=np.int64))
    tab4, ovl4 = hitio.load_bsn(path)
    assert ovl4.shape == (0, 3) and all(type(a) is type(b) and a == b for a, b in zip(tab3.reshape(-1).tolist(), tab4.reshape(-1).tolist()))
