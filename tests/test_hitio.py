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


def _deep_equal(a, b):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        if not (isinstance(a, np.ndarray) and isinstance(b, np.ndarray)) or a.shape != b.shape or a.dtype != b.dtype:
            return False
        if a.dtype == object:
            return all(_deep_equal(x, y) for x, y in zip(a.reshape(-1), b.reshape(-1)))
        return np.array_equal(a, b)
    if isinstance(a, (list, tuple)):
        return type(a) is type(b) and len(a) == len(b) and all(_deep_equal(x, y) for x, y in zip(a, b))
    return type(a) is type(b) and a == b


def test_typed_codec_round_trips_nested_pipeline_values():
    rows = np.array([[1, 'a', 2.5, [1, 2.0, 'x'], np.arange(5, dtype=np.uint8), np.array([[1, 'q'], [2, 'r']], dtype=object)],
                     [np.int64(7), np.str_('b'), np.float64(1.5), [], np.zeros(0), None]], dtype=object)
    assert _deep_equal(hitio.decode_value(hitio.encode_value(rows)), rows)
    for v in (1 << 100, -3, 2.5, 'x', None, True, (1, 'a'), np.zeros([2, 0, 3]), np.array(['ab', 'c'])):
        assert _deep_equal(hitio.decode_value(hitio.encode_value(v)), v)
    with pytest.raises(TypeError):
        hitio.encode_value({'a': 1})


def test_flat_store_behaves_like_the_references_mapbsn(tmp_path):
    """FlatStore against the reference's own MapBsn (PEPPAN.py:27-114), operation for operation; and the reference's
    compare_prediction (PEPPAN.py:869-902) run on a FlatStore gives the table it gives on a MapBsn."""
    import sys
    ref = os.environ.get('PEPPAN_REFERENCE', '/root/reference')
    if not os.path.exists(os.path.join(ref, 'PEPPAN.py')):
        pytest.skip('reference checkout not present')
    from test_reference_consumer_cpu import PEPPAN as _fx
    P = _fx.__wrapped__() if hasattr(_fx, '__wrapped__') else None
    if P is None:
        pytest.skip('reference fixture not importable')
    if not hasattr(np.lib.npyio, 'format'):
        np.lib.npyio.format = np.lib.format
    rng = np.random.default_rng(5)

    def table(key, n):
        return np.array([[key, int(rng.integers(1, 1000)), int(rng.integers(1000, 2000)), '+-'[int(rng.integers(0, 2))], float(rng.random()),
                          np.arange(int(rng.integers(0, 4)), dtype=np.uint8)] for _ in range(n)], dtype=object)
    stores = (P.MapBsn(os.path.join(tmp_path, 'a.npz'), 'w'), hitio.FlatStore(os.path.join(tmp_path, 'b.npz'), 'w'))
    t1, t2, t3 = table(11, 3), table('g2', 2), table(13, 1)
    for s in stores:
        s._save(s.conn, '11', t1); s._save(s.conn, 'g2', t2)
        s.conn.close() if isinstance(s, P.MapBsn) else s.close()
    stores = (P.MapBsn(os.path.join(tmp_path, 'a.npz'), 'a'), hitio.FlatStore(os.path.join(tmp_path, 'b.npz'), 'a'))
    for s in stores:
        assert s.size() == 2 and s.exists(11) and s.exists('11') and not s.exists(12) and sorted(s.keys()) == ['11', 'g2']
        assert s.get('nope') == [] and s.get('nope', None) is None
    assert _deep_equal(stores[0].get(11), stores[1].get(11)) and _deep_equal(stores[0]['g2'], stores[1]['g2'])
    more = [np.array([[11, 5, 6, '+', 0.5, np.zeros(1, dtype=np.uint8)]], dtype=object), t3]
    for s in stores:
        s.update(more)
    for s in stores:
        assert sorted(s.keys()) == ['11', '13', 'g2'] and len(s.get(11)) == 4
    for k in ('11', '13', 'g2'):
        assert _deep_equal(stores[0].get(k), stores[1].get(k))
    assert _deep_equal(sorted(k for k, _ in stores[0].items()), sorted(k for k, _ in stores[1].items()))
    for s in stores:
        assert len(s.pop('g2')) == 2 and not s.exists('g2') and s.size() == 2
    stores[0].conn.close(); stores[1].close()

    # the reference's compare_prediction on both kinds of store
    old_rows = np.array([[7, 101, 400, '+'], [8, 900, 1500, '-']], dtype=object)
    blastab = np.array([[1, 500, 0.9, 300, 0, 0, 1, 300, 101, 400, 0.0, 290., 300, 5000, '300M', 0],
                        [2, 500, 0.8, 600, 0, 0, 1, 600, 1500, 901, 0.0, 500., 600, 5000, '600M', 1],
                        [3, 500, 0.7, 100, 0, 0, 1, 100, 3000, 3099, 0.0, 90., 100, 5000, '100M', 2]], dtype=object)
    outs = []
    for cls, name in ((P.MapBsn, 'c.npz'), (hitio.FlatStore, 'd.npz')):
        st = cls(os.path.join(tmp_path, name), 'w')
        st._save(st.conn, '500', old_rows)
        st.conn.close() if cls is P.MapBsn else st.close()
        saved = P.MapBsn
        P.MapBsn = cls
        try:
            outs.append(P.compare_prediction(blastab.copy(), os.path.join(tmp_path, name)))
        finally:
            P.MapBsn = saved
    assert outs[0].shape == outs[1].shape and all(_deep_equal(x, y) for x, y in zip(outs[0].reshape(-1).tolist(), outs[1].reshape(-1).tolist()))
    col10 = sorted(float(r[10]) for r in outs[1])
    assert col10[0] == 0.1 and col10[1] > 0.99 and col10[2] == 1.0


def test_typed_codec_stores_tables_column_by_column():
    """2-D object arrays of >= 8 rows are written column by column (typed columns, nested tables concatenated, typed arrays
    concatenated): the value that comes back has the same shape, cell types and values -- the layout of the per-genome
    `bsn` array of PEPPAN.iter_map_bsn (rows of [gene, contig, score, identity, encoded sequence, id, hit rows])."""
    rng = np.random.default_rng(11)
    rows = []
    for i in range(40):
        k = 1 + int(rng.integers(0, 3))
        hits = np.array([[int(rng.integers(1, 99)), 1001, float(rng.random()), 300, 2, 0, 1, 300, 5, 304, 0.1 if i % 3 else float(rng.random()),
                          float(rng.integers(100, 900)), 300, 5000, '%dM' % int(rng.integers(50, 300)), int(rng.integers(0, 9999))] for _ in range(k)], dtype=object)
        rows.append([int(rng.integers(1, 99)), 1001, np.float64(rng.random() * 100), float(rng.random()), rng.integers(0, 125, int(rng.integers(0, 40))).astype(np.uint8), i, hits])
    bsn = np.empty([len(rows), 7], dtype=object)
    for i, r in enumerate(rows):
        for j, v in enumerate(r):
            bsn[i, j] = v
    blob = hitio.encode_value(bsn)
    assert blob[0] == hitio._V_TABLE
    back = hitio.decode_value(blob)
    assert _deep_equal(back, bsn)
    assert type(back[0, 2]) is np.float64 and type(back[0, 3]) is float and type(back[0, 0]) is int and back[0, 4].dtype == np.uint8
    assert type(back[0, 6][0, 14]) is str and type(back[0, 6][0, 10]) is float
    # a column of mixed types falls back to cell-by-cell values; small arrays keep the generic form
    mixed = np.empty([9, 2], dtype=object)
    for i in range(9):
        mixed[i, 0] = i if i % 2 else float(i); mixed[i, 1] = None if i == 4 else 'x%d' % i
    assert _deep_equal(hitio.decode_value(hitio.encode_value(mixed)), mixed)
    small = bsn[:3].copy()
    assert hitio.encode_value(small)[0] == hitio._V_OBJ and _deep_equal(hitio.decode_value(hitio.encode_value(small)), small)
    empty_nested = np.empty([8, 1], dtype=object)
    for i in range(8):
        empty_nested[i, 0] = np.empty([0, 16], dtype=object)
    assert _deep_equal(hitio.decode_value(hitio.encode_value(empty_nested)), empty_nested)


def test_typed_codec_stores_vectors_of_tables_and_arrays_concatenated():
    """1-D object arrays of same-width object tables (a `.mat` chunk of get_map_bsn) or of typed arrays (a `.seq` chunk) are
    written as one concatenated table / array with the row counts; shapes, cell types and values come back."""
    rng = np.random.default_rng(12)
    mats = np.empty(30, dtype=object)
    for i in range(30):
        k = 1 + int(rng.integers(0, 3))
        mats[i] = np.array([[int(rng.integers(1, 99)), 1001, float(rng.random()), '%dM' % int(rng.integers(50, 300)), np.float64(rng.random())] for _ in range(k)], dtype=object)
    blob = hitio.encode_value(mats)
    assert blob[0] == hitio._V_COLUMN
    back = hitio.decode_value(blob)
    assert back.shape == (30,) and _deep_equal(back, mats)
    assert type(back[3][0, 0]) is int and type(back[3][0, 2]) is float and type(back[3][0, 3]) is str and type(back[3][0, 4]) is np.float64
    seqs = np.empty(12, dtype=object)
    for i in range(12):
        seqs[i] = rng.integers(0, 125, int(rng.integers(0, 50))).astype(np.uint8)
    back = hitio.decode_value(hitio.encode_value(seqs))
    assert hitio.encode_value(seqs)[0] == hitio._V_COLUMN and back.shape == (12,) and _deep_equal(back, seqs) and back[0].dtype == np.uint8
    ragged = np.empty(9, dtype=object)                       # tables of different widths keep the generic form
    for i in range(9):
        ragged[i] = np.empty([1, 2 + i % 2], dtype=object); ragged[i][:] = 1
    assert hitio.encode_value(ragged)[0] == hitio._V_OBJ and _deep_equal(hitio.decode_value(hitio.encode_value(ragged)), ragged)


def test_typed_codec_round_trips_random_nested_values():
    """property test (hypothesis): whatever nesting of the carried types -- scalars of Python and numpy, strings, big integers,
    lists / tuples, typed arrays, object vectors and object tables (below and above the column-wise threshold, with uniform and
    mixed columns, nested tables of one width, typed arrays) -- comes back with the same shapes, cell types and values."""
    hyp = pytest.importorskip('hypothesis')
    from hypothesis import given, settings, strategies as st, HealthCheck
    scalars = st.one_of(st.none(), st.booleans(), st.integers(-(1 << 70), 1 << 70), st.floats(allow_nan=False), st.text(max_size=8),
                        st.integers(-1000, 1000).map(np.int64), st.floats(allow_nan=False, width=32).map(np.float64),
                        st.text(alphabet='ACGT', max_size=6).map(np.str_), st.booleans().map(np.bool_))
    typed = st.one_of(st.lists(st.integers(0, 255), max_size=12).map(lambda v: np.array(v, dtype=np.uint8)),
                      st.lists(st.integers(-5, 5), min_size=0, max_size=6).map(lambda v: np.array(v, dtype=np.int64).reshape(-1, 1)),
                      st.lists(st.floats(allow_nan=False), max_size=5).map(lambda v: np.array(v, dtype=np.float64)),
                      st.lists(st.text(alphabet='MID0123456789', max_size=5), min_size=1, max_size=4).map(lambda v: np.array(v)))

    def obj_vector(items):
        a = np.empty(len(items), dtype=object)
        for i, v in enumerate(items):
            a[i] = v
        return a

    def obj_table(rows_cols):
        rows, cols, cells = rows_cols
        a = np.empty([rows, cols], dtype=object)
        for i in range(rows):
            for j in range(cols):
                a[i, j] = cells[(i * cols + j) % len(cells)] if cells else None
        return a

    leaf = st.one_of(scalars, typed)
    value = st.recursive(leaf, lambda inner: st.one_of(
        st.lists(inner, max_size=4), st.lists(inner, max_size=3).map(tuple), st.lists(inner, max_size=12).map(obj_vector),
        st.tuples(st.integers(0, 12), st.integers(1, 4), st.lists(inner, max_size=7)).map(obj_table),
        # a uniform column layout, as the pipeline's tables have: every row the same kinds
        st.tuples(st.integers(8, 14), st.lists(st.sampled_from(['i', 'f', 's', 'a', 't']), min_size=1, max_size=5), st.integers(0, 1 << 30)).map(
            lambda x: _uniform_table(*x))), max_leaves=12)

    @settings(max_examples=int(os.environ.get("PB_CODEC_EXAMPLES", 150)), deadline=None, derandomize="PB_CODEC_EXAMPLES" not in os.environ, suppress_health_check=list(HealthCheck))
    @given(value)
    def check(v):
        assert _deep_equal(hitio.decode_value(hitio.encode_value(v)), v)
    check()


def _uniform_table(rows, kinds, seed):
    rng = np.random.default_rng(seed)
    a = np.empty([rows, len(kinds)], dtype=object)
    for i in range(rows):
        for j, k in enumerate(kinds):
            if k == 'i':
                a[i, j] = int(rng.integers(-9, 9))
            elif k == 'f':
                a[i, j] = float(rng.random())
            elif k == 's':
                a[i, j] = '%dM' % int(rng.integers(1, 300))
            elif k == 'a':
                a[i, j] = rng.integers(0, 125, int(rng.integers(0, 9))).astype(np.uint8)
            else:
                t = np.empty([int(rng.integers(0, 3)), 3], dtype=object)
                for r in range(t.shape[0]):
                    t[r] = [int(rng.integers(0, 9)), float(rng.random()), 'x']
                a[i, j] = t
    return a
