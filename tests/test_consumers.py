"""N4: peppan_b200.consumers.map_bsn_groups against the REFERENCE'S OWN iter_map_bsn (PEPPAN.py:759-867) on the
same blastab: the reference function runs on the uberBlast shim's output (module swap, search answered by the oracle as in
tests/test_reference_consumer_cpu.py); the table compare_prediction hands to the grouping / scoring loop is captured and
given to the columnar implementation, whose groups, scores, encoded sequences and overlap table must equal what the
reference saved."""
import os

import numpy as np
import pytest

from peppan_b200 import consumers, uberBlast as ub, workloads
from test_reference_consumer_cpu import PEPPAN, REF  # noqa: F401  (fixture)

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'PEPPAN.py')), reason='reference checkout not present')


def _same(a, b):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        a, b = np.asarray(a), np.asarray(b)
        if a.shape != b.shape:
            return False
        if a.dtype == object or b.dtype == object:
            return all(_same(x, y) for x, y in zip(a.reshape(-1), b.reshape(-1)))
        return a.dtype == b.dtype and np.array_equal(a, b)
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return a == b


@pytest.mark.parametrize('seed,identity', [(41, 0.5), (42, 0.9)])
def test_columnar_grouping_and_scoring_equals_the_reference_loop(PEPPAN, oracle_as_search, monkeypatch, tmp_path, seed, identity):
    monkeypatch.setattr(PEPPAN, 'uberBlast', ub.uberBlast)
    if not hasattr(np.lib.npyio, 'format'):
        monkeypatch.setattr(np.lib.npyio, 'format', np.lib.format, raising=False)
    pool = workloads.GenePool(50, 50, seed=workloads.SEED + seed)
    seq, annot = workloads.synth_genome(pool, 0, n_acc_per_genome=25, seed=workloads.SEED + seed)
    # cut the genome inside a gene so that the search returns fragments that linearMerge joins (merge groups of > 1 hit)
    mid = annot[len(annot) // 2]
    cut = (int(mid[1]) + int(mid[2])) // 2
    contigs = [(1001, seq[:cut]), (1002, seq[cut:])]
    clust = os.path.join(tmp_path, 'exemplar.fa')
    with open(clust, 'w') as f:
        for n, s in pool.fasta_items():
            f.write('>%s\n%s\n' % (n, s))
    old = os.path.join(tmp_path, 'old.npz')
    store = PEPPAN.MapBsn(old, 'w')
    rows = [[a[0], a[1] + 1, a[2], '+' if a[3] > 0 else '-'] for a in annot if a[2] <= cut]
    store._save(store.conn, '1001', np.array(sorted(rows, key=lambda r: r[1]), dtype=object))
    store.conn.close()
    ortho = os.path.join(tmp_path, 'ortho.npy')
    genes = sorted(set(int(a[0]) for a in annot))
    pairs = np.array([[genes[i], genes[i + 1], (1 if i % 2 else -1) * 9000] for i in range(0, min(len(genes) - 1, 30))], dtype=int)
    np.save(ortho, pairs, allow_pickle=True)
    params = dict(gtable=11, noDiamond=False, match_identity=identity, match_frag_len=50., match_frag_prop=0.25, link_gap=600., link_diff=1.5,
                  match_prop=0.5, match_len=250., match_prop1=0.8, match_len1=100., match_prop2=0.4, match_len2=400.)
    captured = {}
    real_ub, real_cp = ub.uberBlast, PEPPAN.compare_prediction

    def spy_ub(argv, *a, **k):
        tab, ovl = real_ub(argv, *a, **k)
        captured['ovl'] = ovl.copy()
        return tab, ovl

    def spy_cp(blastab, old_prediction):
        out = real_cp(blastab, old_prediction)
        captured['blastab'] = np.array([[(list(c) if isinstance(c, list) else c) for c in r] for r in out], dtype=object)
        return out

    monkeypatch.setattr(PEPPAN, 'uberBlast', spy_ub)
    monkeypatch.setattr(PEPPAN, 'compare_prediction', spy_cp)
    prefix = os.path.join(tmp_path, 'run')
    out = PEPPAN.iter_map_bsn((prefix, clust, 0, 'taxon', contigs, ortho, old, params))
    res = np.load(out + '.bsn.npz', allow_pickle=True)
    ref_bsn, ref_ovl = res['bsn'], res['ovl']

    bsn, ovl = consumers.map_bsn_groups(captured['blastab'], captured['ovl'], contigs, params, np.load(ortho, allow_pickle=True))
    assert bsn.shape == ref_bsn.shape and len(bsn) >= 40
    assert any(len(g[6]) > 1 for g in ref_bsn) or identity > 0.8        # merge groups are exercised at the permissive setting
    for g, r in zip(bsn, ref_bsn):
        assert g[0] == r[0] and g[1] == r[1] and g[5] == r[5]
        assert g[2] == r[2] and g[3] == r[3], (g[:4], r[:4])            # group score: same floating-point operations, same value
        assert g[4].dtype == np.uint8 and np.array_equal(g[4], r[4])
        assert _same(g[6], r[6])
    assert ovl.shape == ref_ovl.shape and np.array_equal(ovl, ref_ovl)


def test_score_hits_handles_gaps_frames_and_stops():
    # one hit with an insertion and a deletion on the minus strand, scored by hand
    seq = [(7, 'AAACCCGGGTTTACGTACGTTAGCCCAAATTT')]
    #        1-based 4..21 on the minus strand: rc('CCCGGGTTTACGTACGTT') = 'AACGTACGTAAACCCGGG'
    row = [1, 7, 1.0, 0, 0, 0, 1, 19, 21, 4, 1.0, 0, 30, 32, '6M2I3M1D8M', 0]
    sc, enc = consumers.score_hits([row], seq)
    ms = 'AACGTA' + '--' + 'CGT' + 'AACCCGGG'        # the deletion skips one subject base ('A')
    assert ''.join(' ACGT'[v] if v else '-' for v in enc[0]) == ms
    # frames: 6 M in frame 0, +2 -> frame 2: 3 M, -1 -> frame 1: 8 M ; codons of ms: AAC GTA --C GTA ACC CGG (no stop) -> 19
    assert int(sc[0]) == min(8, len(ms) + 3)
    row2 = [1, 7, 1.0, 0, 0, 0, 1, 12, 10, 21, 1.0, 0, 30, 32, '12M', 0]      # plus strand: 'TTACGTACGTTA' holds TTA CGT ACG TTA: no stop
    sc2, enc2 = consumers.score_hits([row2], seq)
    assert int(sc2[0]) == 12 and len(enc2[0]) == 12
    row3 = [1, 7, 1.0, 0, 0, 0, 1, 12, 11, 22, 1.0, 0, 30, 32, '12M', 0]      # 'TACGTACGTTAG': TAC GTA CGT TAG -> stop at 9: max(9, 3) + 3
    sc3, _ = consumers.score_hits([row3], seq)
    assert int(sc3[0]) == 12
    row4 = [1, 7, 1.0, 0, 0, 0, 1, 15, 20, 34 - 2, 1.0, 0, 30, 32, '13M', 0]
    seq4 = [(7, 'AAACCCGGGTTTACGTACGTAGCCCAAATTTGG')]                          # 20..32: 'TAGCCCAAATTTG': stop first -> spans 0 and 13
    sc4, _ = consumers.score_hits([row4], seq4)
    assert int(sc4[0]) == 13


def test_get_similar_pairs_equals_the_reference(PEPPAN, oracle_as_search, monkeypatch, tmp_path):
    """consumers.get_similar_pairs against PEPPAN.get_similar_pairs (PEPPAN.py:194-294) on the same exemplar file: the same
    ortholog pairs with the same average identities, the same exemplar file afterwards, the same merge records."""
    import shutil
    monkeypatch.setattr(PEPPAN, 'uberBlast', ub.uberBlast)
    monkeypatch.setattr(PEPPAN, 'pool', None, raising=False)
    rng = np.random.default_rng(78)
    gp = workloads.GenePool(36, 0, seed=workloads.SEED + 43)
    allg = {i: g for i, g in enumerate(gp.genes)}
    allg.update({100 + i: workloads._diverge(rng, gp.genes[i], 0.95) for i in range(0, 10)})          # merged into their exemplar
    allg.update({200 + i: workloads._diverge(rng, gp.genes[i], 0.75) for i in range(10, 22)})         # ortholog pairs
    allg.update({300 + i: workloads._diverge(rng, gp.genes[i], 0.62) for i in range(22, 30)})         # weak pairs: mostly protein hits
    allg.update({400 + i: gp.genes[i][:len(gp.genes[i]) // 6 * 3] for i in range(30, 34)})            # half-length fragments
    out = []
    for tag in ('ref', 'ours'):
        d = os.path.join(tmp_path, tag); os.makedirs(d)
        clust = os.path.join(d, 'x.clust.exemplar')
        with open(clust, 'w') as f:
            for n, g in allg.items():
                f.write('>%d\n%s\n' % (n, workloads._NT[g].tobytes().decode()))
        np.save(os.path.join(d, 'x.clust.npy'), np.zeros([0, 3], dtype=int))
        priorities = {n: [n % 3, n] for n in allg}
        params = dict(clust=clust, incompleteCDS='', noDiamond=False, n_thread=2, gtable=11, match_identity=0.5, match_frag_len=50., match_frag_prop=0.25,
                      match_prop=0.5, match_len=250., match_prop1=0.8, match_len1=100., match_prop2=0.4, match_len2=400., clust_identity=0.9, clust_match_prop=0.8)
        pairs = PEPPAN.get_similar_pairs(clust, priorities, params) if tag == 'ref' else consumers.get_similar_pairs(clust, priorities, params)
        out.append((pairs, open(clust).read(), np.load(os.path.join(d, 'x.clust.npy'), allow_pickle=True)))
    (p0, f0, c0), (p1, f1, c1) = out
    assert p0.shape == p1.shape and len(p0) >= 12 and np.array_equal(p0, p1)
    assert f0 == f1 and c0.shape == c1.shape and len(c0) >= 8 and np.array_equal(c0.astype(int), c1.astype(int))


def test_iter_map_bsn_writes_what_the_reference_writes(PEPPAN, oracle_as_search, monkeypatch, tmp_path):
    """consumers.iter_map_bsn (search -> compare_prediction -> grouping / scoring -> .bsn.npz) against PEPPAN.iter_map_bsn on
    the same genome, exemplars, old predictions and ortholog pairs: the saved arrays are equal cell for cell."""
    monkeypatch.setattr(PEPPAN, 'uberBlast', ub.uberBlast)
    if not hasattr(np.lib.npyio, 'format'):
        monkeypatch.setattr(np.lib.npyio, 'format', np.lib.format, raising=False)
    pool = workloads.GenePool(50, 50, seed=workloads.SEED + 47)
    seq, annot = workloads.synth_genome(pool, 0, n_acc_per_genome=25, seed=workloads.SEED + 47)
    mid = annot[len(annot) // 3]
    cut = (int(mid[1]) + int(mid[2])) // 2
    # the second contig in lower case: the search upper-cases what it reads, the scoring loop sees the raw string (plus-strand hits
    # encode to zeros, minus-strand hits go through rc(), which upper-cases: modules/configure.py:153-154)
    contigs = [(1001, seq[:cut]), (1002, seq[cut:].lower())]
    clust = os.path.join(tmp_path, 'exemplar.fa')
    with open(clust, 'w') as f:
        for n, s in pool.fasta_items():
            f.write('>%s\n%s\n' % (n, s))
    # old predictions on both contigs, both strands, some shifted so that the frame test and the 60 % rule both matter
    old = os.path.join(tmp_path, 'old.npz')
    store = PEPPAN.MapBsn(old, 'w')
    r1 = [[a[0], a[1] + 1 + (3 if k % 4 == 0 else 0) + (1 if k % 7 == 0 else 0), a[2] + (90 if k % 5 == 0 else 0), '+' if a[3] > 0 else '-']
          for k, a in enumerate(annot) if a[2] <= cut]
    r2 = [[a[0], a[1] + 1 - cut, a[2] - cut, '+' if a[3] > 0 else '-'] for k, a in enumerate(annot) if a[1] >= cut and k % 2 == 0]
    store._save(store.conn, '1001', np.array(sorted(r1, key=lambda r: r[1]), dtype=object))
    store._save(store.conn, '1002', np.array(sorted(r2, key=lambda r: r[1]), dtype=object))
    store.conn.close()
    ortho = os.path.join(tmp_path, 'ortho.npy')
    genes = sorted(set(int(a[0]) for a in annot))
    np.save(ortho, np.array([[genes[i], genes[i + 1], (1 if i % 2 else -1) * 9000] for i in range(0, 30)], dtype=int), allow_pickle=True)
    params = dict(gtable=11, noDiamond=False, match_identity=0.5, match_frag_len=50., match_frag_prop=0.25, link_gap=600., link_diff=1.5,
                  match_prop=0.5, match_len=250., match_prop1=0.8, match_len1=100., match_prop2=0.4, match_len2=400.)
    a = PEPPAN.iter_map_bsn((os.path.join(tmp_path, 'ref'), clust, 0, 'taxon', contigs, ortho, old, params))
    b = consumers.iter_map_bsn((os.path.join(tmp_path, 'ours'), clust, 0, 'taxon', contigs, ortho, old, params), store=PEPPAN.MapBsn)
    ra, rb = np.load(a + '.bsn.npz', allow_pickle=True), np.load(b + '.bsn.npz', allow_pickle=True)
    assert ra['bsn'].shape == rb['bsn'].shape and len(ra['bsn']) >= 40
    assert _same(rb['bsn'], ra['bsn'])
    assert ra['ovl'].shape == rb['ovl'].shape and np.array_equal(ra['ovl'], rb['ovl'])
    col10 = [float(t[10]) for g in rb['bsn'] for t in g[6]]
    assert sum(1 for v in col10 if v > 0.1) >= 20 and sum(1 for v in col10 if 0.1 < v < 0.99) >= 3      # overlap fractions were exercised


class _SerialPool(object):
    """stands in for the multiprocessing pool behind PEPPAN.get_map_bsn (`pool.imap_unordered`, PEPPAN.py:924)"""
    def imap_unordered(self, fn, tasks):
        return map(fn, tasks)

    def close(self):
        pass

    def join(self):
        pass


def _stores(cls, d, mode='w'):
    return [cls(os.path.join(d, name), mode) for name in ('tab.npz', 'seq.npz', 'mat.npz', 'clf.npz')]


def _store_equal(a, b):
    if sorted(a.keys()) != sorted(b.keys()):
        return False
    return all(_same(a.get(k), b.get(k)) for k in a.keys())


def _synthetic_result(rng, gid, n_groups):
    """a (bsn, ovl) pair shaped like iter_map_bsn's output: rows [gene, contig, score, identity, encoded sequence, id, hit rows]"""
    bsn = np.empty([n_groups, 7], dtype=object)
    for i in range(n_groups):
        k = 1 + int(rng.integers(0, 3))
        hits = np.array([[int(rng.integers(1, 60)), 5000 + gid, float(rng.random()), 300, 2, 0, 1, 300, 5, 304, 0.1, float(rng.integers(100, 900)), 300, 5000,
                          '%dM' % int(rng.integers(50, 300)), int(rng.integers(0, 9999))] for _ in range(k)], dtype=object)
        # coarse scores so that equal scores occur (the order among ties is the reference's argsort of an object column)
        bsn[i] = [int(rng.integers(1, 60)), 5000 + gid, np.float64(int(rng.integers(1, 40)) * 12.3456789), float(int(rng.integers(50, 100)) / 100. + 1e-5),
                  rng.integers(0, 125, int(rng.integers(1, 30))).astype(np.uint8), i, hits]
    m = int(rng.integers(0, 3 * n_groups)) if gid % 7 else 0
    ovl = np.stack([rng.integers(0, n_groups, m), rng.integers(0, n_groups, m), rng.integers(0, 3, m)], axis=1).astype(np.int64).reshape(-1, 3)
    return bsn, ovl


@pytest.mark.parametrize('n_genomes,n_groups,save_seq', [(3, 40, True), (7, 450, False), (503, 61, True)])
def test_bsn_merger_fills_the_stores_like_the_references_get_map_bsn(PEPPAN, monkeypatch, tmp_path, n_genomes, n_groups, save_seq):
    """consumers.get_map_bsn / BsnMerger against PEPPAN.get_map_bsn (:907-983) on the same per-genome results: the four stores
    (integer hit table per gene, sequence chunks, hit-row chunks, conflict lists per 30,000 group ids) hold equal values.
    503 genomes x 61 groups cross the 500-genome flush of the table, the 1,000-value chunks and the 30,000-id bucket."""
    if not hasattr(np.lib.npyio, 'format'):
        monkeypatch.setattr(np.lib.npyio, 'format', np.lib.format, raising=False)
    from peppan_b200 import hitio
    rng = np.random.default_rng(5 + n_genomes)
    results = [_synthetic_result(rng, g, n_groups + (g % 3)) for g in range(n_genomes)]
    genomes = {5000 + g: [700 + g, 'ACGT'] for g in range(n_genomes)}

    def ref_task(data):
        prefix, gid = data[0], data[2]
        bsn, ovl = results[gid]
        np.savez_compressed('%s.%d.bsn.npz' % (prefix, gid), bsn=bsn.copy(), ovl=ovl.copy())
        return '%s.%d' % (prefix, gid)

    monkeypatch.setattr(PEPPAN, 'pool', _SerialPool(), raising=False)
    monkeypatch.setattr(PEPPAN, 'iter_map_bsn', ref_task)
    monkeypatch.setattr(PEPPAN, 'logger', lambda *a, **k: None)
    dr, do = os.path.join(tmp_path, 'ref'), os.path.join(tmp_path, 'ours')
    os.makedirs(dr); os.makedirs(do)
    ref = _stores(PEPPAN.MapBsn, dr)
    PEPPAN.get_map_bsn(os.path.join(dr, 'run'), 'x', genomes, 'o', 'p', ref[0], ref[1], ref[2], ref[3], save_seq)
    for s in ref:
        s.conn.close()
    ours = _stores(hitio.FlatStore, do)
    consumers.get_map_bsn(os.path.join(do, 'run'), 'x', genomes, 'o', 'p', ours[0], ours[1], ours[2], ours[3], save_seq, params={},
                          mapper=lambda data: tuple(x.copy() for x in results[data[2]]))
    for s in ours:
        s.close()
    ref, ours = _stores(PEPPAN.MapBsn, dr, 'r'), _stores(hitio.FlatStore, do, 'r')
    total = sum(len(r[0]) for r in results)
    assert ref[0].size() >= 50 and sum(len(v) for v in ours[0].values()) == total
    for a, b in zip(ref, ours):
        assert _store_equal(a, b)
    assert ref[1].size() == ((total + 999) // 1000 if save_seq else 0) and ref[2].size() == (total + 999) // 1000
    assert ref[3].size() == (total + 29999) // 30000
    # the same through files and a pool (typed flat files instead of pickled npz), into the reference's own store class
    if n_genomes <= 7:
        d2 = os.path.join(tmp_path, 'pool'); os.makedirs(d2)

        def flat_task(data):
            with hitio.FlatStore('%s.%d.bsn.pbs' % (data[0], data[2]), 'w') as st:
                st.save('bsn', results[data[2]][0]); st.save('ovl', results[data[2]][1])
            return '%s.%d' % (data[0], data[2])

        third = _stores(PEPPAN.MapBsn, d2)
        consumers.get_map_bsn(os.path.join(d2, 'run'), 'x', genomes, 'o', 'p', third[0], third[1], third[2], third[3], save_seq, params={},
                              pool=_SerialPool(), mapper=flat_task)
        for s in third:
            s.conn.close()
        assert not [f for f in os.listdir(d2) if f.endswith('.pbs')]
        for a, b in zip(ref, _stores(PEPPAN.MapBsn, d2, 'r')):
            assert _store_equal(a, b)


def test_get_map_bsn_end_to_end_equals_the_reference(PEPPAN, oracle_as_search, monkeypatch, tmp_path):
    """Three genomes through search -> comparison -> grouping -> merge: the reference's get_map_bsn (its own iter_map_bsn per
    genome, pickled .bsn.npz files, MapBsn stores) beside consumers.get_map_bsn (results kept in memory, flat stores)."""
    monkeypatch.setattr(PEPPAN, 'uberBlast', ub.uberBlast)
    if not hasattr(np.lib.npyio, 'format'):
        monkeypatch.setattr(np.lib.npyio, 'format', np.lib.format, raising=False)
    from peppan_b200 import hitio
    pool = workloads.GenePool(30, 30, seed=workloads.SEED + 53)
    clust = os.path.join(tmp_path, 'exemplar.fa')
    with open(clust, 'w') as f:
        for n, s in pool.fasta_items():
            f.write('>%s\n%s\n' % (n, s))
    genomes = {}
    for g in range(3):
        seq, annot = workloads.synth_genome(pool, g, n_acc_per_genome=15, seed=workloads.SEED + 53)
        mid = annot[len(annot) // 2]
        cut = (int(mid[1]) + int(mid[2])) // 2          # inside a gene: its two halves form a merge group (overlap entries, without
        genomes[1000 + 2 * g] = [900 + g, seq[:cut]]    # which the reference's function ends in an unbound `del ovl`, :990)
        genomes[1001 + 2 * g] = [900 + g, seq[cut:]]
    ortho = os.path.join(tmp_path, 'ortho.npy')
    np.save(ortho, np.zeros([0, 3], dtype=int), allow_pickle=True)
    params = dict(gtable=11, noDiamond=False, match_identity=0.5, match_frag_len=50., match_frag_prop=0.25, link_gap=600., link_diff=1.5,
                  match_prop=0.5, match_len=250., match_prop1=0.8, match_len1=100., match_prop2=0.4, match_len2=400.)
    monkeypatch.setattr(PEPPAN, 'pool', _SerialPool(), raising=False)
    monkeypatch.setattr(PEPPAN, 'params', params)
    monkeypatch.setattr(PEPPAN, 'logger', lambda *a, **k: None)
    dr, do = os.path.join(tmp_path, 'ref'), os.path.join(tmp_path, 'ours')
    os.makedirs(dr); os.makedirs(do)
    for cls, d in ((PEPPAN.MapBsn, dr), (hitio.FlatStore, do)):                # empty old-annotation stores, each in its own format
        st = cls(os.path.join(d, 'old.npz'), 'w')
        st._save(st.conn, '0', np.zeros([0, 4], dtype=object))
        st.close() if hasattr(st, 'close') else st.conn.close()
    ref = _stores(PEPPAN.MapBsn, dr)
    PEPPAN.get_map_bsn(os.path.join(dr, 'run'), clust, genomes, ortho, os.path.join(dr, 'old.npz'), ref[0], ref[1], ref[2], ref[3], True)
    for s in ref:
        s.conn.close()
    ours = _stores(hitio.FlatStore, do)
    consumers.get_map_bsn(os.path.join(do, 'run'), clust, genomes, ortho, os.path.join(do, 'old.npz'), ours[0], ours[1], ours[2], ours[3], True, params)
    for s in ours:
        s.close()
    ref, ours = _stores(PEPPAN.MapBsn, dr, 'r'), _stores(hitio.FlatStore, do, 'r')
    assert ref[0].size() >= 40 and ref[2].size() == 1
    for a, b in zip(ref, ours):
        assert _store_equal(a, b)
    assert sorted(os.listdir(do)) == ['clf.npz', 'mat.npz', 'old.npz', 'seq.npz', 'tab.npz']          # no per-genome files left behind


def test_batched_stage_equals_the_sequential_one(oracle, oracle_as_search, tmp_path):
    """consumers.get_map_bsn_batched (grouped search of several genomes by the process that owns the device, post-search chain
    and consumer loops in spawned worker processes that get the record tables handed in, merge in genome order) fills the
    four stores with the values of consumers.get_map_bsn (one uberBlast call with its own searches per genome).  The grouped
    search is stood in by the oracle genome by genome -- on the GPU, search.search_grouped returns per-genome tables equal to
    per-genome searches (tests/test_real_genomes.py)."""
    from peppan_b200 import hitio, seqcodec
    pool = workloads.GenePool(30, 30, seed=workloads.SEED + 59)
    clust = os.path.join(tmp_path, 'exemplar.fa')
    with open(clust, 'w') as f:
        for n, s in pool.fasta_items():
            f.write('>%s\n%s\n' % (n, s))
    genomes = {}
    for g in range(3):
        seq, annot = workloads.synth_genome(pool, g, n_acc_per_genome=15, seed=workloads.SEED + 59)
        mid = annot[len(annot) // 2]
        cut = (int(mid[1]) + int(mid[2])) // 2
        genomes[1000 + 2 * g] = [900 + g, seq[:cut].lower() if g == 1 else seq[:cut]]      # one genome in lower case
        genomes[1001 + 2 * g] = [900 + g, seq[cut:]]
    ortho = os.path.join(tmp_path, 'ortho.npy')
    np.save(ortho, np.zeros([0, 3], dtype=int), allow_pickle=True)
    params = dict(gtable=11, noDiamond=False, match_identity=0.5, match_frag_len=50., match_frag_prop=0.25, link_gap=600., link_diff=1.5,
                  match_prop=0.5, match_len=250., match_prop1=0.8, match_len1=100., match_prop2=0.4, match_len2=400.)
    old = os.path.join(tmp_path, 'old.pbs')
    with hitio.FlatStore(old, 'w') as st:
        st.save('1000', np.array([[5, 100, 900, '+'], [6, 1200, 2000, '-']], dtype=object))
    calls = []

    def grouped(ctx, qb, qo, tb, to, groups, mode, min_id, min_cov, min_ratio, gtable):
        res = []
        for g in range(int(groups.max()) + 1):
            idx = np.flatnonzero(groups == g)
            a = int(to[idx[0]])
            res.append(oracle.search(qb, qo, tb[a:int(to[idx[-1] + 1])], to[idx[0]:idx[-1] + 2] - a, mode, seqcodec.BLOSUM62.reshape(-1),
                                     min_id=min_id, min_cov=min_cov, min_ratio=min_ratio, gtable=gtable))
        calls.append((mode, len(res)))
        return res, {}

    outs = {}
    for tag, workers in (('sequential', None), ('inline', 0), ('pool', 2)):
        d = os.path.join(tmp_path, tag); os.makedirs(d)
        stores = _stores(hitio.FlatStore, d)
        if workers is None:
            consumers.get_map_bsn(os.path.join(d, 'run'), clust, genomes, ortho, old, stores[0], stores[1], stores[2], stores[3], True, params)
        else:
            consumers.get_map_bsn_batched(os.path.join(d, 'run'), clust, genomes, ortho, old, stores[0], stores[1], stores[2], stores[3], True, params,
                                          workers=workers, batch=2, grouped_search=grouped, timeout=120.)
        for s in stores:
            s.close()
        outs[tag] = _stores(hitio.FlatStore, d, 'r')
    assert calls.count((1, 2)) == 2 and calls.count((2, 2)) == 2 and calls.count((1, 1)) == 2 and calls.count((2, 1)) == 2      # batches of 2 and 1 genomes, both modes, twice
    assert outs['sequential'][0].size() >= 30
    for tag in ('inline', 'pool'):
        for a, b in zip(outs['sequential'], outs[tag]):
            assert _store_equal(a, b), tag


def test_a_genome_without_hits_adds_nothing(oracle_as_search, tmp_path):
    """Where the reference stops with an exception (np.max of an empty column, PEPPAN.py:775), the stage here carries on: the
    genome yields an empty result and the merged stores hold the other genomes only."""
    from peppan_b200 import hitio
    pool = workloads.GenePool(20, 10, seed=workloads.SEED + 61)
    clust = os.path.join(tmp_path, 'exemplar.fa')
    with open(clust, 'w') as f:
        for n, s in pool.fasta_items():
            f.write('>%s\n%s\n' % (n, s))
    rng = np.random.default_rng(3)
    seq, annot = workloads.synth_genome(pool, 0, n_acc_per_genome=5, seed=workloads.SEED + 61)
    mid = annot[len(annot) // 2]
    cut = (int(mid[1]) + int(mid[2])) // 2
    genomes = {2000: [800, ''.join('ACGT'[i] for i in rng.integers(0, 4, 30000))],      # random: no exemplar matches it
               2001: [801, seq[:cut]], 2002: [801, seq[cut:]]}
    ortho = os.path.join(tmp_path, 'ortho.npy')
    np.save(ortho, np.zeros([0, 3], dtype=int), allow_pickle=True)
    params = dict(gtable=11, noDiamond=False, match_identity=0.5, match_frag_len=50., match_frag_prop=0.25, link_gap=600., link_diff=1.5,
                  match_prop=0.5, match_len=250., match_prop1=0.8, match_len1=100., match_prop2=0.4, match_len2=400.)
    old = os.path.join(tmp_path, 'old.pbs')
    with hitio.FlatStore(old, 'w') as st:
        st.save('0', np.zeros([0, 4], dtype=object))
    bsn, ovl = consumers.iter_map_bsn((os.path.join(tmp_path, 'x'), clust, 0, 800, [[2000, genomes[2000][1]]], ortho, old, params), out='memory')
    assert len(bsn) == 0 and ovl.shape == (0, 3)
    outs = []
    for tag, gen in (('with', genomes), ('without', {k: v for k, v in genomes.items() if k != 2000})):
        d = os.path.join(tmp_path, tag); os.makedirs(d)
        stores = _stores(hitio.FlatStore, d)
        consumers.get_map_bsn(os.path.join(d, 'run'), clust, gen, ortho, old, stores[0], stores[1], stores[2], stores[3], True, params)
        for s in stores:
            s.close()
        outs.append(_stores(hitio.FlatStore, d, 'r'))
    assert outs[0][0].size() >= 15
    for a, b in zip(*outs):
        assert _store_equal(a, b)


def test_real_sequences_through_both_iter_map_bsn(PEPPAN, oracle_as_search, monkeypatch, tmp_path):
    """Real sequences (the committed slice of the bundled E. coli genomes: 164 CDS of GCF_000010485 against the syntenic 138 kb
    of GCF_001566635) with the REAL annotation of that region as old predictions: the reference's iter_map_bsn and
    consumers.iter_map_bsn write the same arrays -- real gene lengths, both strands, partial genes at the slice ends and the overlap
    fractions with annotated genes (the two strains are close: the alignments of this slice carry mismatches but no gaps; gapped
    hits are covered by the synthetic cases above)."""
    import gzip
    import json
    import re
    from peppan_b200 import ingest
    monkeypatch.setattr(PEPPAN, 'uberBlast', ub.uberBlast)
    if not hasattr(np.lib.npyio, 'format'):
        monkeypatch.setattr(np.lib.npyio, 'format', np.lib.format, raising=False)
    here = os.path.dirname(os.path.abspath(__file__))
    fx = json.load(gzip.open(os.path.join(here, 'golden', 'real_slice.json.gz'), 'rt'))
    contig, lo, hi = re.search(r'vs GCF_001566635 (\S+)\[(\d+):(\d+)\]', fx['source']).groups()
    lo, hi = int(lo), int(hi)
    clust = os.path.join(tmp_path, 'exemplar.fa')
    with open(clust, 'w') as f:
        for n, s in fx['queries']:
            f.write('>%d\n%s\n' % (int(n) + 1, s))
    _, cds = ingest.iter_readGFF((os.path.join(REF, 'examples', 'GCF_001566635.combined.gff.gz'), 'CDS', 11))
    rows = sorted([[5000 + k, c[2] - lo, c[3] - lo, c[4]] for k, c in enumerate(cds.values()) if c[1] == contig and c[2] > lo and c[3] <= hi], key=lambda r: r[1])
    assert len(rows) >= 100
    old = os.path.join(tmp_path, 'old.npz')
    store = PEPPAN.MapBsn(old, 'w')
    store._save(store.conn, '1001', np.array(rows, dtype=object))
    store.conn.close()
    ortho = os.path.join(tmp_path, 'ortho.npy')
    np.save(ortho, np.array([[i, i + 1, 9000 if i % 2 else -9000] for i in range(1, 120)], dtype=int), allow_pickle=True)
    params = dict(gtable=11, noDiamond=False, match_identity=0.5, match_frag_len=50., match_frag_prop=0.25, link_gap=600., link_diff=1.5,
                  match_prop=0.5, match_len=250., match_prop1=0.8, match_len1=100., match_prop2=0.4, match_len2=400.)
    contigs = [(1001, fx['target'][0][1])]
    a = PEPPAN.iter_map_bsn((os.path.join(tmp_path, 'ref'), clust, 0, 'taxon', contigs, ortho, old, params))
    b = consumers.iter_map_bsn((os.path.join(tmp_path, 'ours'), clust, 0, 'taxon', contigs, ortho, old, params), store=PEPPAN.MapBsn)
    ra, rb = np.load(a + '.bsn.npz', allow_pickle=True), np.load(b + '.bsn.npz', allow_pickle=True)
    assert ra['bsn'].shape == rb['bsn'].shape and len(ra['bsn']) >= 120
    assert _same(rb['bsn'], ra['bsn'])
    assert ra['ovl'].shape == rb['ovl'].shape and np.array_equal(ra['ovl'], rb['ovl'])
    col10 = [float(t[10]) for g in rb['bsn'] for t in g[6]]
    assert sum(1 for v in col10 if v > 0.9) >= 80                                          # most hits sit on an annotated gene
