"""Host post-search stages against golden vectors produced by the reference's own Python
(tests/golden/make_golden.py -> modules/uberBlast.py RunBlast.run)."""
import copy
import json
import os

import numpy as np
import pytest

from peppan_b200 import postfilter as pf
import postfilter_mirror as pfm

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _load(name):
    with open(os.path.join(GOLD, name + '.json')) as f:
        return json.load(f)


def _same(a, b, path=''):
    if isinstance(a, float) or isinstance(b, float):
        assert abs(float(a) - float(b)) <= 1e-9 * max(1.0, abs(float(b))), '%s: %r != %r' % (path, a, b)
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), '%s: length %d != %d (%r vs %r)' % (path, len(a), len(b), a, b)
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, '%s[%d]' % (path, i))
    else:
        assert a == b, '%s: %r != %r' % (path, a, b)


def test_cigar2score_modes_match_reference():
    for c in _load('cigar2score'):
        iden, score = pf.cigar2score(c['cigar'], pf.encode_nuc(c['r']), pf.encode_nuc(c['q']), c['frame'], c['mode'])
        assert abs(iden - c['iden']) < 1e-12 and abs(score - c['score']) < 1e-9, c['mode']


def test_np_round_half_even():
    for x, y in _load('np_round'):
        assert float(np.round(x, 3)) == y


@pytest.mark.parametrize('scen', range(14))
def test_post_chain_matches_reference(scen):
    g = _load('post_chain')[scen]
    ref_enc = {k: pf.encode_nuc(v) for k, v in g['contigs'].items()}
    qry_enc = {k: pf.encode_nuc(v) for k, v in g['genes'].items()}
    for run in g['runs']:
        o = run['opts']
        rows = copy.deepcopy(g['rows_in'])
        for i, t in enumerate(rows):
            t.append(i)
        if o['re_score']:
            rows = pf.rescore(rows, ref_enc, qry_enc, o['re_score'], g['min_id'])
        if o['filter'][0]:
            rows = pfm.ovl_filter(rows, o['filter'][1], o['filter'][2])
        if o['linear_merge'][0]:
            rows = pfm.linear_merge(rows, o['linear_merge'][1], o['linear_merge'][2])
        pfm.fix_end(rows, o['fix_end'][0], o['fix_end'][1])
        ovl = pfm.overlaps(rows, o['return_overlap'][1], o['return_overlap'][2]) if o['return_overlap'][0] else None
        rows = pfm.final_sort(rows)
        _same(rows, run['tab_out'], 'tab')
        if ovl is not None:
            _same(ovl.tolist(), run['overlap_out'], 'overlap')


def test_rescore_m1_table_equals_cigar2score():
    """pb_rescore_m1 (library, whole table) against the per-hit Python mirror pinned to the reference above"""
    rng = np.random.default_rng(7)
    nt = np.frombuffer(b'ACGTN', dtype=np.uint8)
    contigs = {'c%d' % i: nt[rng.choice(5, 3000, p=[.24, .24, .24, .24, .04])].tobytes().decode() for i in range(3)}
    genes, rows = {}, []
    comp = str.maketrans('ACGTN', 'TGCAN')
    for h in range(200):
        cname = 'c%d' % rng.integers(3); ctg = contigs[cname]
        # random alignment: M / I / D runs
        cig, qlen, slen = [], 0, 0
        for k in range(int(rng.integers(1, 8))):
            n = int(rng.integers(1, 120)); cig.append([n, 'M']); qlen += n; slen += n
            if rng.random() < 0.7:
                g = int(rng.integers(1, 9)); op = 'I' if rng.random() < 0.5 else 'D'
                cig.append([g, op]); qlen += g if op == 'I' else 0; slen += g if op == 'D' else 0
        if cig[-1][1] != 'M':
            cig.append([5, 'M']); qlen += 5; slen += 5
        s0 = int(rng.integers(0, 3000 - slen)); minus = rng.random() < 0.5
        sub = ctg[s0:s0 + slen]
        if minus:
            sub = sub.translate(comp)[::-1]
        # query: the subject bases along the path with some substitutions
        q, ri = [], 0
        for n, op in cig:
            if op == 'M':
                seg = np.frombuffer(sub[ri:ri + n].encode(), dtype=np.uint8).copy()
                m = rng.random(n) < 0.15; seg[m] = nt[rng.integers(0, 5, int(m.sum()))]
                q.append(seg.tobytes().decode()); ri += n
            elif op == 'D':
                ri += n
            else:
                q.append(nt[rng.integers(0, 4, n)].tobytes().decode())
        pre, post = int(rng.integers(0, 10)), int(rng.integers(0, 10))
        gname = str(h)
        genes[gname] = 'A' * pre + ''.join(q) + 'C' * post
        ss, se = (s0 + 1, s0 + slen) if not minus else (s0 + slen, s0 + 1)
        rows.append([gname, cname, 0.5, 0, 0, 0, pre + 1, pre + qlen, ss, se, 0.0, 1, len(genes[gname]), 3000, cig, h])
    from peppan_b200 import seqio
    want = pf.rescore(copy.deepcopy(rows), {k: pf.encode_nuc(v) for k, v in contigs.items()}, {k: pf.encode_nuc(v) for k, v in genes.items()}, 1, 0.8)
    got = pf.rescore_m1_table(copy.deepcopy(rows), seqio.to_seqset(genes), seqio.to_seqset(contigs), 0.8)
    assert 0 < len(want) < len(rows)
    _same([list(t) for t in got], [list(t) for t in want])
    assert all(type(t[2]) is float and type(t[11]) is float for t in got)
    # the golden post-chain scenarios with re_score = 1 through the table path
    for g in _load('post_chain'):
        for run in g['runs']:
            if run['opts']['re_score'] != 1:
                continue
            rows_in = copy.deepcopy(g['rows_in'])
            for i, t in enumerate(rows_in):
                t.append(i)
            a = pf.rescore(copy.deepcopy(rows_in), {k: pf.encode_nuc(v) for k, v in g['contigs'].items()},
                           {k: pf.encode_nuc(v) for k, v in g['genes'].items()}, 1, g['min_id'])
            b = pf.rescore_m1_table(copy.deepcopy(rows_in), seqio.to_seqset(g['genes']), seqio.to_seqset(g['contigs']), g['min_id'])
            _same([list(t) for t in b], [list(t) for t in a])


def test_rescore_m1_table_rejects_inconsistent_rows():
    """a CIGAR that runs past its coordinate slice, or coordinates outside the sequence, are errors, not silent garbage"""
    from peppan_b200 import seqio
    genes = {'g': 'ACGTACGTACGTACGTACGT'}; contigs = {'c': 'TTACGTACGTACGTACGTACGTTT'}
    good = ['g', 'c', 0.0, 0, 0, 0, 1, 20, 3, 22, 0.0, 0, 20, 24, [[20, 'M']], 0]
    out = pf.rescore_m1_table([list(good)], seqio.to_seqset(genes), seqio.to_seqset(contigs), 0.5)
    assert out[0][2] == 1.0 and out[0][11] == 60.0
    for bad in (dict(cig=[[25, 'M']]), dict(send=30), dict(qend=40)):
        row = list(good)
        row[14] = bad.get('cig', row[14]); row[9] = bad.get('send', row[9]); row[7] = bad.get('qend', row[7])
        with pytest.raises(RuntimeError):
            pf.rescore_m1_table([row], seqio.to_seqset(genes), seqio.to_seqset(contigs), 0.5)
    assert pf.rescore_m1_table([], seqio.to_seqset(genes), seqio.to_seqset(contigs), 0.5) == []


def _python_chain(rows, o):
    if o['filter'][0]:
        rows = pfm.ovl_filter(rows, o['filter'][1], o['filter'][2])
    if o['linear_merge'][0]:
        rows = pfm.linear_merge(rows, o['linear_merge'][1], o['linear_merge'][2])
    pfm.fix_end(rows, o['fix_end'][0], o['fix_end'][1])
    ovl = pfm.overlaps(rows, o['return_overlap'][1], o['return_overlap'][2]) if o['return_overlap'][0] else None
    return pfm.final_sort(rows), ovl


@pytest.mark.parametrize('scen', range(14))
def test_library_post_chain_matches_reference_golden(scen):
    """pb_post_chain (host C++, whole table) on the reference's golden scenarios"""
    g = _load('post_chain')[scen]
    ref_enc = {k: pf.encode_nuc(v) for k, v in g['contigs'].items()}
    qry_enc = {k: pf.encode_nuc(v) for k, v in g['genes'].items()}
    for run in g['runs']:
        o = run['opts']
        rows = copy.deepcopy(g['rows_in'])
        for i, t in enumerate(rows):
            t.append(i)
        if o['re_score']:
            rows = pf.rescore(rows, ref_enc, qry_enc, o['re_score'], g['min_id'])
        rows, ovl = pf.post_chain_table(rows, o['filter'], o['linear_merge'], o['fix_end'], o['return_overlap'])
        _same(rows, run['tab_out'], 'tab')
        if o['return_overlap'][0]:
            _same(ovl.tolist(), run['overlap_out'], 'overlap')


def _random_table(rng, n_genes, n_contigs, n_rows):
    """hits that overlap, nest, chain and sit at contig ends often enough to reach every branch of the chain"""
    rows = []
    clen = [int(rng.integers(3000, 9000)) for _ in range(n_contigs)]
    glen = [int(rng.integers(150, 1500)) for _ in range(n_genes)]
    anchors = [(int(rng.integers(n_contigs)), int(rng.integers(0, 2)), float(rng.random())) for _ in range(n_genes)]
    for h in range(n_rows):
        gi = int(rng.integers(n_genes)); L = glen[gi]
        ci, minus, pos = anchors[gi] if rng.random() < 0.8 else (int(rng.integers(n_contigs)), int(rng.integers(0, 2)), float(rng.random()))
        qa = int(rng.integers(1, max(2, L // 2))) if rng.random() < 0.6 else int(rng.integers(1, 8))
        qb = int(rng.integers(qa + 30, L + 1)) if qa + 30 < L else L
        span = qb - qa + 1
        cig = [[span, 'M']]
        sspan = span
        if rng.random() < 0.4 and span > 40:
            a = int(rng.integers(10, span - 10)); g_ = int(rng.integers(1, 12))
            if rng.random() < 0.5:
                cig = [[a, 'M'], [g_, 'D'], [span - a, 'M']]; sspan = span + g_
            else:
                cig = [[a, 'M'], [g_, 'I'], [span - a - g_, 'M']] if span - a - g_ > 0 else [[span, 'M']]
                sspan = span - g_ if len(cig) == 3 else span
        C_ = clen[ci]
        base = int(pos * (C_ - 2 * L - 700)) + 300 if rng.random() < 0.85 else int(rng.integers(1, 40))
        base = max(1, min(base, C_ - sspan - 1))
        s0 = base + (qa if not minus else (L - qb)) + int(rng.integers(-3, 4)) * int(rng.random() < 0.3)
        s0 = max(1, min(s0, C_ - sspan + 1))
        ss, se = (s0, s0 + sspan - 1) if not minus else (s0 + sspan - 1, s0)
        iden = round(float(rng.uniform(0.45, 1.0)), 3)
        score = round(float(span * (4 * iden - 1) + rng.integers(-5, 6)), 3)
        rows.append([str(100 + gi), 'ctg%d' % ci, iden, sum(k for k, _ in cig), 0, len(cig) // 2, qa, qb, ss, se, 1e-50, score, L, C_, cig, h])
    return rows


@pytest.mark.parametrize('seed', range(12))
def test_library_post_chain_equals_python_mirror_on_random_tables(seed):
    rng = np.random.default_rng(1000 + seed)
    rows = _random_table(rng, n_genes=int(rng.integers(3, 25)), n_contigs=int(rng.integers(1, 4)), n_rows=int(rng.integers(5, 160)))
    for o in (dict(filter=[True, 0.9, 0.], linear_merge=[True, 600., 1.5], fix_end=[0., 3.], return_overlap=[True, 300, 0.6]),
              dict(filter=[False, 0.9, 0.], linear_merge=[False, 300., 1.2], fix_end=[3., 3.], return_overlap=[False, 300, 0.6]),
              dict(filter=[True, 0.5, 10.], linear_merge=[True, 300., 1.2], fix_end=[6., 6.], return_overlap=[True, 100, 0.3]),
              dict(filter=[False, 0.9, 0.], linear_merge=[True, 2000., 3.0], fix_end=[0., 0.], return_overlap=[True, 300, 0.6])):
        want, wovl = _python_chain(copy.deepcopy(rows), o)
        got, govl = pf.post_chain_table(copy.deepcopy(rows), o['filter'], o['linear_merge'], o['fix_end'], o['return_overlap'])
        _same([list(t) for t in got], [list(t) for t in want], 'seed %d' % seed)
        if o['return_overlap'][0]:
            # with >= 32 hits of one query the Python code walks a set of row indices whose order is a CPython hash-table
            # detail; it only permutes the rows of the overlap table (pb_post.cu header)
            many = max(np.unique([t[0] for t in rows], return_counts=True)[1]) >= 32
            assert (sorted(map(tuple, govl.tolist())) == sorted(map(tuple, wovl.tolist()))) if many else (govl.tolist() == wovl.tolist())
        else:
            assert govl is None


@pytest.mark.parametrize('scen', range(14))
def test_runblast_run_wiring_on_golden_scenarios(scen):
    """RunBlast.run with the search tool replaced by the scenario's rows (as tests/golden/make_golden.py did with the
    reference's own class): the whole host chain of the shim -- library re-scoring, library post-chain, object array --
    must reproduce the reference's output table and overlap list."""
    from peppan_b200.uberBlast import RunBlast
    g = _load('post_chain')[scen]

    class Fake(RunBlast):
        def runBlast(self, ref, qry):
            arr = np.empty([len(g['rows_in']), 15], dtype=object)
            for i, r in enumerate(copy.deepcopy(g['rows_in'])):
                for j in range(15):
                    arr[i, j] = r[j]
            return arr

    for run in g['runs']:
        o = run['opts']
        rb = Fake()
        rb.qrySeq = dict(g['genes']); rb.refSeq = dict(g['contigs'])
        res = rb.run('unused_ref', 'unused_qry', ['blastn'], g['min_id'], g['min_cov'], g['min_ratio'], re_score=o['re_score'],
                     filter=o['filter'], linear_merge=o['linear_merge'], return_overlap=o['return_overlap'], fix_end=o['fix_end'])
        tab, ovl = res if o['return_overlap'][0] else (res, None)
        _same([list(r) for r in tab], run['tab_out'], 'tab')
        if ovl is not None:
            _same(ovl.tolist(), run['overlap_out'], 'overlap')


def test_rescore_m1_counts_matching_columns_like_nucencoder_on_any_bytes():
    """pb_rescore_m1 compares 16 columns at a time where it can: single-run alignments of every length 1..70 at every offset
    parity, both strands, over an alphabet with ambiguity codes, gaps and lower case, against the class rule of the reference
    (nucEncoder, modules/uberBlast.py:270-271: A 0, C 1, G 3, T 4, anything else 2; minus strand: 4 - class of the reversed
    subject)."""
    from peppan_b200 import seqio
    rng = np.random.default_rng(23)
    alpha = np.frombuffer(b'ACGTACGTACGTNRYacgt-*', dtype=np.uint8)
    enc = np.full(256, 2, dtype=np.int64); enc[[ord(c) for c in 'ACGT']] = (0, 1, 3, 4)
    genes = {'g': alpha[rng.integers(0, len(alpha), 400)].tobytes().decode()}
    contigs = {'c': alpha[rng.integers(0, len(alpha), 400)].tobytes().decode()}
    # plant similarity: copy stretches of the gene (and of its reverse complement class-wise) into the contig
    g = np.frombuffer(genes['g'].encode(), dtype=np.uint8); c = np.frombuffer(contigs['c'].encode(), dtype=np.uint8).copy()
    c[50:150] = g[50:150]
    comp = np.arange(256, dtype=np.uint8); comp[[ord(x) for x in 'ACGT']] = [ord(x) for x in 'TGCA']
    c[200:300] = comp[g[100:200]][::-1]
    contigs['c'] = c.tobytes().decode()
    rows, want = [], []
    for n in range(1, 71):
        for minus in (False, True):
            qa = int(rng.integers(0, 300)); ta = int(rng.integers(40, 300)) if not minus else int(rng.integers(190, 320))
            if n % 2 == 0:                 # on the planted stretches: mostly matching columns
                k = int(rng.integers(0, 100 - n + 1))
                qa, ta = (50 + k, 50 + k) if not minus else (100 + k, 200 + 100 - k - n)
            q = g[qa:qa + n]; t = c[ta:ta + n]
            tt = t if not minus else t[::-1]
            same = enc[q] == (enc[tt] if not minus else 4 - enc[tt])
            nm = int(same.sum())
            rows.append(['g', 'c', 0.0, n, 0, 0, qa + 1, qa + n, (ta + 1) if not minus else (ta + n), (ta + n) if not minus else (ta + 1), 0.0, 0, 400, 400, [[n, 'M']], len(rows)])
            want.append((round(nm / float(n), 3), float(nm * 3 - (n - nm))))
    got = pf.rescore_m1_table([list(r) for r in rows], seqio.to_seqset(genes), seqio.to_seqset(contigs), -1.0)
    assert len(got) == len(rows)
    for r, (iden, score) in zip(got, want):
        assert abs(r[2] - iden) < 1e-12 and r[11] == score, (r[3], r[8] > r[9], r[2], iden, r[11], score)
    assert sum(1 for w in want if w[0] > 0.9) >= 10 and sum(1 for w in want if w[0] < 0.5) >= 10
