"""Host post-search stages against golden vectors produced by the reference's own Python
(tests/golden/make_golden.py -> modules/uberBlast.py RunBlast.run)."""
import copy
import json
import os

import numpy as np
import pytest

from peppan_b200 import postfilter as pf

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
            rows = pf.ovl_filter(rows, o['filter'][1], o['filter'][2])
        if o['linear_merge'][0]:
            rows = pf.linear_merge(rows, o['linear_merge'][1], o['linear_merge'][2])
        pf.fix_end(rows, o['fix_end'][0], o['fix_end'][1])
        ovl = pf.overlaps(rows, o['return_overlap'][1], o['return_overlap'][2]) if o['return_overlap'][0] else None
        rows = pf.final_sort(rows)
        _same(rows, run['tab_out'], 'tab')
        if ovl is not None:
            _same(ovl.tolist(), run['overlap_out'], 'overlap')
