"""End-to-end uberBlast() shim on the GPU with the flag sets PEPPAN itself uses
(PEPPAN.py:226-230 get_similar_pairs, :767-772 iter_map_bsn)."""
import os
import re

import numpy as np
import pytest

from peppan_b200 import workloads
from peppan_b200.uberBlast import uberBlast

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp('ub')
    pool = workloads.GenePool(80, 120, seed=workloads.SEED + 11)
    seq, annot = workloads.synth_genome(pool, 0, n_acc_per_genome=40, seed=workloads.SEED + 11)
    qry = os.path.join(d, 'exemplar.fa'); ref = os.path.join(d, 'genome.fa')
    with open(qry, 'w') as f:
        for n, s in pool.fasta_items():
            f.write('>%s\n%s\n' % (n, s))
    with open(ref, 'w') as f:
        f.write('>contig1 test\n')
        for i in range(0, len(seq), 70):
            f.write(seq[i:i + 70] + '\n')
    return dict(qry=qry, ref=ref, pool=pool, annot=annot, glen=len(seq))


def _cigar_spans(c):
    q = s = 0
    for n, t in re.findall(r'(\d+)([MID])', c):
        n = int(n)
        if t in 'MI':
            q += n
        if t in 'MD':
            s += n
    return q, s


def test_iter_map_bsn_flags(files):
    args = '-r {ref} -q {qry} -f -m -O --blastn --diamond --min_id 0.4 --min_cov 50 --min_ratio 0.25 --merge_gap 600 --merge_diff 1.5 -t 1 -s 1 -e 0,3 --gtable 11'.format(**files).split()
    blastab, overlap = uberBlast(args)
    assert blastab.dtype == object and blastab.shape[1] == 17 and overlap.shape[1] == 3 and overlap.dtype.kind == 'i'
    assert len(blastab) > 100
    keys = [(r[0], r[1], r[11]) for r in blastab]
    assert keys == sorted(keys)                                   # sort_values([0, 1, 11])
    for r in blastab:
        assert isinstance(r[0], str) and isinstance(r[1], str) and isinstance(r[14], str) and isinstance(r[15], int)
        assert 0.4 <= r[2] <= 1.0 and round(r[2], 3) == r[2]
        assert 1 <= r[6] <= r[7] <= r[12] and 1 <= min(r[8], r[9]) and max(r[8], r[9]) <= r[13] == files['glen']
        qspan, sspan = _cigar_spans(r[14])
        assert qspan == r[7] - r[6] + 1 and sspan == abs(r[9] - r[8]) + 1
        assert r[16][0] >= r[11] - 1e-9 and r[15] in r[16][3:]
    ids = set(int(r[15]) for r in blastab)
    assert set(overlap[:, 0].tolist()) | set(overlap[:, 1].tolist()) <= ids
    # every gene copy planted at >= 90 % identity is recovered over >= 80 % of its length
    got = {}
    for r in blastab:
        got[int(r[0])] = max(got.get(int(r[0]), 0), (r[7] - r[6] + 1) / float(r[12]))
    planted = [a for a in files['annot'] if a[4] >= 0.9 and (a[2] - a[1]) >= 0.99 * len(files['pool'].genes[a[0]])]
    found = sum(1 for a in planted if got.get(a[0], 0) >= 0.8)
    assert found == len(planted), (found, len(planted))


def test_get_similar_pairs_flags(files):
    args = '-r {qry} -q {qry} --blastn --diamond -s 1 --min_id 0.45 --min_cov 50 -t 4 --min_ratio 0.25 -e 3,3 -p --gtable 11'.format(**files).split()
    blastab = uberBlast(args, extPool='ignored')
    assert blastab.shape[1] == 16
    selfhits = [r for r in blastab if r[0] == r[1] and r[6] == 1 and r[7] == r[12]]
    assert len(set(r[0] for r in selfhits)) == 200                # every exemplar aligns to itself end to end
    assert all(r[2] == 1.0 for r in selfhits)


def test_no_methods_and_output_file(files, tmp_path):
    out = uberBlast(['-r', files['ref'], '-q', files['qry']])
    assert out.shape == (0, 16)
    tab, ovl = uberBlast(['-r', files['ref'], '-q', files['qry'], '-O'])
    assert tab.shape == (0, 16) and ovl.shape == (0, 3)
    path = os.path.join(tmp_path, 'o.tsv')
    res = uberBlast(['-r', files['ref'], '-q', files['qry'], '--blastn', '-o', path])
    assert len(open(path).read().strip().split('\n')) == len(res)
