"""clust() / getClust() file contract on the CPU: pb_cluster is replaced -- in this test only -- by the oracle's search +
scalar greedy (the definition the GPU path is checked against in tests/test_clust_gpu.py)."""
import os

import numpy as np
import pytest

from peppan_b200 import clust, seqio

from test_clust_gpu import _genes


def test_getclust_files_cpu(oracle_as_cluster, tmp_path):
    items = _genes(3, n_anc=25)
    fa = os.path.join(tmp_path, 'genes.fa')
    with open(fa, 'w') as f:
        for n, s in items:
            f.write('>%s some description\n' % n)
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + '\n')
    prefix = os.path.join(tmp_path, 'out')
    ex, tab = clust.getClust(prefix, fa, dict(identity=0.9, coverage=0.8, n_thread=4, translate=False))
    assert ex == prefix + '.clust.exemplar' and tab == prefix + '.clust.tab'
    pairs = [l.rstrip('\n').split('\t') for l in open(tab)]
    assert [p[0] for p in pairs] == sorted(n for n, _ in items)          # sorted by gene name (modules/clust.py:104)
    reps = set(p[1] for p in pairs)
    ex_names = [l[1:].split()[0] for l in open(ex) if l.startswith('>')]
    assert set(ex_names) == reps and ex_names == [n for n, _ in items if n in reps]   # input order
    src, out = open(fa).read(), open(ex).read()
    for rec in out.split('>')[1:]:                                        # records are re-emitted verbatim (:72-88)
        assert ('>' + rec) in src
    assert all(dict(pairs)[r] == r for r in reps) and 25 <= len(reps) < len(items)
    # every member's representative precedes it in the input order (the first member of a cluster is its representative)
    pos = {n: i for i, (n, _) in enumerate(items)}
    assert all(pos[p[1]] <= pos[p[0]] for p in pairs)
    ex2, tab2 = clust.clust(['-i', fa, '-p', prefix + '2', '-d', '0.9', '-c', '0.8'])
    assert open(tab2).read() == open(tab).read() and open(ex2).read() == open(ex).read()
    # -a: exemplar records are re-emitted as one-line nucleotide records (modules/clust.py:95-100)
    ex3, tab3 = clust.clust(['-i', fa, '-p', prefix + '3', '-d', '0.9', '-c', '0.8', '-a'])
    reps3 = set(l.rstrip('\n').split('\t')[1] for l in open(tab3))
    lines3 = open(ex3).read().split('\n')
    seqs = dict(items)
    assert [l[1:] for l in lines3 if l.startswith('>')] == [n for n, _ in items if n in reps3]
    assert all(lines3[i + 1] == seqs[lines3[i][1:]] for i in range(0, len(lines3) - 1, 2))
