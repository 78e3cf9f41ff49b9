"""GPU parity of pb_cluster / getClust: the cluster assignment must equal the scalar greedy of
oracle/pb_oracle.c applied to the edges the scalar search oracle verifies."""
import os

import numpy as np
import pytest

from peppan_b200 import clust, seqcodec, seqio, workloads

pytestmark = pytest.mark.gpu


def _genes(seed, n_anc=60, copies=4):
    """priority-ordered gene set: diverged copies (some truncated) of ancestral genes, shuffled"""
    rng = np.random.default_rng(seed)
    pool = workloads.GenePool(n_anc, 0, seed=workloads.SEED + seed)
    items = []
    for a in range(n_anc):
        for c in range(int(rng.integers(1, copies + 1))):
            g = workloads._diverge(rng, pool.genes[a], float(rng.uniform(0.85, 1.0)))
            if rng.random() < 0.2:
                g = g[:int(g.size * rng.uniform(0.5, 0.95))]
            items.append(workloads._NT[g].tobytes().decode())
    items.sort(key=lambda s: -len(s))            # PEPPAN orders by priority, then longer first (PEPPAN.py:1027)
    return [(str(i), s) for i, s in enumerate(items)]


def _oracle_clusters(oracle, items, identity, coverage, translate=False):
    names, buf, off = seqio.to_seqset(items)
    hits, cig = oracle.search(buf, off, buf, off, 3 if translate else (1 | 256), seqcodec.BLOSUM62.reshape(-1), min_id=identity - 0.005,
                              min_cov=0, min_ratio=max(0.0, coverage - 0.005), max_hits=0x7fffffff)
    ea, eb = [], []
    for h in hits:
        a, b = int(h['s_id']), int(h['q_id'])
        if a >= b or (translate and int(h['frame']) != 1):
            continue
        ops = cig[int(h['cigar_off']):int(h['cigar_off']) + int(h['cigar_n'])]
        gapb = int(sum(int(o) >> 2 for o in ops if int(o) & 3))
        nm = int(h['aln_len']) - int(h['mismatch']) - gapb
        iden = nm / float(h['aln_len'])
        qc = (h['q_end'] - h['q_start'] + 1) / float(h['q_len']); sc = (h['s_end'] - h['s_start'] + 1) / float(h['s_len'])
        if iden + 1e-9 >= np.float32(identity) and qc + 1e-9 >= np.float32(coverage) and sc + 1e-9 >= np.float32(coverage):
            ea.append(a); eb.append(b)
    order = np.lexsort((ea, eb))
    return oracle.greedy_cluster(len(items), np.array(ea, np.int32)[order], np.array(eb, np.int32)[order])


@pytest.mark.parametrize('identity,coverage', [(0.9, 0.8), (0.99, 0.8), (1.0, 0.8), (0.95, 0.5)])
def test_cluster_matches_oracle_greedy(ctx, oracle, identity, coverage):
    items = _genes(1)
    names, buf, off = seqio.to_seqset(items)
    rep, st = clust.cluster(ctx, buf, off, identity, coverage)
    want = _oracle_clusters(oracle, items, identity, coverage)
    assert np.array_equal(rep, want)
    assert (rep <= np.arange(len(rep))).all() and (rep[rep] == rep).all()
    assert st['n_reps'] == int((rep == np.arange(len(rep))).sum()) and st['kernel_launches'] > 0


@pytest.mark.parametrize('identity,coverage', [(0.9, 0.8), (0.7, 0.5)])
def test_translated_cluster_matches_oracle_greedy(ctx, oracle, identity, coverage):
    # clust -a: genes compared as frame-1 proteins (modules/clust.py:38-46)
    items = _genes(4, n_anc=50)
    names, buf, off = seqio.to_seqset(items)
    rep, st = clust.cluster(ctx, buf, off, identity, coverage, translate=True)
    want = _oracle_clusters(oracle, items, identity, coverage, translate=True)
    assert np.array_equal(rep, want)
    assert 0 < st['n_reps'] < len(items)


def test_family_larger_than_any_hit_cap_stays_one_cluster(ctx):
    # 1,100 alleles of one gene: gene 0 carries three SNPs, every other gene one SNP of its own, so each later gene scores
    # higher against its ~1,100 siblings than against gene 0 -- a per-query top-1000 by score would cut the edge to the only
    # representative and split the family.  Identity to gene 0 is 146/150 = 0.973 >= 0.97: one cluster.
    rng = np.random.default_rng(5)
    base = workloads._random_gene(rng, 150)
    genes = []
    for k in range(1100):
        g = base.copy()
        for p in ([10, 50, 90] if k == 0 else [12 + (k % 120)]):
            g[p] = (g[p] + 1 + (k // 120) % 3) % 4 if k else (g[p] + 1) % 4
        genes.append(workloads._NT[g].tobytes().decode())
    names, buf, off = seqio.to_seqset([(str(i), s) for i, s in enumerate(genes)])
    rep, st = clust.cluster(ctx, buf, off, 0.97, 0.8)
    assert (rep == 0).all() and st['n_reps'] == 1


def test_remembered_alignments_do_not_change_the_ladder(ctx, oracle):
    # iterClust's shape: nested gene sets at falling thresholds.  The second and third rung find most of their pairs in the
    # context's memo (pb_cluster_forget / pb_memo.h); the assignments must be those of a context that remembers nothing,
    # and those of the oracle.
    items = _genes(6, n_anc=70)
    clust.forget(ctx)
    cur = items
    remembered = []
    for identity in (0.97, 0.93, 0.9):
        names, buf, off = seqio.to_seqset(cur)
        rep, st = clust.cluster(ctx, buf, off, identity, 0.8)
        remembered.append(int(st['n_pairs_remembered']))
        want = _oracle_clusters(oracle, cur, identity, 0.8)
        assert np.array_equal(rep, want), identity
        cur = [cur[i] for i in np.nonzero(rep == np.arange(len(rep)))[0]]
    assert remembered[0] == 0 and remembered[1] > 0 and remembered[2] > 0
    clust.forget(ctx)
    names, buf, off = seqio.to_seqset(items)
    rep, st = clust.cluster(ctx, buf, off, 0.97, 0.8)
    assert int(st['n_pairs_remembered']) == 0


def test_cluster_blocked_equals_single_block(ctx, monkeypatch):
    items = _genes(2, n_anc=80)
    names, buf, off = seqio.to_seqset(items)
    rep1, st1 = clust.cluster(ctx, buf, off, 0.9, 0.8)
    monkeypatch.setenv('PB_CLUSTER_BLOCK', '20000')
    rep2, st2 = clust.cluster(ctx, buf, off, 0.9, 0.8)
    assert st2['n_blocks'] > 3 and np.array_equal(rep1, rep2)


def test_getclust_files(ctx, tmp_path):
    from peppan_b200 import uberBlast
    uberBlast.set_context(ctx)
    items = _genes(3, n_anc=40)
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
    assert set(ex_names) == reps and ex_names == [n for n, _ in items if n in reps]   # input order, verbatim records
    assert all(dict(pairs)[r] == r for r in reps) and 40 <= len(reps) < len(items)
    # the CLI wrapper produces the same files
    ex2, tab2 = clust.clust(['-i', fa, '-p', prefix + '2', '-d', '0.9', '-c', '0.8'])
    assert open(tab2).read() == open(tab).read()
    # -a: translated clustering; exemplar records are re-emitted as one-line nucleotide records (modules/clust.py:95-100)
    ex3, tab3 = clust.clust(['-i', fa, '-p', prefix + '3', '-d', '0.9', '-c', '0.8', '-a'])
    pairs3 = [l.rstrip('\n').split('\t') for l in open(tab3)]
    reps3 = set(p[1] for p in pairs3)
    lines3 = open(ex3).read().split('\n')
    seqs = dict(items)
    assert [l[1:] for l in lines3 if l.startswith('>')] == [n for n, _ in items if n in reps3]
    assert all(lines3[i + 1] == seqs[lines3[i][1:]] for i in range(0, len(lines3) - 1, 2))
