"""GPU parity of pb_search against the scalar restatement of the same search specification
(oracle/pb_search_oracle.c): identical hit tables, field by field, CIGAR by CIGAR."""
import numpy as np
import pytest

from peppan_b200 import search, seqcodec, seqio, workloads

pytestmark = pytest.mark.gpu


def _world(n_core=60, n_acc=90, genomes=1, seed=0):
    pool = workloads.GenePool(n_core, n_acc, seed=workloads.SEED + seed)
    contigs = []
    for g in range(genomes):
        seq, annot = workloads.synth_genome(pool, g, n_acc_per_genome=n_acc // 3, seed=workloads.SEED + seed)
        # split into a few contigs so multi-contig bookkeeping is exercised
        cut = len(seq) // 3
        contigs += [('g%d_a' % g, seq[:cut]), ('g%d_b' % g, seq[cut:])]
    return pool, contigs


def _check(ctx, oracle, q_items, t_items, mode, **kw):
    qn, qb, qo = seqio.to_seqset(q_items); tn, tb, to = seqio.to_seqset(t_items)
    hits, cigar, st = search.search(ctx, qb, qo, tb, to, mode, **kw)
    ref, rcig = oracle.search(qb, qo, tb, to, mode, seqcodec.BLOSUM62.reshape(-1), **kw)
    assert len(hits) == len(ref), 'hit count gpu %d oracle %d' % (len(hits), len(ref))
    for k in hits.dtype.names:
        if k in ('identity', 'evalue'):
            assert np.allclose(hits[k], ref[k], rtol=1e-6, atol=1e-30), k
        else:
            bad = np.nonzero(hits[k] != ref[k])[0]
            assert len(bad) == 0, '%s differs at hit %d: gpu %s oracle %s' % (k, bad[0], hits[bad[0]], ref[bad[0]])
    assert np.array_equal(cigar, rcig)
    assert st['kernel_launches'] > 0
    return hits, st


def test_nucleotide_search_matches_oracle(ctx, oracle):
    pool, contigs = _world()
    hits, st = _check(ctx, oracle, pool.fasta_items(), contigs, search.MODE_NT, min_id=0.4, min_cov=50, min_ratio=0.25)
    assert len(hits) >= 60 and (hits['s_start'] > hits['s_end']).any() and (hits['s_start'] < hits['s_end']).any()


def test_protein_6frame_search_matches_oracle(ctx, oracle):
    pool, contigs = _world(seed=1)
    hits, st = _check(ctx, oracle, pool.fasta_items(), contigs, search.MODE_PROT6, min_id=0.4, min_cov=50, min_ratio=0.25)
    assert len(hits) >= 60 and set(np.unique(hits['frame'])) <= {1, 2, 3, 4, 5, 6} and (hits['frame'] > 3).any()
    assert (hits['aln_len'] % 3 == 0).all()


def test_protein_self_search_matches_oracle(ctx, oracle):
    pool, _ = _world(n_core=40, n_acc=20, seed=2)
    items = pool.fasta_items()
    # add diverged copies so that non-self hits exist
    rng = np.random.default_rng(5)
    extra = []
    for i in range(0, 40, 3):
        g = workloads._diverge(rng, pool.genes[i], 0.8)
        extra.append(('x%d' % i, workloads._NT[g].tobytes().decode()))
    hits, st = _check(ctx, oracle, items + extra, items + extra, search.MODE_PROT3_SELF, min_id=0.3, min_cov=40, min_ratio=0.05)
    assert (hits['q_id'] != hits['s_id']).any() and (hits['frame'] <= 3).all()


def test_self_nucleotide_all_vs_all(ctx, oracle):
    pool, _ = _world(n_core=50, n_acc=0, seed=3)
    items = pool.fasta_items()
    hits, st = _check(ctx, oracle, items, items, search.MODE_NT, min_id=0.45, min_cov=50, min_ratio=0.25)
    assert set(hits['q_id'][hits['q_id'] == hits['s_id']].tolist()) == set(range(50))


def test_ambiguous_bases_gtable4_and_empty(ctx, oracle):
    pool, contigs = _world(n_core=30, n_acc=10, seed=4)
    name, seq = contigs[0]
    s = list(seq)
    rng = np.random.default_rng(9)
    for k in rng.integers(0, len(s), size=200):
        s[k] = 'N'
    contigs[0] = (name, ''.join(s))
    _check(ctx, oracle, pool.fasta_items(), contigs, search.MODE_NT, min_id=0.3, min_cov=40, min_ratio=0.05)
    _check(ctx, oracle, pool.fasta_items(), contigs, search.MODE_PROT6, min_id=0.3, min_cov=40, min_ratio=0.05, gtable=4)
    # a query set with nothing to find
    junk = [('j%d' % i, ''.join(rng.choice(list('ACGT'), size=300).tolist())) for i in range(5)]
    hits, st = _check(ctx, oracle, junk, contigs, search.MODE_NT, min_id=0.4, min_cov=50, min_ratio=0.25)
    assert len(hits) == 0


def test_transeq_device_matches_reference_golden(ctx):
    """seqcodec.transeq (pb_transeq on the device) against the vectors generated from modules/configure.py:160-194"""
    import json, os
    cases = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'transeq.json')))
    n = 0
    for c in cases:
        got = seqcodec.transeq({'a': c['seq']}, frame=c['frame'], transl_table=c['table'], ctx=ctx)['a']
        assert got == c['out'], (c['seq'][:30], c['frame'], c['table'])
        n += len(got)
    assert n > 300
    # batched call, list container, markStarts
    seqs = [(str(i), c['seq']) for i, c in enumerate(cases) if c['frame'] == '7' and c['table'] == 11]
    got = seqcodec.transeq(seqs, frame='7', transl_table=11, ctx=ctx)
    want = [c['out'] for c in cases if c['frame'] == '7' and c['table'] == 11]
    assert [g[1] for g in got] == want and [g[0] for g in got] == [s[0] for s in seqs]
    assert seqcodec.transeq({'a': 'TAATAGTGAATGGTGTTGCTGATTATCATA'}, frame='1', transl_table=11, markStarts=True, ctx=ctx)['a'] == ['XXXMMMLIII']
    assert seqcodec.transeq({'a': 'ATGNNNAC-GTA'}, frame='1', ctx=ctx)['a'] == ['MX-V']
    assert seqcodec.transeq({'a': 'TAATAGTGAATGGTGTTGCTG'}, frame='1', transl_table=4, ctx=ctx)['a'] == ['XXWMVLL']
    assert seqcodec.transeq({}, frame='7', ctx=ctx) == {}


@pytest.mark.gpu
def test_grouped_search_in_two_halves_and_views_of_the_library_buffers(ctx):
    """search_grouped_local + take_hits (the calls worker threads and the exchanging thread make) give the table of
    search_grouped_raw, also when the arrays are views of the library's buffers (copy=False)."""
    import gc
    pool = workloads.GenePool(40, 40, seed=workloads.SEED + 5)
    seqs = [workloads.synth_genome(pool, g, n_acc_per_genome=20, seed=workloads.SEED + 5)[0] for g in range(3)]
    qn, qb, qo = seqio.to_seqset(pool.fasta_items()); tn, tb, to = seqio.to_seqset([('g%d' % i, s) for i, s in enumerate(seqs)])
    groups = np.arange(3, dtype=np.int32)
    for mode in (search.MODE_NT, search.MODE_PROT6):
        h0, c0, g0, _ = search.search_grouped_raw(ctx, qb, qo, tb, to, groups, mode, min_id=0.4, min_cov=50, min_ratio=0.25)
        for copy in (True, False):
            out, goff, st = search.search_grouped_local(ctx, qb, qo, tb, to, groups, mode, min_id=0.4, min_cov=50, min_ratio=0.25)
            h, c, roff = search.take_hits(ctx, out, False, copy=copy)
            assert roff is None and np.array_equal(goff, g0) and len(h) == len(h0) > 100
            assert h.tobytes() == h0.tobytes() and np.array_equal(c, c0)
            del h, c
            gc.collect()
