"""Real sequences at whole-genome scale (BASELINE.json configs[0], example.bash:2): the first 1,000 valid CDS of the bundled
E. coli genome GCF_000010485 against the complete genomes GCF_000214765 (105 contigs, ambiguous bases) and GCF_001566635
(4 contigs), all three search modes.  tests/golden/real_genomes.npz holds the sequences and the hit tables of the scalar
search oracle (tests/golden/make_real_genomes.py, run where /root/reference exists); the CUDA path must reproduce every
field of every hit and every CIGAR op."""
import os

import numpy as np
import pytest

from peppan_b200 import seqcodec

HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = {}


def _fx():
    if not _CACHE:
        with np.load(os.path.join(HERE, 'golden', 'real_genomes.npz')) as z:
            _CACHE.update({k: z[k] for k in z.files})
    return _CACHE


def _same(hits, cigar, ref, rcig):
    assert len(hits) == len(ref)
    for k in ref.dtype.names:
        if k in ('identity', 'evalue'):
            assert np.allclose(hits[k], ref[k], rtol=1e-6, atol=1e-30), k
        else:
            assert np.array_equal(hits[k], ref[k]), k
    assert np.array_equal(cigar, rcig)


def test_fixture_is_the_oracles_table(oracle):
    # pins the committed table to the oracle (one genome x the protein mode: ~10 s); also the shape of the workload
    fx = _fx()
    assert len(fx['q_off']) - 1 == 1000 and len(fx['g765_off']) - 1 == 105 and len(fx['g635_off']) - 1 == 4
    hits, cig = oracle.search(fx['q_bytes'], fx['q_off'], fx['g765_bytes'], fx['g765_off'], 2, seqcodec.BLOSUM62.reshape(-1),
                              min_id=0.4, min_cov=50, min_ratio=0.25, cap=2000000, cigar_cap=40000000)
    _same(hits, cig, fx['g765_m2_hits'], fx['g765_m2_cigar'])
    per_q = np.bincount(hits['q_id'], minlength=1000)
    assert per_q.max() >= 20 and (per_q > 0).sum() >= 950          # insertion sequences; nearly every gene has an ortholog


@pytest.mark.gpu
@pytest.mark.parametrize('genome', ['g765', 'g635'])
@pytest.mark.parametrize('mode', [1, 2])
def test_gpu_equals_oracle_on_whole_real_genomes(ctx, genome, mode):
    from peppan_b200 import search
    fx = _fx()
    hits, cigar, st = search.search(ctx, fx['q_bytes'], fx['q_off'], fx[genome + '_bytes'], fx[genome + '_off'], mode,
                                    min_id=0.4, min_cov=50, min_ratio=0.25)
    _same(hits, cigar, fx['%s_m%d_hits' % (genome, mode)], fx['%s_m%d_cigar' % (genome, mode)])
    assert st['kernel_launches'] > 0


@pytest.mark.gpu
def test_gpu_equals_oracle_on_real_self_search(ctx):
    from peppan_b200 import search
    fx = _fx()
    hits, cigar, st = search.search(ctx, fx['q_bytes'], fx['q_off'], fx['q_bytes'], fx['q_off'], 3, min_id=0.4, min_cov=50, min_ratio=0.25)
    _same(hits, cigar, fx['self_m3_hits'], fx['self_m3_cigar'])


@pytest.mark.gpu
def test_gpu_grouped_search_equals_per_genome_search(ctx):
    # both genomes in ONE call (target groups): the per-genome tables come back unchanged, contig ids local to their genome
    from peppan_b200 import search
    fx = _fx()
    tb = np.concatenate([fx['g765_bytes'], fx['g635_bytes']])
    to = np.concatenate([fx['g765_off'], fx['g635_off'][1:] + fx['g765_off'][-1]])
    groups = np.concatenate([np.zeros(105, np.int32), np.ones(4, np.int32)])
    for mode in (1, 2):
        res, st = search.search_grouped(ctx, fx['q_bytes'], fx['q_off'], tb, to, groups, mode, min_id=0.4, min_cov=50, min_ratio=0.25)
        assert len(res) == 2
        for (hits, cigar), genome in zip(res, ('g765', 'g635')):
            _same(hits, cigar, fx['%s_m%d_hits' % (genome, mode)], fx['%s_m%d_cigar' % (genome, mode)])
