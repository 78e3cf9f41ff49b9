"""GFF3 / FASTA ingest + exact-duplicate collapse (peppan_b200/ingest.py) against the REFERENCE'S OWN
iter_readGFF / checkPseu / writeGenes (PEPPAN.py:117-182, :992-1010, :1023-1039) on the bundled E. coli genomes (where
/root/reference exists), and on a hand-made GFF that exercises every rejection code (everywhere)."""
import gzip
import os
import stat
import sys
import tempfile
import types

import numpy as np
import pytest

from peppan_b200 import ingest

REF = os.environ.get('PEPPAN_REFERENCE', '/root/reference')
HAVE_REF = os.path.exists(os.path.join(REF, 'PEPPAN.py'))


@pytest.fixture(scope='module')
def PEPPAN():
    stubs = tempfile.mkdtemp(prefix='pb_stubs_')
    for name in ('mmseqs', 'makeblastdb', 'diamond', 'blastn'):
        p = os.path.join(stubs, name)
        with open(p, 'w') as f:
            f.write('#!/bin/sh\nexit 0\n')
        os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)
    os.environ['PATH'] = stubs + os.pathsep + os.path.join(REF, 'dependencies') + os.pathsep + os.environ['PATH']
    if 'ete3' not in sys.modules:
        m = types.ModuleType('ete3'); m.Tree = object
        sys.modules['ete3'] = m
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import PEPPAN as P
    P.params = dict(min_cds=120., incompleteCDS='')
    return P


@pytest.mark.skipif(not HAVE_REF, reason='reference checkout not present')
@pytest.mark.parametrize('acc', ['GCF_000214765', 'GCF_001577325'])
def test_ingest_equals_reference_on_bundled_genomes(PEPPAN, acc):
    fn = os.path.join(REF, 'examples', acc + '.combined.gff.gz')
    seq0, cds0 = PEPPAN.iter_readGFF((fn, 'CDS', 11))
    seq1, cds1 = ingest.iter_readGFF((fn, 'CDS', 11), min_cds=120., incomplete='')
    assert list(seq0) == list(seq1) and all(seq0[k] == seq1[k] for k in seq0)
    assert list(cds0) == list(cds1)
    for k in cds0:
        assert cds0[k] == cds1[k], k
    codes = [c[5] for c in cds1.values() if not c[6]]
    assert set(codes) >= {1, 2, 3, 4, 5}                  # every rejection reason occurs in real annotations
    # exact-duplicate collapse in priority order, against the reference's writeGenes
    genes = {n: c for n, c in cds1.items()}
    order = sorted(genes, key=lambda n: (-len(genes[n][6]), n))
    priority = {n: i for i, n in enumerate(order)}
    with tempfile.TemporaryDirectory() as tmp:
        f0, g0 = PEPPAN.writeGenes(os.path.join(tmp, 'a.fa'), cds0, priority)
        f1, g1 = ingest.write_genes(os.path.join(tmp, 'b.fa'), genes, priority)
        assert open(f0).read() == open(f1).read() and g0 == g1
    names, buf, off = ingest.genes_seqset(genes)
    assert len(names) == sum(1 for c in genes.values() if c[6]) and off[-1] == len(buf)


_GFF = '''##gff-version 3
ctg1\tx\tgene\t1\t300\t.\t+\t.\tID=gene0;locus_tag=L0
ctg1\tx\tCDS\t11\t160\t.\t+\t0\tID=cds0;Parent=gene0
ctg1\tx\tCDS\t201\t350\t.\t-\t0\tID=cds1;locus_tag=L1
ctg1\tx\tCDS\t401\t430\t.\t+\t0\tID=cds2;Name=short
ctg1\tx\tCDS\t451\t601\t.\t+\t0\tID=cds3;locus_tag=L3
ctg1\tx\tCDS\t651\t800\t.\t+\t0\tID=cds4;locus_tag=L4
ctg1\tx\tCDS\t851\t1000\t.\t+\t0\tID=cds5;locus_tag=L5
ctg1\tx\tCDS\t1051\t1200\t.\t+\t0\tID=cds6;locus_tag=L6
ctg2\tx\tCDS\t1\t150\t.\t+\t0\tID=cds7;locus_tag=L7
ctg9\tx\tCDS\t1\t150\t.\t+\t0\tID=cds8;locus_tag=L8
##FASTA
'''


def _toy(tmp_path):
    rng = np.random.default_rng(1)

    def orf(n, start='ATG', stop='TAA', inner=None):
        cod = []
        while len(cod) < n // 3 - 2:
            c = ''.join(rng.choice(list('ACGT'), 3))
            if c not in ('TAA', 'TAG', 'TGA'):
                cod.append(c)
        if inner:
            cod[5] = inner
        return start + ''.join(cod) + stop
    ctg1 = list(''.join(rng.choice(list('ACGT'), 1300)))

    def put(a, s):
        ctg1[a - 1:a - 1 + len(s)] = list(s)
    good = orf(150)
    put(11, good)                                    # L0: accepted, plus strand
    put(201, ingest.rc(good))                        # L1: accepted, minus strand, exact duplicate of L0
    put(401, orf(30))                                # short: code 1
    put(451, orf(150) + 'A')                         # L3: 151 nt, frameshift: code 2
    put(651, orf(150, start='CCC'))                  # L4: no start: code 3
    put(851, orf(150, stop='CCC'))                   # L5: no stop: code 4
    put(1051, orf(150, inner='TAG'))                 # L6: internal stop: code 5
    ctg2 = orf(150, start='GTG')                     # L7: alternative start GTG counts as M (markStarts)
    fa = '>ctg1 first contig\n' + '\n'.join(''.join(ctg1)[i:i + 70] for i in range(0, 1300, 70)) + '\n>ctg2\n' + ctg2.lower() + '\n'
    fn = os.path.join(tmp_path, 'toy.combined.gff.gz')
    with gzip.open(fn, 'wt') as f:
        f.write(_GFF + fa)
    return fn, good


def test_ingest_codes_on_a_toy_annotation(tmp_path):
    fn, good = _toy(tmp_path)
    seq, cds = ingest.iter_readGFF((fn, 'CDS', 11))
    assert list(seq) == ['toy:ctg1', 'toy:ctg2'] and len(seq['toy:ctg1'][1]) == 1300 and seq['toy:ctg2'][1].isupper()
    code = {k.split(':')[1]: (v[5] if not v[6] else 0) for k, v in cds.items()}
    assert code == {'L0': 0, 'L1': 0, 'short': 1, 'L3': 2, 'L4': 3, 'L5': 4, 'L6': 5, 'L7': 0, 'L8': 6}
    assert cds['toy:L0'][6] == good == cds['toy:L1'][6] and cds['toy:L0'][5] == cds['toy:L1'][5] and cds['toy:L1'][4] == '-'
    assert cds['toy:L0'][:5] == [fn, 'toy:ctg1', 11, 160, '+']
    priority = {n: i for i, n in enumerate(sorted(cds))}
    out, groups = ingest.write_genes(os.path.join(tmp_path, 'genes.fa'), cds, priority)
    assert groups == [['toy:L0', 'toy:L1', 10000]]
    assert [l[1:].strip() for l in open(out) if l.startswith('>')] == ['toy:L0', 'toy:L7']


@pytest.mark.gpu
def test_device_screen_equals_host_screen(ctx, tmp_path):
    # the pseudogene screen through pb_transeq (one batch per genome) gives the codes of the numpy statement
    fn, _ = _toy(tmp_path)
    s0, c0 = ingest.iter_readGFF((fn, 'CDS', 11))
    s1, c1 = ingest.iter_readGFF((fn, 'CDS', 11), ctx=ctx)
    assert c0 == c1 and s0 == s1
    s4, c4 = ingest.iter_readGFF((fn, 'CDS', 4), ctx=ctx)
    assert c4 == ingest.iter_readGFF((fn, 'CDS', 4))[1]
