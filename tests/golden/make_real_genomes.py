#!/usr/bin/env python
"""Whole-genome REAL-sequence fixture (BASELINE.json configs[0], example.bash:2): the first 1,000 valid CDS of the bundled
E. coli genome GCF_000010485 -- extracted by the REFERENCE'S OWN iter_readGFF / checkPseu (PEPPAN.py:117-182, :992-1010) --
as queries, and the complete genomes GCF_000214765 (105 contigs) and GCF_001566635 (4 contigs) as targets, together with
the hit tables the scalar search oracle computes for them (nucleotide, protein 6-frame per genome; protein 3-frame self).
Writes tests/golden/real_genomes.npz (sequence bytes, offsets, expected hit tables + CIGARs).
Needs /root/reference; run in the authoring container:  python tests/golden/make_real_genomes.py"""
import os, stat, sys, tempfile, time, types
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import pb_oracle
from peppan_b200 import seqcodec, seqio
REF = os.environ.get('PEPPAN_REFERENCE', '/root/reference')
NQ = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
stubs = tempfile.mkdtemp(prefix='pb_stubs_')
for name in ('mmseqs', 'makeblastdb', 'diamond', 'blastn'):
    p = os.path.join(stubs, name); open(p, 'w').write('#!/bin/sh\nexit 0\n'); os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)
os.environ['PATH'] = stubs + os.pathsep + os.path.join(REF, 'dependencies') + os.pathsep + os.environ['PATH']
m = types.ModuleType('ete3'); m.Tree = object; sys.modules['ete3'] = m
sys.path.insert(0, REF); sys.dont_write_bytecode = True
import warnings; warnings.simplefilter('ignore')
import PEPPAN
PEPPAN.params = dict(min_cds=120., incompleteCDS='')

out = {}
seqA, cdsA = PEPPAN.iter_readGFF((os.path.join(REF, 'examples', 'GCF_000010485.combined.gff.gz'), 'CDS', 11))
genes = [(str(i), c[6]) for i, (n, c) in enumerate((n, c) for n, c in cdsA.items() if isinstance(c[6], str) and len(c[6]) >= 120)][:NQ]
qn, qb, qo = seqio.to_seqset(genes)
out['q_bytes'] = qb; out['q_off'] = qo
mat = seqcodec.BLOSUM62.reshape(-1)
for tag, acc in (('g765', 'GCF_000214765'), ('g635', 'GCF_001566635')):
    seqB, _ = PEPPAN.iter_readGFF((os.path.join(REF, 'examples', acc + '.combined.gff.gz'), 'CDS', 11))
    contigs = [(n, s[1]) for n, s in seqB.items()]
    tn, tb, to = seqio.to_seqset(contigs)
    out[tag + '_bytes'] = tb; out[tag + '_off'] = to
    for mode in (1, 2):
        t0 = time.time()
        hits, cig = pb_oracle.search(qb, qo, tb, to, mode, mat, min_id=0.4, min_cov=50, min_ratio=0.25, cap=2000000, cigar_cap=40000000)
        print(acc, 'mode', mode, len(hits), 'hits', len(cig), 'cigar ops, %.1f s' % (time.time() - t0), flush=True)
        out['%s_m%d_hits' % (tag, mode)] = hits; out['%s_m%d_cigar' % (tag, mode)] = cig
t0 = time.time()
hits, cig = pb_oracle.search(qb, qo, qb, qo, 3, mat, min_id=0.4, min_cov=50, min_ratio=0.25, cap=2000000, cigar_cap=40000000)
print('self mode 3', len(hits), 'hits, %.1f s' % (time.time() - t0))
out['self_m3_hits'] = hits; out['self_m3_cigar'] = cig
np.savez_compressed(os.path.join(HERE, 'real_genomes.npz'), **out)
print('written', os.path.getsize(os.path.join(HERE, 'real_genomes.npz')), 'bytes')
