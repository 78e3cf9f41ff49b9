#!/usr/bin/env python
"""Cuts a small REAL-sequence fixture out of the reference's bundled example genomes (examples/*.combined.gff.gz): the valid
CDS (reference's own iter_readGFF / checkPseu) of a 150 kb window of GCF_000010485 as queries, and the syntenic region of
GCF_001566635 (found with the search oracle) as target.  Writes tests/golden/real_slice.json.gz (sequence data only).
Run in the authoring container:  python tests/golden/make_real_slice.py"""
import gzip, json, os, stat, sys, tempfile, types
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import pb_oracle
from peppan_b200 import seqcodec, seqio
REF = os.environ.get('PEPPAN_REFERENCE', '/root/reference')
stubs = tempfile.mkdtemp(prefix='pb_stubs_')
for name in ('mmseqs', 'makeblastdb', 'diamond', 'blastn'):
    p = os.path.join(stubs, name); open(p, 'w').write('#!/bin/sh\nexit 0\n'); os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)
os.environ['PATH'] = stubs + os.pathsep + os.path.join(REF, 'dependencies') + os.pathsep + os.environ['PATH']
m = types.ModuleType('ete3'); m.Tree = object; sys.modules['ete3'] = m
sys.path.insert(0, REF); sys.dont_write_bytecode = True
import warnings; warnings.simplefilter('ignore')
import PEPPAN
PEPPAN.params = dict(min_cds=120., incompleteCDS='')
seqA, cdsA = PEPPAN.iter_readGFF((os.path.join(REF, 'examples', 'GCF_000010485.combined.gff.gz'), 'CDS', 11))
seqB, cdsB = PEPPAN.iter_readGFF((os.path.join(REF, 'examples', 'GCF_001566635.combined.gff.gz'), 'CDS', 11))
genes = [(n, c) for n, c in cdsA.items() if isinstance(c[6], str) and len(c[6]) >= 120 and 1000000 <= c[2] <= 1150000]
queries = [(str(i), c[6]) for i, (n, c) in enumerate(genes)]
contigs = [(n, s[1]) for n, s in seqB.items()]
qn, qb, qo = seqio.to_seqset(queries); tn, tb, to = seqio.to_seqset(contigs)
hits, _ = pb_oracle.search(qb, qo, tb, to, 1, seqcodec.BLOSUM62.reshape(-1), min_id=0.8, min_cov=100, min_ratio=0.5, cap=2000000, cigar_cap=40000000)
best = np.bincount(hits['s_id']).argmax()
pos = np.minimum(hits['s_start'], hits['s_end'])[hits['s_id'] == best]
lo, hi = int(np.percentile(pos, 5)) - 5000, int(np.percentile(pos, 95)) + 8000
lo = max(0, lo); target = contigs[best][1][lo:hi]
out = dict(source='GCF_000010485 CDS with start in [1.0, 1.15] Mbp vs GCF_001566635 %s[%d:%d]' % (tn[best], lo, hi),
           queries=queries, target=[('slice', target)])
with gzip.open(os.path.join(HERE, 'real_slice.json.gz'), 'wt') as f:
    json.dump(out, f)
print(len(queries), 'queries', sum(len(s) for _, s in queries), 'nt; target slice', len(target), 'bp')
