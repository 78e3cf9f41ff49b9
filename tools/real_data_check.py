"""Real-sequence sanity check of the search specification (CPU: scalar search oracle; the GPU path reproduces its hit tables):
the first N valid CDS of one bundled E. coli ST131 genome (extracted by the REFERENCE'S OWN iter_readGFF / checkPseu,
PEPPAN.py:117-182) searched against another bundled genome, nucleotide + protein 6-frame, PEPPAN's iter_map_bsn thresholds.
Needs /root/reference (examples/*.combined.gff.gz); authoring-container only.  python tools/real_data_check.py [N]"""
import json, os, stat, sys, tempfile, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import pb_oracle
from peppan_b200 import seqcodec, seqio

REF = os.environ.get('PEPPAN_REFERENCE', '/root/reference')
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
stubs = tempfile.mkdtemp(prefix='pb_stubs_')
for name in ('mmseqs', 'makeblastdb', 'diamond', 'blastn'):
    p = os.path.join(stubs, name); open(p, 'w').write('#!/bin/sh\nexit 0\n'); os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)
os.environ['PATH'] = stubs + os.pathsep + os.path.join(REF, 'dependencies') + os.pathsep + os.environ['PATH']
m = types.ModuleType('ete3'); m.Tree = object; sys.modules['ete3'] = m
sys.path.insert(0, REF); sys.dont_write_bytecode = True
import warnings
warnings.simplefilter('ignore')
import PEPPAN
PEPPAN.params = dict(min_cds=120., incompleteCDS='')        # the module-level settings checkPseu reads (PEPPAN.py:992-1010)

ga, gb = [os.path.join(REF, 'examples', f) for f in ('GCF_000010485.combined.gff.gz', 'GCF_001566635.combined.gff.gz')]
t0 = time.time()
seqA, cdsA = PEPPAN.iter_readGFF((ga, 'CDS', 11))
seqB, cdsB = PEPPAN.iter_readGFF((gb, 'CDS', 11))
genes = [(n, c[6]) for n, c in cdsA.items() if isinstance(c[6], str) and len(c[6]) >= 120][:N]
contigs = [(n, s[1]) for n, s in seqB.items()]
print('parsed in %.1f s: %d query genes (%d nt), target %d contigs, %d bp' % (time.time() - t0, len(genes), sum(len(s) for _, s in genes),
                                                                              len(contigs), sum(len(s) for _, s in contigs)))
qn, qb, qo = seqio.to_seqset(genes); tn, tb, to = seqio.to_seqset(contigs)
out = {'queries': len(genes), 'target_bp': int(to[-1]), 'target_contigs': len(contigs)}
best = {}
for name, mode in (('nt', 1), ('prot6', 2)):
    t0 = time.time()
    hits, cig = pb_oracle.search(qb, qo, tb, to, mode, seqcodec.BLOSUM62.reshape(-1), min_id=0.4, min_cov=50, min_ratio=0.25, cap=2000000, cigar_cap=40000000)
    dt = time.time() - t0
    span = (hits['q_end'] - hits['q_start'] + 1) / hits['q_len']
    full = set(hits['q_id'][(span >= 0.8) & (hits['identity'] >= 0.9)].tolist())
    anyhit = set(hits['q_id'].tolist())
    per_q = np.bincount(hits['q_id'], minlength=len(genes))
    out[name] = {'hits': int(len(hits)), 'oracle_seconds': round(dt, 1), 'queries_with_a_hit': len(anyhit), 'queries_ge80pct_span_ge90pct_id': len(full),
                 'max_hits_of_one_query': int(per_q.max()), 'queries_with_more_than_10_hits': int((per_q > 10).sum())}
    best[name] = full
out['either_mode_ge80pct_span_ge90pct_id'] = len(best['nt'] | best['prot6'])
print(json.dumps(out, indent=1))
