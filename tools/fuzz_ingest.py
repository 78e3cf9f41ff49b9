"""Differential fuzz of ingest.iter_readGFF against the REFERENCE'S OWN iter_readGFF / checkPseu (PEPPAN.py:117-182, 992-1010): the
annotation of a bundled genome with half of the features dropped, coordinates jittered, strands flipped, records duplicated; every
setting of incompleteCDS, both genetic tables, three minimum lengths.  Sequences, gene records and rejection codes must be equal.
Set-up from tools/fuzz_consumers.py.  Needs /root/reference.
    python tools/fuzz_ingest.py 0 10 >> profiles/r02_consumer_fuzz.txt"""
import os, sys, gzip, tempfile
_HERE = os.path.dirname(os.path.abspath(__file__))
exec(open(os.path.join(_HERE, 'fuzz_consumers.py')).read().split("bad = 0\nfor case in range")[0].split('\"\"\"', 2)[2].replace('os.path.dirname(os.path.dirname(os.path.abspath(__file__)))', repr(os.path.dirname(_HERE))))
from peppan_b200 import ingest
text = gzip.open(os.path.join(REF, 'examples', 'GCF_000214765.combined.gff.gz'), 'rt').read()
cut = text.find('\n>') + 1
gff, fasta = text[:cut].split('\n'), text[cut:]
feat = [l for l in gff if l and not l.startswith('#')]
bad = 0
for case in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(900 + case)
    lines = []
    for l in feat:
        if rng.random() < 0.5: continue
        p = l.split('\t')
        if len(p) > 8 and rng.random() < 0.3:
            r = rng.random()
            if r < 0.4: p[3] = str(max(1, int(p[3]) + int(rng.integers(-3, 4))))
            elif r < 0.7: p[4] = str(int(p[4]) + int(rng.integers(-3, 4)))
            elif r < 0.85: p[6] = '-' if p[6] == '+' else '+'
            else: p[3], p[4] = str(int(p[3])), str(int(p[3]) + int(rng.integers(10, 200)))
        lines.append('\t'.join(p))
        if rng.random() < 0.02: lines.append('\t'.join(p))          # duplicated record
    fn = os.path.join(tempfile.mkdtemp(prefix='fi%d_' % case), 'G%d.gff' % case)
    open(fn, 'w').write('##gff-version 3\n' + '\n'.join(lines) + '\n##FASTA\n' + fasta)
    inc = ['', 's', 'e', 'f', 'sef'][case % 5]; gt = [11, 4][case % 2]; mc = [120., 60., 300.][case % 3]
    P.params = dict(min_cds=mc, incompleteCDS=inc)
    try:
        s0, c0 = P.iter_readGFF((fn, 'CDS', gt)); e0 = None
    except Exception as e:
        e0 = type(e).__name__
    try:
        s1, c1 = ingest.iter_readGFF((fn, 'CDS', gt), min_cds=mc, incomplete=inc); e1 = None
    except Exception as e:
        e1 = type(e).__name__
    if e0 or e1:
        print('case', case, 'errors', e0, e1); bad += e0 != e1; continue
    ok = list(s0) == list(s1) and all(s0[k] == s1[k] for k in s0) and list(c0) == list(c1) and all(c0[k] == c1[k] for k in c0)
    codes = {}
    for c in c1.values(): codes[c[5] if not c[6] else 0] = codes.get(c[5] if not c[6] else 0, 0) + 1
    print('case', case, 'features', len(lines), 'genes', len(c0), 'by code', dict(sorted(codes.items())), 'incomplete', repr(inc), 'gtable', gt, 'ok' if ok else 'DIFF', flush=True)
    if not ok:
        for k in c0:
            if k not in c1 or c0[k] != c1[k]: print('  first diff', k, c0[k][:6], c1.get(k, [None]*6)[:6]); break
    bad += not ok
print('bad', bad)
