"""Would a seed-score pre-filter cut the random protein seeds without losing hits?    Runs the scalar
search oracle in protein 6-frame mode with ORC_MIN_SEED_SCORE = T (seeds whose own 7-residue BLOSUM62 score is below T are
not extended; an experiment flag, NOT part of the specification the GPU implements) on the planted-identity genome of
tools/recall_curve.py (third-position-biased drift) and on the real-sequence fixture.  CPU only."""
import json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import gzip, json, os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'oracle')); sys.path.insert(0, os.path.join({root!r}, 'tools'))
import numpy as np, pb_oracle
from peppan_b200 import seqcodec, seqio
if {which!r} == 'planted':
    import recall_curve as rc
    rc.PER_LEVEL = 40
    res, n = rc.run(True, 3)
    print('RESULT' + json.dumps({{'found_by_level': {{('%.2f' % k): v for k, v in res['prot6'].items()}}}}))
else:
    d = json.load(gzip.open(os.path.join({root!r}, 'tests', 'golden', 'real_slice.json.gz'), 'rt'))
    qn, qb, qo = seqio.to_seqset([tuple(x) for x in d['queries']]); tn, tb, to = seqio.to_seqset([tuple(x) for x in d['target']])
    h, c = pb_oracle.search(qb, qo, tb, to, 2, seqcodec.BLOSUM62.reshape(-1), min_id=0.4, min_cov=50, min_ratio=0.25)
    print('RESULT' + json.dumps({{'hits': len(h), 'table': h.tobytes().hex()[:0] + str(hash(h.tobytes() + c.tobytes()))}}))
'''
out = {}
for which in ('planted', 'real'):
    out[which] = {}
    base = None
    for T in (-1000000, 10, 15, 20, 25):
        env = dict(os.environ, ORC_MIN_SEED_SCORE=str(T), ORC_SEED_STATS='1', PYTHONHASHSEED='0')
        p = subprocess.run([sys.executable, '-c', CODE.format(root=ROOT, which=which)], capture_output=True, text=True, env=env)
        st = re.findall(r'mode 2: seeds extended (\d+), skipped by the diagonal rule \d+, ungapped HSPs (\d+), below the seed-score threshold (\d+)', p.stderr)
        ext = sum(int(x[0]) for x in st); low = sum(int(x[2]) for x in st); hsp = sum(int(x[1]) for x in st)
        r = json.loads(p.stdout.split('RESULT')[1])
        r.update(seeds_extended=ext, seeds_filtered=low, ungapped_hsps=hsp)
        if base is None:
            base = r
        elif 'table' in r:
            r['hit_table_identical_to_unfiltered'] = r['table'] == base['table']
        r.pop('table', None) if T != -1000000 else None
        out[which]['none' if T < 0 else 'T=%d' % T] = r
    out[which]['none'].pop('table', None)
print(json.dumps(out, indent=1))
