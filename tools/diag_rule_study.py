"""What would a BLAST-style "diagonal already covered" rule change?    Runs the scalar search oracle
twice -- as specified, and with ORC_DIAG_COVERED=1 (a seed lying inside the HSP last produced on its (query, diagonal) is not
extended again) -- on a synthetic genome and on the real-sequence fixture, and reports the seeds extended / skipped, the
ungapped HSPs, and whether the final hit table changes.  CPU only; the rule is NOT part of the specification the GPU implements."""
import gzip, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import gzip, json, os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'oracle'))
import numpy as np, pb_oracle
from peppan_b200 import seqcodec, seqio, workloads
if {which!r} == 'synthetic':
    pool = workloads.GenePool(300, 600)
    seq, annot = workloads.synth_genome(pool, 0, n_acc_per_genome=150)
    q = pool.fasta_items(); t = [('ctg', seq)]
else:
    d = json.load(gzip.open(os.path.join({root!r}, 'tests', 'golden', 'real_slice.json.gz'), 'rt'))
    q = [tuple(x) for x in d['queries']]; t = [tuple(x) for x in d['target']]
qn, qb, qo = seqio.to_seqset(q); tn, tb, to = seqio.to_seqset(t)
out = {{}}
for mode in (1, 2):
    h, c = pb_oracle.search(qb, qo, tb, to, mode, seqcodec.BLOSUM62.reshape(-1), min_id=0.4, min_cov=50, min_ratio=0.25)
    out[mode] = [h.tobytes().hex(), c.tobytes().hex(), len(h)]
print('RESULT' + json.dumps(out))
'''
res = {}
for which in ('synthetic', 'real'):
    res[which] = {}
    runs = {}
    for rule in ('0', '1'):
        env = dict(os.environ, ORC_DIAG_COVERED=rule, ORC_SEED_STATS='1')
        p = subprocess.run([sys.executable, '-c', CODE.format(root=ROOT, which=which)], capture_output=True, text=True, env=env)
        stats = re.findall(r'mode (\d): seeds extended (\d+), skipped by the diagonal rule (\d+), ungapped HSPs (\d+)', p.stderr)
        tables = json.loads(p.stdout.split('RESULT')[1])
        runs[rule] = (stats, tables)
    for i, mode in enumerate(('1', '2')):
        s0 = [x for x in runs['0'][0] if x[0] == mode][0]; s1 = [x for x in runs['1'][0] if x[0] == mode][0]
        t0, t1 = runs['0'][1][mode], runs['1'][1][mode]
        res[which]['nt' if mode == '1' else 'prot6'] = {
            'seeds_extended': int(s0[1]), 'seeds_extended_with_rule': int(s1[1]), 'seeds_skipped_by_rule': int(s1[2]),
            'ungapped_hsps': int(s0[3]), 'ungapped_hsps_with_rule': int(s1[3]),
            'hits': t0[2], 'hits_with_rule': t1[2], 'hit_table_identical': t0[0] == t1[0] and t0[1] == t1[1]}
print(json.dumps(res, indent=1))
