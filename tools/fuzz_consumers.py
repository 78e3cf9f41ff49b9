"""Differential fuzz of the N4 consumer chain against the REFERENCE'S OWN iter_map_bsn (PEPPAN.py:759-867): random synthetic genomes
with ambiguous bases, a lower-case stretch, point indels (frame shifts), long insertions and duplicated gene pieces inside genes
(split hits that linearMerge joins, overlapping groups), 1-3 contigs, jittered old predictions, random ortholog signs, gtable 11 / 4,
with and without the protein search, several thresholds.  Both functions call this repository's uberBlast() with the search answered
by the oracle; the arrays they save must be equal cell for cell.  Needs /root/reference.
    python tools/fuzz_consumers.py 0 40 > profiles/r02_consumer_fuzz.txt"""
import os, sys, stat, tempfile, types, warnings, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/oracle'); sys.path.insert(0, ROOT+'/tests')
import numpy as np
REF = os.environ.get('PEPPAN_REFERENCE', '/root/reference')
stubs = tempfile.mkdtemp(prefix='pb_stubs_')
for name in ('mmseqs', 'makeblastdb', 'diamond', 'blastn'):
    p = os.path.join(stubs, name); open(p, 'w').write('#!/bin/sh\nexit 0\n'); os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)
os.environ['PATH'] = stubs + os.pathsep + os.environ['PATH']
m = types.ModuleType('ete3'); m.Tree = object; sys.modules['ete3'] = m
sys.path.insert(0, REF); warnings.simplefilter('ignore')
import PEPPAN as P
if not hasattr(np.lib.npyio, 'format'): np.lib.npyio.format = np.lib.format
import pb_oracle
from peppan_b200 import consumers, uberBlast as ub, workloads, seqcodec
def fake_search(ctx, qb, qo, rb, ro, mode, min_id=0.3, min_cov=40., min_ratio=0.05, gtable=11, max_hits=0, allgather=False):
    h, c = pb_oracle.search(qb, qo, rb, ro, mode, seqcodec.BLOSUM62.reshape(-1), min_id=min_id, min_cov=min_cov, min_ratio=min_ratio, gtable=gtable, max_hits=max_hits)
    return h, c, dict(kernel_launches=0)
ub._srch.search = fake_search; ub.get_context = lambda: None; ub.logger = lambda *a, **k: None
P.uberBlast = ub.uberBlast; P.logger = lambda *a, **k: None
def same(a, b):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        a, b = np.asarray(a), np.asarray(b)
        if a.shape != b.shape: return False
        if a.dtype == object or b.dtype == object: return all(same(x, y) for x, y in zip(a.reshape(-1), b.reshape(-1)))
        return a.dtype == b.dtype and np.array_equal(a, b)
    if isinstance(a, (list, tuple)): return len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
    return a == b
bad = 0
for case in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(1000 + case)
    tmp = tempfile.mkdtemp(prefix='fz%d_' % case)
    pool = workloads.GenePool(int(rng.integers(20, 50)), int(rng.integers(10, 40)), seed=workloads.SEED + 100 + case)
    seq, annot = workloads.synth_genome(pool, 0, n_acc_per_genome=int(rng.integers(5, 20)), seed=workloads.SEED + 100 + case)
    s = np.frombuffer(seq.encode(), dtype=np.uint8).copy()
    # ambiguous bases, a lower-case stretch, and point indels inside genes (frame shifts)
    for p in rng.integers(0, len(s), 30): s[p] = ord('N')
    a0 = int(rng.integers(0, len(s) - 3000)); s[a0:a0 + 2000] = np.frombuffer(s[a0:a0 + 2000].tobytes().lower(), dtype=np.uint8)
    seq = s.tobytes().decode()
    for k in range(6):
        g = annot[int(rng.integers(0, len(annot)))]; p = int(rng.integers(int(g[1]) + 30, int(g[2]) - 30))
        seq = seq[:p] + ('' if rng.random() < 0.5 else 'ACGT'[int(rng.integers(4))] * int(rng.integers(1, 3))) + seq[p + (int(rng.integers(1, 3)) if rng.random() < 0.5 else 0):]
    for k in range(8):                     # long insertions inside genes: split hits that linearMerge joins; duplicated gene pieces: overlaps
        g = annot[int(rng.integers(0, len(annot)))]; p = int(rng.integers(int(g[1]) + 60, int(g[2]) - 60))
        ins = ''.join('ACGT'[i] for i in rng.integers(0, 4, int(rng.integers(80, 500)))) if rng.random() < 0.6 else seq[int(g[1]):int(g[1]) + int(rng.integers(150, 400))]
        seq = seq[:p] + ins + seq[p:]
    ncut = int(rng.integers(1, 4)); cuts = sorted(int(x) for x in rng.integers(1000, len(seq) - 1000, ncut - 1))
    parts = [seq[a:b] for a, b in zip([0] + cuts, cuts + [len(seq)])]
    contigs = [(1001 + i, p) for i, p in enumerate(parts)]
    clust = os.path.join(tmp, 'ex.fa')
    with open(clust, 'w') as f:
        for n, x in pool.fasta_items(): f.write('>%s\n%s\n' % (n, x))
    old = os.path.join(tmp, 'old.npz'); st = P.MapBsn(old, 'w')
    off = 0
    for (cn, p) in contigs:
        rows = [[int(a[0]), int(a[1]) + 1 - off + int(rng.integers(-4, 5)), int(a[2]) - off + int(rng.integers(-4, 5)), '+' if a[3] > 0 else '-'] for a in annot if a[1] >= off and a[2] <= off + len(p) and rng.random() < 0.7]
        rows = [r for r in rows if r[1] >= 1]
        if rows: st._save(st.conn, str(cn), np.array(sorted(rows, key=lambda r: r[1]), dtype=object))
        off += len(p)
    st.conn.close()
    genes = sorted(set(int(a[0]) for a in annot))
    ortho = os.path.join(tmp, 'ortho.npy')
    np.save(ortho, np.array([[genes[i], genes[(i * 7 + 3) % len(genes)], int(rng.integers(-3, 4)) * 3000] for i in range(len(genes))], dtype=int), allow_pickle=True)
    params = dict(gtable=int(rng.choice([11, 11, 4])), noDiamond=bool(rng.random() < 0.3), match_identity=float(rng.choice([0.5, 0.65, 0.8, 0.9])), match_frag_len=float(rng.choice([40., 50., 80.])),
                  match_frag_prop=float(rng.choice([0.1, 0.25, 0.4])), link_gap=float(rng.choice([300., 600.])), link_diff=float(rng.choice([1.2, 1.5])),
                  match_prop=0.5, match_len=250., match_prop1=0.8, match_len1=100., match_prop2=0.4, match_len2=400.)
    t0 = time.time()
    try:
        a = P.iter_map_bsn((os.path.join(tmp, 'ref'), clust, 0, 'taxon', contigs, ortho, old, params))
        ra = np.load(a + '.bsn.npz', allow_pickle=True); ref_err = None
    except Exception as e:
        ref_err = '%s: %s' % (type(e).__name__, e)
    try:
        b = consumers.iter_map_bsn((os.path.join(tmp, 'ours'), clust, 0, 'taxon', contigs, ortho, old, params), store=P.MapBsn)
        rb = np.load(b + '.bsn.npz', allow_pickle=True); our_err = None
    except Exception as e:
        our_err = '%s: %s' % (type(e).__name__, e)
    if ref_err or our_err:
        print('case', case, 'ref_err', ref_err, 'our_err', our_err, params); bad += bool(our_err) != bool(ref_err); continue
    ok = ra['bsn'].shape == rb['bsn'].shape and same(rb['bsn'], ra['bsn']) and ra['ovl'].shape == rb['ovl'].shape and np.array_equal(ra['ovl'], rb['ovl'])
    print('case', case, 'groups', len(ra['bsn']), 'ovl', len(ra['ovl']), 'multi', sum(1 for g in ra['bsn'] if len(g[6]) > 1), 'ok' if ok else 'DIFF', {k: params[k] for k in ('gtable', 'noDiamond', 'match_identity')}, '%.0fs' % (time.time() - t0), flush=True)
    bad += not ok
print('bad', bad)
