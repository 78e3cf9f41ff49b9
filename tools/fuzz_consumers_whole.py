"""The N4 chain on WHOLE real genomes against the REFERENCE'S OWN iter_map_bsn (PEPPAN.py:759-867): 1,000 real CDS of GCF_000010485
as exemplars, the complete bundled genomes GCF_000214765 (105 contigs) / GCF_001566635 (4 contigs) with their REAL annotation as old
predictions; both functions call this repository's uberBlast() with the hit tables of the committed oracle fixture
(tests/golden/real_genomes.npz) handed in (uberBlast(tables=...)).  The saved arrays must be equal cell for cell.  Needs /root/reference.
    python tools/fuzz_consumers_whole.py >> profiles/r02_consumer_fuzz.txt"""
import os, sys, tempfile, time
_HERE = os.path.dirname(os.path.abspath(__file__))
exec(open(os.path.join(_HERE, 'fuzz_consumers.py')).read().split("bad = 0\nfor case in range")[0].split('"""', 2)[2].replace('os.path.dirname(os.path.dirname(os.path.abspath(__file__)))', repr(os.path.dirname(_HERE))))
from peppan_b200 import ingest
with np.load(os.path.join(ROOT, 'tests', 'golden', 'real_genomes.npz')) as z:
    fx = {k: z[k] for k in z.files}
cut = lambda b, o: [b[o[i]:o[i + 1]].tobytes().decode() for i in range(len(o) - 1)]
tmp = tempfile.mkdtemp(prefix='fw_')
clust = os.path.join(tmp, 'exemplar.fa')
with open(clust, 'w') as f:
    for i, x in enumerate(cut(fx['q_bytes'], fx['q_off'])):
        f.write('>%d\n%s\n' % (i + 1, x))
real_ub = ub.uberBlast
bad = 0
for tag, acc in (('g635', 'GCF_001566635'), ('g765', 'GCF_000214765')):
    seqs = cut(fx[tag + '_bytes'], fx[tag + '_off'])
    contigs = [(7001 + i, s) for i, s in enumerate(seqs)]
    tables = {1: (fx[tag + '_m1_hits'], fx[tag + '_m1_cigar']), 2: (fx[tag + '_m2_hits'], fx[tag + '_m2_cigar'])}
    seqB, cds = ingest.iter_readGFF((os.path.join(REF, 'examples', acc + '.combined.gff.gz'), 'CDS', 11))
    names = list(seqB)
    assert [s[1] for s in seqB.values()] == seqs
    old = os.path.join(tmp, tag + '.old.npz'); st = P.MapBsn(old, 'w')
    for ci, cn in enumerate(names):
        rows = sorted([[50000 + k, c[2], c[3], c[4]] for k, c in enumerate(cds.values()) if c[1] == cn], key=lambda r: r[1])
        if rows:
            st._save(st.conn, str(7001 + ci), np.array(rows, dtype=object))
    st.conn.close()
    ortho = os.path.join(tmp, tag + '.ortho.npy')
    np.save(ortho, np.array([[i, i + 1, 9000 if i % 2 else -9000] for i in range(1, 900)], dtype=int), allow_pickle=True)
    for k, (ident, nod) in enumerate(((0.5, False), (0.8, False), (0.5, True))):
        params = dict(gtable=11, noDiamond=nod, match_identity=ident, match_frag_len=50., match_frag_prop=0.25, link_gap=600., link_diff=1.5,
                      match_prop=0.5, match_len=250., match_prop1=0.8, match_len1=100., match_prop2=0.4, match_len2=400.)
        P.uberBlast = lambda argv, *a, **kw: real_ub(argv, tables=tables)
        t0 = time.time()
        a = P.iter_map_bsn((os.path.join(tmp, tag + 'ref%d' % k), clust, 0, 'taxon', contigs, ortho, old, params)); t_ref = time.time() - t0
        t0 = time.time()
        b = consumers.iter_map_bsn((os.path.join(tmp, tag + 'ours%d' % k), clust, 0, 'taxon', contigs, ortho, old, params), store=P.MapBsn, tables=tables); t_our = time.time() - t0
        ra, rb = np.load(a + '.bsn.npz', allow_pickle=True), np.load(b + '.bsn.npz', allow_pickle=True)
        ok = ra['bsn'].shape == rb['bsn'].shape and same(rb['bsn'], ra['bsn']) and ra['ovl'].shape == rb['ovl'].shape and np.array_equal(ra['ovl'], rb['ovl'])
        print(acc, 'contigs', len(contigs), 'match_identity', ident, 'noDiamond', nod, 'groups', len(ra['bsn']), 'merge groups', sum(1 for g in ra['bsn'] if len(g[6]) > 1), 'overlaps', len(ra['ovl']),
              'on an annotated gene', sum(1 for g in ra['bsn'] for t in g[6] if t[10] > 0.9), 'ok' if ok else 'DIFF', 'reference %.1f s, ours %.1f s' % (t_ref, t_our), flush=True)
        bad += not ok
print('bad', bad)
